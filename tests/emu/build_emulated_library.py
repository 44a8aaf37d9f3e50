"""TEST INFRASTRUCTURE: build the whole library for the CPU on top of the SIMT emulator.

    python tests/emu/build_emulated_library.py        ->  tests/emu/libfus_b200_emulated.so

The product sources are not touched.  fus_capi.cu and fus_halo.cu are copied into tests/emu/_gen/
with every `kernel<<<grid, block, smem, stream>>>(args)` rewritten into
`FUS_EMU_LAUNCH((kernel), grid, block, smem, stream, args)` and compiled as host C++ against
tests/emu/cuda_runtime.h (a synchronous stand-in for the CUDA runtime calls the library makes) and
tests/emu/simt_emu.hpp (the kernels on host threads).  The result exports the same C ABI as
libfus_b200.so, which lets `pytest -m gpu --emulated-device` drive the host plumbing of every entry
point -- context creation, options, models, the RK4 loop on eager issue, the FP32 and 2-D paths --
in the build container, where there is no GPU.  It is a checker like oracle/: only tests/ load it,
and only when asked to; the product has no CPU fallback.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "fenicsx-fus_b200", "csrc")
GEN = os.path.join(HERE, "_gen")
LIB = os.path.join(HERE, "libfus_b200_emulated.so")


def _match_back_angle(s, i):
    """s[i] == '>': index of the matching '<'."""
    depth = 0
    while i >= 0:
        if s[i] == ">":
            depth += 1
        elif s[i] == "<":
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced template arguments before <<<")


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(src):
    out, pos, count = "", 0, 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            return out + src[pos:], count
        # kernel expression: identifier, optionally followed by template arguments
        j = k - 1
        if src[j] == ">":
            j = _match_back_angle(src, j) - 1
        while j >= 0 and (src[j].isalnum() or src[j] in "_:"):
            j -= 1
        kexpr = src[j + 1:k]
        e = src.index(">>>", k)
        cfg = _split_top(src[k + 3:e])
        if len(cfg) != 4:
            raise ValueError(f"launch configuration with {len(cfg)} entries: {src[k:e + 3]}")
        a = e + 3
        while src[a].isspace():
            a += 1
        assert src[a] == "(", src[k:a + 20]
        depth, b = 0, a
        while True:
            if src[b] == "(":
                depth += 1
            elif src[b] == ")":
                depth -= 1
                if depth == 0:
                    break
            b += 1
        args = src[a + 1:b].strip()
        out += src[pos:j + 1] + f"FUS_EMU_LAUNCH(({kexpr}), {', '.join(cfg)}, {args})"
        pos = b + 1
        count += 1


def build(verbose=False):
    os.makedirs(GEN, exist_ok=True)
    gen = []
    for name in ("fus_capi.cu", "fus_halo.cu"):
        with open(os.path.join(CSRC, name)) as f:
            text, n = rewrite_launches(f.read())
        if verbose:
            print(f"{name}: {n} launches rewritten")
        dst = os.path.join(GEN, name.replace(".cu", ".emu.cpp"))
        with open(dst, "w") as f:
            f.write(f"// GENERATED from fenicsx-fus_b200/csrc/{name} by build_emulated_library.py\n" + text)
            if name == "fus_capi.cu":          # test-only export: how many graph replays ran
                f.write('\nextern "C" long long fus_emu_graph_launches(void) '
                        '{ return fus_emu::graph_launches(); }\n')
        gen.append(dst)
    cmd = ["/usr/bin/g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-DFUS_HOST_EMULATION=1",
           "-I" + HERE, "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-o", LIB] + gen + [
        os.path.join(CSRC, "fus_host.cpp"), os.path.join(CSRC, "fus_partition.cpp"), "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulated build failed:\n" + res.stdout + res.stderr[-6000:])
    return LIB


def stale():
    if not os.path.exists(LIB):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, f) for f in ("cuda_runtime.h", "simt_emu.hpp", "build_emulated_library.py")]
    deps.append(os.path.join(ROOT, "include", "fus_b200.h"))
    return os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps)


if __name__ == "__main__":
    print(build(verbose=True))
