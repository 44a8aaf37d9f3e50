#!/usr/bin/env python
"""Compare the SASS of every kernel in two builds of libfus_b200.so (whitespace/address-insensitive).

    python scripts/sass_compare.py old.so new.so [name-filter]

Used when a change must not touch an already-measured kernel (there is no GPU in the build
container): an identical instruction stream means the profile and parity results still apply.
"""
import re
import subprocess
import sys


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    res, name = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            res[name] = []
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", ln)
        if m:
            res[name].append(re.sub(r"\s+", " ", m.group(1)).strip())
    return res


def demangle(names):
    return subprocess.run(["c++filt"], input="\n".join(names), capture_output=True,
                          text=True).stdout.split("\n")


def norm(n):
    n = re.sub(r"\(.*", "", n).replace("void fus::", "")
    # bool template arguments print as true/false, ints as digits: make them comparable; a trailing
    # default scalar type (added when the line kernel was templated on it) is not part of the name
    n = n.replace("true", "1").replace("false", "0")
    # ... nor is the trailing HALO = false added with the fused halo exchange
    n = re.sub(r"(stiffness_line_kernel<[^>]*), double, 0>", r"\1>", n)
    return n.replace(", double>", ">")


a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
flt = sys.argv[3] if len(sys.argv) > 3 else ""
da = {norm(d): a[m] for m, d in zip(a, demangle(list(a)))}
db = {norm(d): b[m] for m, d in zip(b, demangle(list(b)))}
same = diff = 0
for k in sorted(set(da) | set(db)):
    if flt and flt not in k:
        continue
    if k not in da:
        print(f"new        {k} ({len(db[k])} instr)")
    elif k not in db:
        print(f"removed    {k}")
    elif da[k] == db[k]:
        same += 1
        if flt:
            print(f"identical  {k} ({len(da[k])} instr)")
    else:
        diff += 1
        print(f"DIFFERENT  {k} ({len(da[k])} -> {len(db[k])} instr)")
print(f"{same} identical, {diff} different")
