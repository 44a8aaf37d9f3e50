#!/usr/bin/env python
"""profiles/traffic.json from an ncu digest (scripts/ncu_digest.py output) of the headline stiffness
kernel: DRAM bytes read + written per launch, stamped with the content hash of the library sources
the capture was taken from.  bench.py reports `roofline.traffic` only when that hash is the hash of
the library it is running (a capture of another build is not a measurement of this one).

    python scripts/ncu_to_traffic.py profiles/r2p_ncu_full_stiffness_P4.json [kernel key]
"""
import importlib.util
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def to_bytes(text):
    m = re.match(r"([\d.,]+)\s*(\w+)", text)
    return float(m.group(1).replace(",", "")) * UNIT[m.group(2)]


def main():
    digest = sys.argv[1]
    key = sys.argv[2] if len(sys.argv) > 2 else "stiffness_line_kernel<5,false>@P4_box54"
    rec = json.load(open(digest))[0]
    rd, wr = to_bytes(rec["dram__bytes_read.sum"]), to_bytes(rec["dram__bytes_write.sum"])
    spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "fenicsx-fus_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {"source_hash": mod._source_hash(),
           "captured_with": f"{os.path.relpath(digest, ROOT)} (ncu --set full --clock-control none, one "
                            "launch inside bench.py --steps 2 --warmup 1)",
           key: {"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                 "kernel": rec["kernel"], "duration": rec.get("gpu__time_duration.sum")}}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
