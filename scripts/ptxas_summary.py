#!/usr/bin/env python
"""Registers / spills / shared memory per kernel from `python fenicsx-fus_b200/build.py --force --verbose`.

    python fenicsx-fus_b200/build.py --force --verbose 2> build.log; python scripts/ptxas_summary.py build.log [filter]
"""
import re
import subprocess
import sys

txt = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
names = re.findall(r"Compiling entry function '([^']+)'", txt)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
blocks = txt.split("Compiling entry function '")[1:]
for nm, blk in zip(dem, blocks):
    short = re.sub(r"\(.*", "", nm).replace("void fus::", "")
    if flt and flt not in short:
        continue
    regs = re.search(r"Used (\d+) registers", blk)
    spill = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", blk)
    print(f"{short:55s} regs {regs.group(1) if regs else '?':>4s}  stack {spill.group(1):>4s}  "
          f"spill st/ld {spill.group(2)}/{spill.group(3)}")
