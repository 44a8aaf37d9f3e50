#!/usr/bin/env python
"""Opcode histogram of selected kernels of libfus_b200.so (cuobjdump -sass): the evidence behind the
statements about instruction mix in DESIGN.md (REDG.E.ADD.F64 scatter, LDG.E.NA.128.CONSTANT stream
of G, DFMA counts, 256-bit LDG/STG in the epilogue, UBLKCP / SYNCS in the TMA-ring variant, no
tensor-core opcodes anywhere).

    python scripts/sass_histogram.py [lib] > profiles/r2_sass_opcodes.json
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fenicsx-fus_b200", "lib", "libfus_b200.so")
WANT = [r"stiffness_line_kernel<5, false, 0, double, false>", r"stiffness_line_kernel<5, false, 0, double, true>",
        r"stiffness_line_kernel<5, true, 0, double, false>", r"stiffness_line_kernel<6, false, 6, double, false>",
        r"stiffness_line_kernel<7, false, 6, double, false>", r"stiffness_line_kernel<8, false, 4, double, false>",
        r"stiffness_line_kernel<5, false, 7, double, false>", r"stiffness_col_kernel<3, false>",
        r"stiffness_col_kernel<4, false>", r"rk4_stage_kernel<1, false, false, false>",
        r"rk4_stage_kernel<1, false, false, true>", r"rk4_stage_kernel<3, true, false, true>"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kern, name = {}, None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1)
        kern[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and name:
        kern[name][m.group(1)] += 1
dem = subprocess.run(["c++filt"], input="\n".join(kern), capture_output=True, text=True).stdout.split("\n")
total = collections.Counter()
for c in kern.values():
    total.update(c)
res = {"library": os.path.relpath(lib, ROOT), "arch": arch, "kernels": len(kern),
       "whole_library": {k: v for k, v in total.items()
                         if re.match(r"REDG|LDG\.E\.NA\.128|DFMA|UBLKCP|SYNCS|UTCMMA|LDTM|UTMALDG|HMMA|DMMA|"
                                     r"LDG\.E[\w.]*256|STG\.E[\w.]*256|LDL|STL", k)},
       "selected": {}}
for mangled, d in zip(kern, dem):
    short = re.sub(r"\(.*", "", d).replace("void fus::", "")
    if any(w == short for w in WANT):
        c = kern[mangled]
        res["selected"][short] = {"instructions": sum(c.values()),
                                  "top": dict(c.most_common(14)),
                                  "spill_LDL_STL": sum(v for k, v in c.items() if k.startswith(("LDL", "STL")))}
json.dump(res, sys.stdout, indent=1)
print()
