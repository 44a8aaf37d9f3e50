#!/usr/bin/env python
"""Hottest lines of an `ncu --page source --csv` dump (stdin): the N rows with the most warp stall
samples, with their source/SASS text and the share of all samples.

    ncu -i report.ncu-rep --page source --csv | python scripts/ncu_source_hot.py 25
"""
import csv
import io
import sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
text = sys.stdin.read()
start = text.find('"')
rows = list(csv.reader(io.StringIO(text[start:]))) if start >= 0 else []
if len(rows) < 2:
    print("no source page in the report")
    sys.exit(0)
# the dump may open with metadata rows ("Kernel Name", ...): the header is the first row naming a
# sampling column
hrow = next((i for i, r in enumerate(rows) if any("Sampl" in c for c in r)), 0)
rows = rows[hrow:]
head = rows[0]
scol = next((i for i, h in enumerate(head) if "Sampling" in h and "All" in h), None)
if scol is None:
    scol = next((i for i, h in enumerate(head) if "Samples" in h), None)
tcol = next((i for i, h in enumerate(head) if h.strip() in ("Source", "SASS", "Instruction")), 1)
if scol is None:
    print("columns:", head)
    sys.exit(0)


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return 0.0


body = [r for r in rows[1:] if len(r) > max(scol, tcol)]
total = sum(num(r[scol]) for r in body) or 1.0
print(f"columns: {head[tcol]} | {head[scol]}; total samples {total:.0f}")
for r in sorted(body, key=lambda r: -num(r[scol]))[:n]:
    print(f"{100.0 * num(r[scol]) / total:6.2f} %  {r[0]:>6s}  {r[tcol][:150]}")
