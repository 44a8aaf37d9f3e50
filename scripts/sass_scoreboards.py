"""Decode the scheduling control bits of a kernel's SASS (cuobjdump -sass): write/read barrier index and
wait mask per instruction -- which scoreboard every load signals and where the code waits on it.

    python scripts/sass_scoreboards.py <mangled kernel name> <library.so> [regex on the instruction text]

Used for DESIGN.md 3.1 (every global load of the cell loop sits on one scoreboard)."""
import re, subprocess, sys
fun, lib = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], capture_output=True, text=True).stdout.splitlines()
ins = []
i = 0
while i < len(out):
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", out[i])
    if m and i + 1 < len(out):
        m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", out[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ctrl = hi >> 41
            stall = ctrl & 0xf; yld = (ctrl >> 4) & 1; wr = (ctrl >> 5) & 7; rd = (ctrl >> 8) & 7
            wait = (ctrl >> 11) & 0x3f
            ins.append((m.group(1), m.group(2).strip(), stall, wr, rd, wait))
            i += 2
            continue
    i += 1
sel = sys.argv[3] if len(sys.argv) > 3 else "LDG|STS|LDS|REDG|BRA|WARPSYNC|BAR|MOV"
for a, s, stall, wr, rd, wait in ins:
    if re.search(sel, s):
        w = ",".join(str(b) for b in range(6) if wait >> b & 1)
        print(f"{a} wr={wr if wr != 7 else '-'} rd={rd if rd != 7 else '-'} wait=[{w}] st={stall:2d}  {s[:80]}")
