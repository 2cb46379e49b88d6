"""Decode the scheduling control fields of sm_100a SASS (stall, write/read scoreboard slot, wait mask) next to each instruction.
usage: cuobjdump -sass -fun <mangled> file.o | python probe/sass_ctrl.py [regex]   (regex: only show matching mnemonics +-context)"""
import re, sys
pat = sys.argv[1] if len(sys.argv) > 1 else None
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lines = sys.stdin.read().split("\n")
out, i = [], 0
while i < len(lines):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
        if m2:
            w = (int(m2.group(1), 16) << 64) | int(m.group(3), 16)
            out.append((m.group(1), m.group(2).strip(), (w >> 105) & 0xF, (w >> 110) & 7, (w >> 113) & 7, (w >> 116) & 0x3F))
            i += 2
            continue
    i += 1
show = set()
for idx, o in enumerate(out):
    if pat is None or re.search(pat, o[1]):
        show.update(range(max(0, idx - ctx), min(len(out), idx + ctx + 1)))
last = -2
for idx in sorted(show):
    if idx != last + 1:
        print("  ...")
    a, ins, st, wb, rb, wm = out[idx]
    print(f"{a} {ins:60s} stall={st:2d} wbar={wb} rbar={rb} wait={wm:06b}")
    last = idx
