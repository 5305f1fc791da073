"""Decode ptxas scoreboard control bits (stall / write-barrier / read-barrier / wait mask) of one kernel's SASS.
usage: sass_sb.py <lib.so> <kernel-name-substring> [lo hi]   (development helper; see B300_MICROARCH.md)"""
import re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
funs = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = funs.split("Function : ")
blk = [b for b in blocks if pat in b.split("\n")[0]]
assert blk, "no function matches"
txt = blk[0].split("\n")
print("#", txt[0][:150])
ins = []
i = 0
while i < len(txt):
    m = re.match(r'\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/', txt[i])
    if m:
        hi = re.search(r'/\* 0x([0-9a-f]{16}) \*/', txt[i + 1])
        word = (int(hi.group(1), 16) << 64) | int(m.group(3), 16)
        ins.append((int(m.group(1), 16), m.group(2).strip(), (word >> 105) & 0xf, (word >> 110) & 7, (word >> 113) & 7, (word >> 116) & 0x3f))
        i += 2
    else:
        i += 1
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi_ = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
print("# %d instructions" % len(ins))
for a, t in [(a, t) for a, t, *_ in ins]:
    m = re.search(r'BRA.*0x([0-9a-f]+)', t)
    if m and int(m.group(1), 16) < a:
        print("# backward branch %x -> %s (%d instrs)" % (a, m.group(1), (a - int(m.group(1), 16)) // 16))
for a, t, st, wb, rb, wm in ins:
    if lo <= a <= hi_:
        print("%5x st=%2d wb=%d rb=%d wait=%s  %s" % (a, st, wb, rb, format(wm, '06b'), t[:90]))
