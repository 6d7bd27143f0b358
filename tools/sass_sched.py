"""Static schedule estimate of a SASS address range from `cuobjdump -sass` output: sums the per-instruction stall
counts of the control codes (bits 105..108 of each 128-bit instruction) and counts instructions per pipe class.

    cuobjdump -sass -fun <mangled> lib.so > k.sass ; python tools/sass_sched.py k.sass 0x1b50 0x5020
"""
import re, sys, collections
path, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
lines = open(path).read().split("\n")
ins = []
i = 0
pat = re.compile(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/")
pat2 = re.compile(r"/\* 0x([0-9a-f]{16}) \*/")
while i < len(lines):
    m = pat.search(lines[i])
    if m and i + 1 < len(lines):
        m2 = pat2.search(lines[i + 1])
        if m2:
            addr = int(m.group(1), 16); text = m.group(2).strip(); hiw = int(m2.group(1), 16)
            ctrl = (hiw >> 41) & 0x1FFFFF
            ins.append((addr, text, ctrl & 0xF, (ctrl >> 4) & 1, (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 0x3F))
            i += 2
            continue
    i += 1
sel = [x for x in ins if lo <= x[0] <= hi]
stall = sum(x[2] for x in sel)
ops = collections.Counter()
for a, t, *_ in sel:
    op = t.split()[0] if not t.startswith("@") else t.split()[1]
    ops[op.split(".")[0]] += 1
print(f"{len(sel)} instructions, sum of stall counts {stall}")
for k, v in ops.most_common():
    print(f"  {k:12s} {v}")
if len(sys.argv) > 4:
    for a, t, s, y, wb, rb, wm in sel:
        print(f"{a:05x} st={s:2d} y={y} wb={wb} rb={rb} wait={wm:06b}  {t}")
