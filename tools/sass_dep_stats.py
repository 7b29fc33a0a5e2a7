"""How well does ptxas interleave independent work?  For every FP64 instruction in the RK4 loop, distance (in
instructions) to the nearest producer of one of its sources.  usage: sass_dep_stats.py <dump> <function-substring>"""
import re, sys, collections
lines = open(sys.argv[1]).read().splitlines(); fn = sys.argv[2]
start = next(i for i, l in enumerate(lines) if "Function :" in l and fn in l)
end = next((i for i in range(start + 1, len(lines)) if "Function :" in lines[i]), len(lines))
ins = []
for l in lines[start:end]:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), re.sub(r"^@!?U?P\d+\s+", "", m.group(2))))
back = []
for a, t in ins:
    m = re.search(r"\bBRA\S*\s+(?:.*?)(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a: back.append(((a - int(m.group(1), 16)) // 16, int(m.group(1), 16), a))
back.sort(reverse=True); b = next(x for x in back if x[0] < 6000)
loop = [t for a, t in ins if b[1] <= a <= b[2]]
last_write = {}
hist = collections.Counter(); est = 0.0
for i, t in enumerate(loop):
    op = t.split()[0]
    regs = [int(x) for x in re.findall(r"(?<![U])R(\d+)", t)]
    if not regs: continue
    wide = op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "F2F.F64", "LDS.64", "MUFU.RCP64H", "MUFU.RSQ64H", "I2F.F64"))
    dst, srcs = regs[0], regs[1:]
    if op.startswith(("DFMA", "DMUL", "DADD")):
        dist = min([i - last_write[s] for s in srcs for s in (s, s + 1) if s in last_write] or [99])
        hist[min(dist, 9)] += 1
        est += max(0, 4 - dist)  # ~8-cycle latency / 2-cycle issue
    if not op.startswith(("ST", "BRA", "ISETP", "DSETP", "FSETP", "BAR")):
        last_write[dst] = i
        if wide: last_write[dst + 1] = i
tot = sum(hist.values())
print(f"{fn}: loop {len(loop)} instrs, {tot} FP64 arithmetic; producer distance histogram (instructions, 9 = >=9):")
print("  " + "  ".join(f"{d}:{hist[d]} ({100 * hist[d] / tot:.0f}%)" for d in sorted(hist)))
