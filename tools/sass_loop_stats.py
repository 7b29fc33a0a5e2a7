"""Opcode histogram of the RK4 loop (largest backward-branch span below 6000 instructions) of a kernel.
usage: python tools/sass_loop_stats.py <cuobjdump -sass dump> <function-substring>"""
import re, sys, collections
lines = open(sys.argv[1]).read().splitlines()
fn = sys.argv[2]
start = next(i for i, l in enumerate(lines) if "Function :" in l and fn in l)
end = next((i for i in range(start + 1, len(lines)) if "Function :" in lines[i]), len(lines))
ins = []
for l in lines[start:end]:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2)))
back = []
for a, t in ins:
    m = re.search(r"\bBRA\S*\s+(?:.*?)(0x[0-9a-f]+)", t)
    if m:
        tg = int(m.group(1), 16)
        if tg < a: back.append(((a - tg) // 16, tg, a))
back.sort(reverse=True)
b = next(x for x in back if x[0] < 6000)
loop = [t for a, t in ins if b[1] <= a <= b[2]]
h = collections.Counter()
for t in loop:
    t = re.sub(r"^@!?U?P\d+\s+", "", t); h[t.split()[0].split(".")[0]] += 1
fp64 = sum(h[k] for k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"kernel {len(ins)} instrs; loop {len(loop)} instrs; FP64-pipe {fp64}; MUFU {h['MUFU']}; local ld/st {h['LDL']}/{h['STL']}; smem ld/st {h['LDS']}/{h['STS']}")
print(h.most_common(28))
