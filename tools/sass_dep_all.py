"""Producer-distance histogram of FP64 arithmetic over a whole function (incl. its device subroutines), split at branches.
usage: sass_dep_all.py <dump> <function-substring>"""
import re, sys, collections
lines = open(sys.argv[1]).read().splitlines(); fn = sys.argv[2]
start = next(i for i, l in enumerate(lines) if "Function :" in l and fn in l)
end = next((i for i in range(start + 1, len(lines)) if "Function :" in lines[i]), len(lines))
ins = []
for l in lines[start:end]:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: ins.append(re.sub(r"^@!?U?P\d+\s+", "", m.group(2)))
last_write = {}; hist = collections.Counter()
for i, t in enumerate(ins):
    op = t.split()[0]
    if op.startswith(("BRA", "RET", "CALL", "BAR", "EXIT")): last_write = {}
    regs = [int(x) for x in re.findall(r"(?<![U])R(\d+)", t)]
    if not regs: continue
    dst, srcs = regs[0], regs[1:]
    if op.startswith(("DFMA", "DMUL", "DADD")):
        dist = min([i - last_write[s] for s0 in srcs for s in (s0, s0 + 1) if s in last_write] or [99])
        hist[min(dist, 9)] += 1
    if not op.startswith(("ST", "BRA", "ISETP", "DSETP", "FSETP", "BAR")):
        last_write[dst] = i; last_write[dst + 1] = i
tot = sum(hist.values())
print(f"{fn}: {len(ins)} instrs, {tot} FP64 arithmetic; producer distance histogram:")
print("  " + "  ".join(f"{d}:{hist[d]} ({100 * hist[d] / tot:.0f}%)" for d in sorted(hist)))
