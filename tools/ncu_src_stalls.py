"""Per-loop stall-sample summary from an ncu source page (csv): which role loop the samples fall into and why they stall.
usage: ncu -i rep --page source --csv > src.csv ; python tools/ncu_src_stalls.py src.csv [sass dump] [function substring]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, ins_ex = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: addr = int(r[ia], 16)
    except ValueError: continue
    st = {h: int(r[i] or 0) for i, h in stall_cols}
    data.append((addr, r[isrc], int(r[ins_ex] or 0), st))
base = data[0][0]
# split at barriers / backward branches into regions: report regions between BAR instructions
regions, cur = [], []
for d in data:
    cur.append(d)
    if d[1].strip().startswith(("BAR", "@")) and "BRA" in d[1] and "0x" in d[1]:
        pass
tot = collections.Counter()
for d in data:
    for h, v in d[3].items(): tot[h] += v
n = sum(tot.values())
print("all samples:", n, {h.replace('stall_', ''): f"{100 * v / n:.1f}%" for h, v in tot.most_common(9)})
# loops = ranges between a BAR.SYNC and the next backward BRA
loops = []
import re
for k, d in enumerate(data):
    m = re.search(r"BRA\S*\s+(0x[0-9a-f]+)", d[1])
    if m:
        tg = int(m.group(1), 16)
        if tg < d[0] - base and d[0] - base - tg > 0x200:
            loops.append((tg + base, d[0]))
for lo, hi in loops:
    c = collections.Counter(); ex = 0; ni = 0
    for d in data:
        if lo <= d[0] <= hi:
            ni += 1; ex = max(ex, d[2])
            for h, v in d[3].items(): c[h] += v
    s = sum(c.values())
    if s == 0: continue
    print(f"loop {lo - base:#07x}..{hi - base:#07x} ({ni:4d} instr, executed {ex:8d}x): {100 * s / n:5.1f}% of samples; " +
          " ".join(f"{h.replace('stall_', '')} {100 * v / s:.0f}%" for h, v in c.most_common(6)))
