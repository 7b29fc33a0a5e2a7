import re, sys
cur = None; rows = {}
for l in open(sys.argv[1]):
    if l.startswith('mask'): cur = l.split()[1]; rows[cur] = {}
    m = re.match(r'warp (\d): group (\d+)\s+barrier1 (\d+)\s+owner (\d+)\s+barrier2 (\d+)', l)
    if m: rows[cur][int(m.group(1))] = (int(m.group(2)), int(m.group(4)))
for k, v in rows.items():
    print(k, ' '.join(f"w{w}:{v[w][0]}/{v[w][1]}" for w in sorted(v)), ' total(w0)=', sum(v[0]) if 0 in v else None)
