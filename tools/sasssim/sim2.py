"""Timing model of kernel C's evaluation loops (glg_roles.cuh) from SASS control words -- see sim.py for the model.

Roles are found by their named barriers: a group loop contains `BAR.ARV 0x1` + `BAR.SYNC 0x2`, the owner loop `BAR.SYNC 0x1` +
`BAR.ARV 0x2`.  A group loop reads its own copy of the stop flag (s_stop[W]): the immediate offset of that 32-bit LDS orders
the loops by group warp; the loop whose flag address has no per-warp immediate is the shared surface role (three warps,
given with --surface w,w,w; default 8,9,11 = the NG = 12 assignment).  Warp w of the CTA runs on sub-partition w % 4; owner
warps are warps 0..NO-1.

usage: python tools/sasssim/sim2.py <sass dump> <function substring> [--ctas N] [--list]
"""
import re
import sys
from parse import functions, parse
from sim import is_fp64, lat_of, FP64_RT

NO = 4
TIMELINE = False
SMEM_RT = 2.0  # cycles of the shared-memory pipe per 64-bit warp access (256 B at 128 B/clk)


def loops_of(ins):
    byaddr = {x.addr: k for k, x in enumerate(ins)}
    out = []
    for k, x in enumerate(ins):
        if x.op.startswith("BRA"):
            m = re.search(r"(0x[0-9a-f]+)", x.text)
            if m:
                tg = int(m.group(1), 16)
                if tg < x.addr and tg in byaddr:
                    body = ins[byaddr[tg]:k + 1]
                    bars = [y.text for y in body if y.op.startswith("BAR")]
                    if any("BAR.ARV 0x1" in b for b in bars) and not any("BAR.ARV 0x2" in b for b in bars):
                        out.append(("group", byaddr[tg], k))
                    elif any("BAR.ARV 0x2" in b for b in bars) and not any("BAR.ARV 0x1" in b for b in bars):
                        out.append(("owner", byaddr[tg], k))
    return out


def linear_path(ins, a, b, skip_min=24):
    """loop body as executed on the common path: a forward conditional branch is taken when it jumps over at least skip_min
    instructions (the rare blocks: micro-step decision, last-stage bookkeeping), unconditional forward branches always"""
    byaddr = {x.addr: k for k, x in enumerate(ins)}
    out, k = [], a
    while k <= b:
        x = ins[k]
        out.append(k)
        if x.op.startswith("BRA") and k != b:
            m = re.search(r"(0x[0-9a-f]+)", x.text)
            tg = byaddr.get(int(m.group(1), 16)) if m else None
            if tg is not None and tg > k and tg <= b and (not x.pred or tg - k >= skip_min):
                k = tg
                continue
        k += 1
    return out


class Warp:
    def __init__(self, wid, cta, role, ins, path):
        self.wid, self.cta, self.role, self.ins, self.path = wid, cta, role, ins, path
        self.pc, self.ready, self.sb = 0, 0.0, [0.0] * 6
        self.wait_bar = None
        self.iters = 0
        self.t_iter = []

    def cur(self):
        return self.ins[self.path[self.pc]]


def simulate(ins, roles, n_ctas=1, iters=6, owner_first=True):
    """roles: list of (kind, path) in warp order (owners first).  Returns cycles per evaluation in steady state."""
    warps = []
    for c in range(n_ctas):
        for w, (kind, path) in enumerate(roles):
            warps.append(Warp(w, c, kind, ins, path))
    nthreads_w = len(roles)
    # barrier state per CTA: arrivals[id] count of warps
    arrived = [{1: 0, 2: 0} for _ in range(n_ctas)]
    fp64_free = [0.0] * 4
    smem_free = 0.0
    t = 0.0
    busy64 = 0.0
    issued = 0
    done_iters = 0
    marks = []
    while True:
        for sp in range(4):
            best = None
            for w in sorted((w for w in warps if w.wid % 4 == sp and w.wait_bar is None), key=lambda w: (-w.wid, w.cta)):
                x = w.cur()
                if w.ready > t:
                    continue
                if x.wait and max(w.sb[s] for s in range(6) if (x.wait >> s) & 1) > t:
                    continue
                if is_fp64(x.op) and fp64_free[sp] > t:
                    continue
                if x.op.startswith(("LDS", "STS")) and smem_free > t + 8:  # shallow queue in front of the shared-memory pipe
                    continue
                best = w
                break
            if best is None:
                continue
            w = best
            x = w.cur()
            issued += 1
            if is_fp64(x.op):
                fp64_free[sp] = t + FP64_RT
                busy64 += FP64_RT
            lat = lat_of(x.op)
            if x.op.startswith(("LDS", "STS")):
                start = max(smem_free, t)
                smem_free = start + (SMEM_RT if ".64" in x.op else SMEM_RT / 2)
                lat = lat + (start - t)
            if x.wbar < 6:
                w.sb[x.wbar] = max(w.sb[x.wbar], t + lat)
            if x.rbar < 6:
                w.sb[x.rbar] = max(w.sb[x.rbar], t + 8)
            w.ready = t + max(1, x.stall)
            if x.op.startswith("BAR"):
                bid = 1 if " 0x1," in x.text else 2
                arrived[w.cta][bid] += 1
                marks.append((t, w.cta, w.wid, w.role, "sync" if "BAR.SYNC" in x.text else "arrive", bid))
                if "BAR.SYNC" in x.text:
                    w.wait_bar = bid
            w.pc += 1
            if w.pc >= len(w.path):
                w.pc = 0
                w.iters += 1
                w.t_iter.append(t)
        for c in range(n_ctas):
            for bid in (1, 2):
                if arrived[c][bid] == nthreads_w:
                    arrived[c][bid] = 0
                    for w in warps:
                        if w.cta == c and w.wait_bar == bid:
                            w.wait_bar = None
                            w.ready = max(w.ready, t + 16)
        if min(w.iters for w in warps) >= iters:
            break
        t += 1.0
        if t > 2e6:
            raise RuntimeError("simulation did not finish (deadlock?)")
    ow = [w for w in warps if w.role == "owner"][0]
    cyc = (ow.t_iter[-1] - ow.t_iter[1]) / (len(ow.t_iter) - 2)
    if TIMELINE:
        t_end = ow.t_iter[-1]
        t_beg = ow.t_iter[-2]
        print("  timeline of the last evaluation (cycles after the owners' previous arrive):")
        for (tm, c, wid, role, what, bid) in marks:
            if t_beg - 50 <= tm <= t_end + 5 and c == 0:
                print(f"    {tm - t_beg:7.0f}  warp {wid:2d} {role:5s} {what:6s} bar {bid}")
    return cyc, busy64 / (4 * t), issued / (4 * t)


def single(ins, path):
    t, sb, n64 = 0.0, [0.0] * 6, 0
    for k in path:
        x = ins[k]
        if x.op.startswith("BAR"):
            continue
        if x.wait:
            t = max(t, max(sb[s] for s in range(6) if (x.wait >> s) & 1))
        if x.wbar < 6:
            sb[x.wbar] = max(sb[x.wbar], t + lat_of(x.op))
        if x.rbar < 6:
            sb[x.rbar] = max(sb[x.rbar], t + 8)
        n64 += is_fp64(x.op)
        t += max(1, x.stall)
    return t, n64


def main():
    dump, fn = sys.argv[1], sys.argv[2]
    n_ctas = int(sys.argv[sys.argv.index("--ctas") + 1]) if "--ctas" in sys.argv else 1
    global TIMELINE
    TIMELINE = "--timeline" in sys.argv
    for name, lines in functions(dump):
        if fn in name:
            ins = parse(lines)
            lp = loops_of(ins)
            groups = [(k, a, b) for k, a, b in lp if k == "group"]
            surf_warps = [int(v) for v in (sys.argv[sys.argv.index("--surface") + 1] if "--surface" in sys.argv else "8,9,11").split(",")]

            def flag_off(a, b):
                offs = [int(m.group(1), 16) for x in ins[a:b + 1] if x.op == "LDS" for m in [re.search(r"\+(0x[0-9a-f]+)\]", x.text)] if m]
                return offs[-1] if offs else -1
            keyed = sorted(groups, key=lambda g: flag_off(g[1], g[2]))
            offs = [flag_off(g[1], g[2]) for g in keyed]
            # the shared loop reads *stop through a register (smallest / missing immediate); the others are 4 bytes apart per warp
            base = min(o for o in offs if o >= 0)
            ng = len(keyed) - 1 + len(surf_warps) if len(keyed) < 12 else len(keyed)
            by_warp = {}
            shared = None
            for g, o in zip(keyed, offs):
                w = (o - base) // 4 if o >= 0 else None
                if w is None or w in by_warp or (len(keyed) < 12 and shared is None and w == 0 and 0 in surf_warps):
                    shared = g
                else:
                    by_warp[w] = g
            if shared is None and len(keyed) < ng:
                shared = keyed[0]
            if shared is not None:
                # flag offsets of the generic loops are relative to s_stop: recover the absolute warp index
                rest = sorted(set(range(ng)) - set(surf_warps))
                by_warp = {w: g for w, g in zip(rest, [g for g in keyed if g is not shared])}
                for w in surf_warps:
                    by_warp[w] = shared
            groups = [by_warp[w] for w in sorted(by_warp)]
            owners = [(k, a, b) for k, a, b in lp if k == "owner"]
            lo = min(a for _, a, _ in lp)
            hi = max(b for _, _, b in lp)
            print(f"{name}\n  {len(groups)} group loops + {len(owners)} owner loop(s); loop code {hex(ins[lo].addr)}..{hex(ins[hi].addr)} = "
                  f"{(ins[hi].addr - ins[lo].addr + 16) / 1024:.1f} KB")
            tot64 = 0
            roles = []
            opath = linear_path(ins, owners[0][1], owners[0][2])
            to, no = single(ins, opath)
            for o in range(NO):
                roles.append(("owner", opath))
            print(f"  owner loop: {len(opath)} instr, {no} FP64, alone {to:.0f} cyc   (x{NO} warps)")
            tot64 += NO * no
            for g, (_, a, b) in enumerate(groups):
                p = linear_path(ins, a, b)
                tg, ng = single(ins, p)
                nl = sum(1 for k in p if ins[k].op.startswith("LDS"))
                ns = sum(1 for k in p if ins[k].op.startswith("STS"))
                nloc = sum(1 for k in p if ins[k].op.startswith(("LDL", "STL")))
                tot64 += ng
                print(f"  group warp {g:2d} (sub-partition {(g + NO) % 4}): {len(p):4d} instr, {ng:3d} FP64, {nl:2d} LDS {ns:2d} STS {nloc} local, alone {tg:5.0f} cyc")
                # a rotated loop starts with its BAR.SYNC (the first evaluation is peeled in front of it): start the model after it
                bars = [i for i, k in enumerate(p) if ins[k].op.startswith("BAR")]
                if bars and "BAR.SYNC" in ins[p[bars[0]]].text:
                    p = p[bars[0] + 1:] + p[:bars[0] + 1]
                roles.append(("group", p))
            per_sp = [0] * 4
            for w, (kind, p) in enumerate(roles):
                per_sp[w % 4] += sum(1 for k in p if is_fp64(ins[k].op))
            print(f"  FP64 per evaluation {tot64}; per sub-partition {per_sp} -> pipe floor {max(per_sp) * FP64_RT:.0f} cycles")
            cyc, b64, iss = simulate(ins, roles, n_ctas=n_ctas)
            print(f"  simulated ({n_ctas} CTA/SM): {cyc:.0f} cycles per evaluation per CTA; FP64 pipe busy {b64:.2f}; issue {iss:.2f}")
            return
    print("function not found")


if __name__ == "__main__":
    main()
