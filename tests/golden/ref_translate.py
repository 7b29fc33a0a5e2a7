"""Dev-time helper (runs only in the build container, never on the GPU box).

Turns the reference's CasADi-SX model source text
(/root/reference/gl_gym/environments/models/aux_states.hpp `update()` and
ode.hpp `ODE()`) into plain-Python callables by a mechanical regex rewrite, so
that golden vectors for the right-hand side can be produced *from the
reference's own source* without CasADi.  Nothing from the reference is copied
into this repository: the text is read, rewritten in memory, exec'd, and only
numeric outputs are stored (see make_golden.py).

Recipe = SURVEY.md Appendix C.
"""
import math
import re

import numpy as np

REF = "/root/reference/gl_gym/environments/models"


def _strip(text):
    out = []
    for ln in text.splitlines():
        ln = re.sub(r"//.*$", "", ln)
        if ln.strip().startswith("#"):
            continue
        out.append(ln)
    return "\n".join(out)


def _body(text, start_pat, end_pat):
    s = re.search(start_pat, text)
    assert s, start_pat
    e = text.index(end_pat, s.end())
    return text[s.end():e]


def _rewrite(stmts):
    py = []
    for st in stmts.split(";"):
        st = " ".join(st.split())
        if not st:
            continue
        st = re.sub(r"\b([pxud])\((\d+)\)", r"\1[\2]", st)
        st = re.sub(r"\ba\((\d+)\)", r"a[\1]", st)
        st = re.sub(r"\bdxdt\((\d+)\)", r"dxdt[\1]", st)
        st = st.replace("M_PI", "math.pi")
        py.append(st)
    return py


def _f32(v):
    return float(np.float32(v))


def make_namespace():
    ns = {"math": math}
    c2k_f = _f32(273.15)  # `const float c2k = 273.15;` in airMv

    def _exp(v):
        try:
            return math.exp(v)
        except OverflowError:
            return math.inf

    ns.update(
        exp=_exp, sqrt=math.sqrt, fabs=abs, tanh=math.tanh, cos=math.cos,
        fmax=max, fmin=min,
        pow=lambda b, e: math.pow(b, e),
        if_else=lambda c, t, f: t if c else f,
        satVP=lambda t: 610.78 * _exp(17.2694 * t / (t + 238.3)),
        co2dens2ppm=lambda t, dens: 1e6 * 8.3144598 * (t + 273.15) * dens / (101325 * 44.01e-3),
        tau12=lambda t1, t2, r1d, r2u: t1 * t2 / (1. - r1d * r2u),
        rhoUp=lambda t1, r1u, r1d, r2u: r1u + (t1 * t1 * r2u) / (1. - r1d * r2u),
        rhoDn=lambda t2, r1d, r2u, r2d: r2d + (t2 * t2 * r1d) / (1. - r1d * r2u),
        degrees2rad=lambda dg: dg * math.pi / 180.,
        fir=lambda a1, e1, e2, f12, t1, t2, sg: a1 * e1 * e2 * f12 * sg * (math.pow(t1 + 273.15, 4.) - math.pow(t2 + 273.15, 4.)),
        sensible=lambda h, t1, t2: abs(h) * (t1 - t2),
        cond=lambda h, v1, v2: 1.0 / (1.0 + _exp(-0.1 * (v1 - v2))) * 6.4e-9 * h * (v1 - v2),
        smoothHar=lambda pv, co, sm, mr: mr * (math.tanh((2.0 * 4.6052 / sm) * (pv - co) / 2.0) + 1.0) / 2.0,
        airMv=lambda f12, v1, v2, t1, t2: 0.002165 * abs(f12) * (v1 / (t1 + c2k_f) - v2 / (t2 + c2k_f)),
        airMc=lambda f12, c1, c2: abs(f12) * (c1 - c2),
    )
    return ns


def load_reference_rhs():
    """Returns (aux(x,u,d,p)->list[239], ode(x,u,d,p)->list[28]) built from the reference source.

    The helper lambdas above restate the 12 inline helpers of aux_states.hpp:5-93;
    they are cross-checked against the header text by `check_helpers()`.
    """
    aux_txt = _strip(open(f"{REF}/aux_states.hpp").read())
    ode_txt = _strip(open(f"{REF}/ode.hpp").read())
    upd = _body(aux_txt, r"std::vector<SX> a\(239\);", "return vertcat(a);")
    ode = _body(ode_txt, r"SX dxdt = SX::zeros\(x\.size\(\)\);", "return dxdt;")
    upd_py = _rewrite(upd)
    ode_py = _rewrite(ode)
    assert len(upd_py) == 239 + 0 or len(upd_py) >= 239, len(upd_py)
    ns = make_namespace()
    src = ["def _aux(x, u, d, p):", "    a = [0.0] * 239"]
    src += ["    " + s for s in upd_py]
    src += ["    return a", "", "def _ode(x, u, d, p):", "    a = _aux(x, u, d, p)", "    dxdt = [0.0] * 28"]
    src += ["    " + s for s in ode_py]
    src += ["    return dxdt"]
    exec(compile("\n".join(src), "<reference-translation>", "exec"), ns)
    return ns["_aux"], ns["_ode"], len(upd_py), len(ode_py)


def check_helpers():
    """Asserts the constants used in the helper lambdas appear in the header text."""
    t = open(f"{REF}/aux_states.hpp").read()
    for tok in ["610.78", "17.2694", "238.3", "8.3144598", "44.01e-3", "101325", "6.4e-9",
                "0.002165", "const float c2k = 273.15", "2.0 * 4.6052 / smooth", "-0.1 * (vp1 - vp2)"]:
        assert tok in t, tok


if __name__ == "__main__":
    check_helpers()
    aux, ode, na, no = load_reference_rhs()
    print("statements:", na, no)
