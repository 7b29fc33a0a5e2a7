"""ctypes binding of the CPU parity oracle (oracle/libglg_oracle.so).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(_ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libglg_oracle.so")
_DP = C.POINTER(C.c_double)


class EnvCfg(C.Structure):
    _fields_ = [("dt", C.c_double), ("n_sub", C.c_int), ("N", C.c_int), ("Np", C.c_int),
                ("delta_u_max_f32", C.c_double), ("u_min", C.c_double * 6), ("u_max", C.c_double * 6),
                ("con_low", C.c_double * 3), ("con_high", C.c_double * 3),
                ("elec_price", C.c_double), ("heating_price", C.c_double), ("co2_price", C.c_double),
                ("fruit_price", C.c_double), ("dmfm", C.c_double), ("uncertainty_scale", C.c_double),
                ("fixed_costs", C.c_double), ("stiff_guard", C.c_int), ("obs_modules", C.c_int * 8)]


class Env(C.Structure):
    _fields_ = [("x", C.c_double * 28), ("x_prev", C.c_double * 28), ("u", C.c_double * 6),
                ("day_of_year", C.c_double), ("hour_of_day", C.c_double), ("timestep", C.c_int),
                ("terminated", C.c_int), ("weather", _DP), ("weather_rows", C.c_int), ("n_micro", C.c_long),
                ("jac", _DP), ("jac_valid", C.c_int), ("state_obs", _DP)]


def build():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        lib = C.CDLL(LIB)
        lib.glgo_aux_rhs.argtypes = [_DP] * 6
        lib.glgo_rhs.argtypes = [_DP] * 5
        lib.glgo_evalf.argtypes = [_DP, _DP, _DP, _DP, C.c_double, C.c_int, _DP]
        lib.glgo_evalf.restype = C.c_int
        lib.glgo_evalf_ex.argtypes = [_DP, _DP, _DP, _DP, C.c_double, C.c_int, C.c_int, _DP, C.POINTER(C.c_long)]
        lib.glgo_evalf_ex.restype = C.c_int
        lib.glgo_evalf_bdf.argtypes = [_DP, _DP, _DP, _DP, C.c_double, C.c_double, C.c_double, _DP, _DP, C.POINTER(C.c_int),
                                       C.POINTER(C.c_long)]
        lib.glgo_evalf_bdf.restype = C.c_int
        lib.glgo_evalf_batch.argtypes = [_DP, _DP, _DP, _DP, C.c_int, C.c_double, C.c_int, _DP, C.c_int, C.c_int]
        lib.glgo_evalf_batch.restype = C.c_int
        lib.glgo_init_state.argtypes = [_DP, _DP]
        lib.glgo_env_reset.argtypes = [C.POINTER(Env), _DP, C.c_int, C.c_double]
        lib.glgo_param_noise.argtypes = [_DP, _DP, _DP]
        lib.glgo_obs_dim.argtypes = [C.POINTER(EnvCfg)]
        lib.glgo_obs_dim.restype = C.c_int
        lib.glgo_env_obs.argtypes = [C.POINTER(EnvCfg), C.POINTER(Env), _DP]
        lib.glgo_env_step.argtypes = [C.POINTER(EnvCfg), C.POINTER(Env), _DP, C.c_void_p, C.c_int, _DP, _DP, _DP, _DP]
        lib.glgo_env_step.restype = C.c_int
        lib.glgo_rule_control.argtypes = [_DP, _DP, _DP, C.c_double, C.c_double, _DP]
        lib.glgo_env_step_rule.argtypes = [C.POINTER(EnvCfg), C.POINTER(Env), _DP, _DP, _DP, _DP, _DP, _DP]
        lib.glgo_env_step_rule.restype = C.c_int
        lib.glgo_rollout.argtypes = [C.POINTER(EnvCfg), _DP, _DP, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                     C.c_int, _DP]
        lib.glgo_rollout.restype = C.c_long
        lib.glgo_batch_create.argtypes = [C.POINTER(EnvCfg), _DP, _DP, C.c_int, C.c_int]
        lib.glgo_batch_create.restype = C.c_void_p
        lib.glgo_batch_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.glgo_batch_destroy.argtypes = [C.c_void_p]
        lib.glgo_batch_work.argtypes = [C.c_void_p]
        lib.glgo_batch_work.restype = C.c_long
        _lib = lib
    return _lib


def P(a):
    return a.ctypes.data_as(_DP)


INTEGRATOR_RK4, INTEGRATOR_GRADED, INTEGRATOR_BDF, INTEGRATOR_BDF_KEEP_JAC = 0, 3, 16, 48  # glgo_env_cfg.stiff_guard


OBS_MODULE_IDS = {"StateObservations": 1, "IndoorClimateObservations": 2, "BasicCropObservations": 3, "ControlObservations": 4,
                  "WeatherObservations": 5, "TimeObservations": 6, "WeatherForecastObservations": 7}


def default_cfg(n_sub=600, N=5760, Np=48, dt=900.0, stiff_guard=0, obs_modules=None):
    c = EnvCfg()
    c.stiff_guard = int(stiff_guard)
    for i, m in enumerate(obs_modules or []):
        c.obs_modules[i] = OBS_MODULE_IDS[m] if isinstance(m, str) else int(m)
    c.dt, c.n_sub, c.N, c.Np = dt, n_sub, N, Np
    c.delta_u_max_f32 = float(np.float32(0.1))
    for i in range(6):
        c.u_min[i], c.u_max[i] = 0.0, 1.0
    lo, hi = (300.0, 15.0, 50.0), (1600.0, 34.0, 85.0)
    for i in range(3):
        c.con_low[i], c.con_high[i] = lo[i], hi[i]
    c.elec_price, c.heating_price, c.co2_price, c.fruit_price, c.dmfm = 0.3, 0.09, 0.3, 1.6, 0.065
    c.uncertainty_scale = 0.0
    c.fixed_costs = (15.0 + 0.015 + 0.07 * 116 + 2.0) / 365 / (86400 // 900)
    return c


def rhs(x, u, d, p):
    f = np.zeros(28)
    load().glgo_rhs(P(x), P(u), P(d), P(p), P(f))
    return f


def aux_rhs(x, u, d, p):
    a, f = np.zeros(239), np.zeros(28)
    load().glgo_aux_rhs(P(x), P(u), P(d), P(p), P(a), P(f))
    return a, f


def evalf(x, u, d, p, dt=900.0, n_sub=600):
    y = np.zeros(28)
    bad = load().glgo_evalf(P(np.ascontiguousarray(x, dtype=np.float64)), P(np.ascontiguousarray(u, dtype=np.float64)),
                            P(np.ascontiguousarray(d, dtype=np.float64)), P(np.ascontiguousarray(p, dtype=np.float64)),
                            dt, n_sub, P(y))
    return y, bad


def evalf_ex(x, u, d, p, dt=900.0, n_sub=600, stiff_guard=0):
    """-> (x_next, bad, RK4 micro-steps executed)"""
    y, n = np.zeros(28), C.c_long(0)
    bad = load().glgo_evalf_ex(P(np.ascontiguousarray(x, dtype=np.float64)), P(np.ascontiguousarray(u, dtype=np.float64)),
                               P(np.ascontiguousarray(d, dtype=np.float64)), P(np.ascontiguousarray(p, dtype=np.float64)),
                               float(dt), int(n_sub), int(stiff_guard), P(y), C.byref(n))
    return y, bool(bad), n.value


def evalf_bdf(x, u, d, p, dt=900.0, rtol=1e-6, atol=1e-6):
    """CVODES-class adaptive implicit solve (oracle/glg_oracle_bdf.c) -> (x_next, bad, {rhs, jac, lu, steps})"""
    y, st = np.zeros(28), (C.c_long * 4)(0, 0, 0, 0)
    bad = load().glgo_evalf_bdf(P(np.ascontiguousarray(x, dtype=np.float64)), P(np.ascontiguousarray(u, dtype=np.float64)),
                                P(np.ascontiguousarray(d, dtype=np.float64)), P(np.ascontiguousarray(p, dtype=np.float64)),
                                float(dt), float(rtol), float(atol), P(y), None, None, st)
    return y, bool(bad), dict(rhs=st[0], jac=st[1], lu=st[2], steps=st[3])


def evalf_batch(x, u, d, p, dt=900.0, n_sub=600, n_threads=0):
    x, u, d, p = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, u, d, p))
    B = x.shape[0]
    y = np.zeros((B, 28))
    stride = 0 if p.ndim == 1 else 208
    n_threads = n_threads or os.cpu_count() or 1
    load().glgo_evalf_batch(P(x), P(u), P(d), P(p), stride, dt, n_sub, P(y), B, n_threads)
    return y


def rule_control(settings29, x, d, hod, doy):
    u = np.zeros(6)
    load().glgo_rule_control(P(np.ascontiguousarray(settings29, dtype=np.float64)), P(np.ascontiguousarray(x, dtype=np.float64)),
                             P(np.ascontiguousarray(d, dtype=np.float64)), float(hod), float(doy), P(u))
    return u


class OracleEnv:
    """One reference-semantics env stepped by the C oracle."""

    def __init__(self, weather, p_nom, cfg=None, start_day=0.0):
        self.cfg = cfg or default_cfg()
        self.W = np.ascontiguousarray(weather, dtype=np.float64)
        self.p = np.ascontiguousarray(p_nom, dtype=np.float64)
        self.e = Env()
        self.start_day = start_day
        self.nobs = load().glgo_obs_dim(C.byref(self.cfg))
        self._state_obs = None
        self.reset()

    def reset(self):
        load().glgo_env_reset(C.byref(self.e), P(self.W), self.W.shape[0], self.start_day)
        obs = np.zeros(self.nobs)
        load().glgo_env_obs(C.byref(self.cfg), C.byref(self.e), P(obs))
        return obs

    def set_state_obs(self, values27):
        """The 27 StateObservations entries of the observations computed from now on (random numbers in the reference)."""
        self._state_obs = None if values27 is None else np.ascontiguousarray(values27, dtype=np.float64)
        self.e.state_obs = None if self._state_obs is None else P(self._state_obs)

    def step(self, action=None, control=None, noise34=None):
        obs, r, info = np.zeros(self.nobs), C.c_double(0.0), np.zeros(11)
        if control is not None:
            a = np.ascontiguousarray(control, dtype=np.float64)
            raw = 1
        else:
            a = np.ascontiguousarray(action, dtype=np.float32)
            raw = 0
        n = None if noise34 is None else P(np.ascontiguousarray(noise34, dtype=np.float64))
        done = load().glgo_env_step(C.byref(self.cfg), C.byref(self.e), P(self.p), a.ctypes.data, raw, n, P(obs),
                                    C.cast(C.byref(r), _DP), P(info))
        return obs, r.value, bool(done), info

    def step_rule(self, settings29, noise34=None):
        """controller in the loop: u = glgo_rule_control(x, weather[k], clock); step_raw_control(u)"""
        obs, r, info = np.zeros(self.nobs), C.c_double(0.0), np.zeros(11)
        s29 = np.ascontiguousarray(settings29, dtype=np.float64)
        n = None if noise34 is None else P(np.ascontiguousarray(noise34, dtype=np.float64))
        done = load().glgo_env_step_rule(C.byref(self.cfg), C.byref(self.e), P(self.p), P(s29), n, P(obs),
                                         C.cast(C.byref(r), _DP), P(info))
        return obs, r.value, bool(done), info

    @property
    def x(self):
        return np.array(self.e.x[:])

    @property
    def u(self):
        return np.array(self.e.u[:])


class OracleBatch:
    """B reference-semantics envs stepped by host threads (the measured CPU baseline)."""

    def __init__(self, weather, p_nom, B, cfg=None, n_threads=None):
        self.cfg = cfg or default_cfg()
        self.B, self.n_threads = int(B), int(n_threads or os.cpu_count() or 1)
        self.W = np.ascontiguousarray(weather, dtype=np.float64)
        self.p = np.ascontiguousarray(p_nom, dtype=np.float64)
        self.h = load().glgo_batch_create(C.byref(self.cfg), P(self.p), P(self.W), self.W.shape[0], self.B)
        self.reward = np.zeros(self.B)
        self.done = np.zeros(self.B, dtype=np.uint8)
        self.obs = np.zeros((self.B, load().glgo_obs_dim(C.byref(self.cfg))), dtype=np.float32)

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float32)
        load().glgo_batch_step(self.h, a.ctypes.data, self.obs.ctypes.data, self.reward.ctypes.data, self.done.ctypes.data,
                               self.n_threads)
        return self.obs, self.reward, self.done

    def work(self):
        """cumulative RK4 micro-steps (RK4 integrators) or right-hand-side evaluations (implicit solver) of all envs"""
        return int(load().glgo_batch_work(self.h))

    def close(self):
        if self.h:
            load().glgo_batch_destroy(self.h)
            self.h = None
