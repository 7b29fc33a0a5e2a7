"""Weather-table builder (hot-path row S9; host-side reset-path prep).

Same contract as the reference's `load_weather_data(weatherDataDir, location, source, growthYear, startDay,
nDays, predHorizon, h, nd)` (gl_gym/environments/utils.py:48-125): reads `<dir>/<location>/<source><year>.csv`
(5-minute rows), derives the 10 GreenLight disturbance columns and resamples them with a PCHIP interpolant onto
`ns = int(dt_raw/h * (Ns+Np))` points of `linspace(t0, t_end, ns)`.  The result is one float64 table
[ns, 10] per (location, source, year, start_day); the batched env uploads a bank of such tables to HBM once
and every env indexes its table by (table_id, timestep).

Columns (utils.py:75-84): 0 iGlob, 1 tOut, 2 vpOut, 3 co2Out [mg m-3], 4 wind, 5 tSky, 6 tSoOut, 7 DLI,
8 isDay, 9 isDaySmooth.

Quirks kept on purpose (they are part of the reference's observable behaviour):
  * the env passes `Np+1` *steps* (49) into the `predHorizon` argument documented in *days*
    (tomato_env.py:250-260) => 60+49 days are read and ns = 10464 for dt=900;
  * linspace over [time[0], time[-1]] with ns points gives a grid spacing of 900.086 s, not 900 s;
  * radiation below 1e-10 after interpolation is zeroed (utils.py:123).
Unlike the reference (which re-parses the CSV on every reset, ~0.3 s) parsed files are cached.
"""
from functools import lru_cache
from os.path import dirname, join

import numpy as np

DEFAULT_WEATHER_DIR = join(dirname(__file__), "data", "weather")
_SECS_DAY = 86400
_CO2_PPM_OUT = 400.0  # utils.py:88
_COLS = ("time", "global radiation", "wind speed", "air temperature", "sky temperature", "RH")


@lru_cache(maxsize=16)
def _read_raw(path):
    import pandas as pd

    df = pd.read_csv(path, sep=",")
    return {c: df[c].to_numpy(dtype=np.float64) for c in _COLS}


def sat_vp(temp):
    """Saturation vapour pressure [Pa] (utils.py:311-323)."""
    return 610.78 * np.exp(17.2694 * temp / (temp + 238.3))


def rh_to_vapor_dens(temp, rh):
    """utils.py:407-444"""
    R, C2K, Mw = 8.3144598, 273.15, 18.01528e-3
    pascals = (rh / 100) * sat_vp(temp)
    return pascals * Mw / (R * (temp + C2K))


def vapor_dens_to_pres(temp, dens):
    """utils.py:281-309"""
    rh = dens / rh_to_vapor_dens(temp, 100)
    return sat_vp(temp) * rh


def co2_ppm_to_dens(temp, ppm):
    """utils.py:325-350"""
    R, C2K, M_CO2, P = 8.3144598, 273.15, 44.01e-3, 101325
    return P * 10 ** -6 * ppm * M_CO2 / (R * (temp + C2K))


def co2_dens_to_ppm(temp, dens):
    """utils.py:352-361"""
    R, C2K, M_CO2, P = 8.3144598, 273.15, 44.01e-3, 101325
    return 1e6 * R * (temp + C2K) * dens / (P * M_CO2)


def vapor_pres_to_rh(temp, vp):
    """utils.py:363-364"""
    return np.clip(100 * vp / sat_vp(temp), a_min=0.0, a_max=100.0)


def soil_temp_nl(time_s):
    """utils.py:251-279"""
    secs_year = 3600 * 24 * 365
    return 10 + 5 * np.sin((2 * np.pi * (time_s + 0.625 * secs_year) / secs_year))


def daily_light_sum(time_s, rad):
    """DLI [MJ m-2 day-1]: for every sample the sum of radiation from the previous to the next midnight
    (utils.py:208-249; inclusive upper slice bound and the `+2` search offset reproduced)."""
    interval = time_s[1] - time_s[0]
    days = time_s / _SECS_DAY
    n = len(days)
    out = np.zeros(n)

    def next_midnight(start):
        hits = np.where(np.diff(np.floor(days[start:])) == 1)[0]
        return n if hits.size == 0 else int(hits[0]) + start

    before = 0
    after = next_midnight(0)
    after = after + 1 if after != n else n
    i = 0
    while i < n:
        seg_end = min(after, n)          # samples i .. after-1 share one sum
        out[i:seg_end] = np.sum(rad[before:after + 1])
        i = seg_end
        if i >= n:
            break
        before = after
        hits = np.where(np.diff(np.floor(days[before + 2:])) == 1)[0]
        after = n if hits.size == 0 else int(hits[0]) + before + 2
    return out * interval * 1e-6


def compute_is_day(rad, dt):
    """isDay / isDaySmooth with a one-hour linear / sigmoid transition around sunrise and sunset
    (utils.py:165-206).  Sequential: later transitions overwrite earlier ones exactly as in the reference."""
    is_day = (rad > 0) * 1.0
    smooth = is_day.copy()
    tsz = int(3600 / dt)
    trans = np.linspace(0, 1, tsz)
    trans_s = 1 / (1 + np.exp(-10 * (trans - 0.5)))
    half = tsz // 2
    sunset = False
    for k in range(tsz, len(is_day) - tsz):
        cur = is_day[k]
        if cur == 0:
            sunset = False
            if is_day[k + 1] == 1:
                is_day[k - half:k + half] = trans
                smooth[k - half:k + half] = trans_s
        elif cur == 1 and is_day[k + 1] == 0 and not sunset:
            is_day[k - half:k + half] = 1 - trans
            smooth[k - half:k + half] = 1 - trans_s
            sunset = True
    return is_day, smooth


def load_weather_data(weatherDataDir, location, source, growthYear, startDay, nDays, predHorizon, h, nd=10):
    from scipy.interpolate import PchipInterpolator

    if weatherDataDir is None:
        weatherDataDir = DEFAULT_WEATHER_DIR
    raw = _read_raw(join(weatherDataDir, location, f"{source}{growthYear}.csv"))
    t_all = raw["time"]
    dt_raw = np.mean(np.diff(t_all - t_all[0]))
    n0 = int(np.ceil(startDay * _SECS_DAY / dt_raw))
    n_season = int(np.ceil(nDays * _SECS_DAY / dt_raw))
    n_pred = int(np.ceil(predHorizon * _SECS_DAY / dt_raw)) + 1
    if n0 + n_season + n_pred > len(t_all):
        # append next year's file, shifted to continue the time axis (utils.py:127-154)
        nxt = _read_raw(join(weatherDataDir, location, f"{source}{growthYear + 1}.csv"))
        raw = {c: np.concatenate([raw[c], nxt[c] + (t_all[-1] + dt_raw if c == "time" else 0.0)]) for c in _COLS}
    sl = slice(n0, n0 + n_season + n_pred)
    t = raw["time"][sl]
    if len(t) < n_season + n_pred:
        raise ValueError(f"not enough weather rows for start day {startDay} ({len(t)} < {n_season + n_pred})")
    W = np.zeros((n_season + n_pred, nd))
    W[:, 0] = raw["global radiation"][sl]
    W[:, 1] = raw["air temperature"][sl]
    W[:, 2] = vapor_dens_to_pres(W[:, 1], rh_to_vapor_dens(W[:, 1], raw["RH"][sl]))
    W[:, 3] = co2_ppm_to_dens(W[:, 1], _CO2_PPM_OUT) * 1e6
    W[:, 4] = raw["wind speed"][sl]
    W[:, 5] = raw["sky temperature"][sl]
    W[:, 6] = soil_temp_nl(t)
    W[:, 7] = daily_light_sum(t, W[:, 0])
    W[:, 8], W[:, 9] = compute_is_day(W[:, 0], dt_raw)

    ns = int((dt_raw / h) * (n_season + n_pred))
    table = PchipInterpolator(t, W)(np.linspace(t[0], t[-1], ns))
    table[:, 0][table[:, 0] < 1e-10] = 0
    return table


def init_state(d0, rhMax=90, time_in_days=0):
    """Initial 28-state vector from the first weather row (utils.py:13-46)."""
    x = np.full(28, 16.5)
    x[0] = x[1] = d0[3]
    x[4] = x[21] = 16.5 + 4
    x[11] = 0.25 * (3.0 * 16.5 + d0[6])
    x[12] = 0.25 * (2.0 * 16.5 + 2 * d0[6])
    x[13] = 0.25 * (16.5 + 3 * d0[6])
    x[14] = d0[6]
    x[15] = x[16] = rhMax / 100.0 * sat_vp(16.5)
    x[22] = 0.0
    x[23], x[24], x[25], x[26] = 9.5283e4, 2.5107e5, 5.5338e4, 3.0978e3
    x[27] = time_in_days
    return x
