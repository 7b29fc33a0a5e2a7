"""GreenLight parameter table (208 entries) for the batched env.

Mirror of `init_default_params` (reference: gl_gym/environments/parameters.py:4-261): same indices, same
values, returned as a float32 array like the reference does (parameters.py:5).  Written as a data table
plus the 22 derived entries (parameters.py:108-126,144,169,171).

NumPy promotion: the reference pins numpy 1.26.4 (requirements.txt:21) where `python_float * np.float32`
is evaluated in float64; numpy >= 2 evaluates it in float32.  Only p[169] (aGroPipe) and p[171] (capGroPipe)
change.  `legacy_promotion=True` (default) reproduces the pinned numpy-1.26 table;
`legacy_promotion=False` reproduces what the reference produces under numpy 2.x (used by the golden check,
tests/golden/params_numpy2.npy).
"""
import math

import numpy as np

NUM_PARAMS = 208
NOISE_LO, NOISE_HI = 128, 162  # noise.py:16  indices perturbed by parametric_crop_uncertainty

PARAM_NAMES = (
    "alfaLeafAir", "L", "sigma", "epsCan", "epsSky", "etaGlobNir",  # 0
    "etaGlobPar", "etaMgPpm", "etaRoofThr", "rhoAir0", "rhoCanPar", "rhoCanNir",  # 6
    "rhoSteel", "rhoWater", "gamma", "omega", "capLeaf", "cEvap1",  # 12
    "cEvap2", "cEvap3Day", "cEvap3Night", "cEvap4Day", "cEvap4Night", "cPAir",  # 18
    "cPSteel", "cPWater", "g", "hSo1", "hSo2", "hSo3",  # 24
    "hSo4", "hSo5", "k1Par", "k2Par", "kNir", "kFir",  # 30
    "mAir", "hSoOut", "mWater", "R", "rCanSp", "rB",  # 36
    "rSMin", "sRs", "etaGlobAir", "psi", "aFlr", "aCov",  # 42
    "hAir", "hGh", "cHecIn", "cHecOut1", "cHecOut2", "cHecOut3",  # 48
    "hElevation", "aRoof", "hVent", "etaInsScr", "aSide", "cDgh",  # 54
    "cLeakage", "cWgh", "hSideRoof", "epsRfFir", "rhoRf", "rhoRfNir",  # 60
    "rhoRfPar", "rhoRfFir", "tauRfNir", "tauRfPar", "tauRfFir", "lambdaRf",  # 66
    "cPRf", "hRf", "epsThScrFir", "rhoThScr", "rhoThScrNir", "rhoThScrPar",  # 72
    "rhoThScrFir", "tauThScrNir", "tauThScrPar", "tauThScrFir", "cPThScr", "hThScr",  # 78
    "kThScr", "epsBlScrFir", "rhoBlScr", "rhoBlScrNir", "rhoBlScrPar", "tauBlScrNir",  # 84
    "tauBlScrPar", "tauBlScrFir", "cPBlScr", "hBlScr", "kBlScr", "epsFlr",  # 90
    "rhoFlr", "rhoFlrNir", "rhoFlrPar", "lambdaFlr", "cPFlr", "hFlr",  # 96
    "rhoCpSo", "lambdaSo", "epsPipe", "phiPipeE", "phiPipeI", "lPipe",  # 102
    "pBoil", "phiExtCo2", "capPipe", "rhoAir", "capAir", "capFlr",  # 108
    "capSo1", "capSo2", "capSo3", "capSo4", "capSo5", "capThScr",  # 114
    "capTop", "capBlScr", "capCo2Air", "capCo2Top", "aPipe", "fCanFlr",  # 120
    "pressure", "energyContentGas", "globJtUmol", "j25LeafMax", "cGamma", "etaCo2AirStom",  # 126
    "eJ", "t25k", "S", "H", "theta", "alpha",  # 132
    "mCh2o", "mCo2", "parJtoUmolSun", "laiMax", "sla", "rgr",  # 138
    "cLeafMax", "cFruitMax", "cFruitG", "cLeafG", "cStemG", "cRgr",  # 144
    "q10m", "cFruitM", "cLeafM", "cStemM", "rgFruit", "rgLeaf",  # 150
    "rgStem", "cBufMax", "cBufMin", "tCan24Max", "tCan24Min", "tCanMax",  # 156
    "tCanMin", "tEndSum", "tEndSumGrowth", "epsGroPipe", "lGroPipe", "phiGroPipeE",  # 162
    "phiGroPipeI", "aGroPipe", "pBoilGro", "capGroPipe", "thetaLampMax", "heatCorrection",  # 168
    "etaLampPar", "etaLampNir", "tauLampPar", "tauLampNir", "tauLampFir", "rhoLampPar",  # 174
    "rhoLampNir", "aLamp", "epsLampTop", "epsLampBottom", "capLamp", "cHecLampAir",  # 180
    "etaLampCool", "zetaLampPar", "intLamps", "vIntLampPos", "fIntLampDown", "capIntLamp",  # 186
    "etaIntLampPar", "etaIntLampNir", "aIntLamp", "epsIntLamp", "thetaIntLampMax", "zetaIntLampPar",  # 192
    "cHecIntLampAir", "tauIntLampFir", "k1IntPar", "k2IntPar", "kIntNir", "kIntFir",  # 198
    "cLeakTop", "minWind", "dmfm", "eps",  # 204
)

# index -> primary (non-derived) value, as written in the GreenLight parameter set
_BASE = {
    0: 5.0, 1: 2450000.0, 2: 5.67e-08, 3: 1.0, 4: 1.0, 5: 0.5,
    6: 0.5, 7: 0.554, 8: 0.9, 9: 1.2, 10: 0.07, 11: 0.35,
    12: 7850.0, 13: 1000.0, 14: 65.8, 15: 1.99e-07, 16: 1200.0, 17: 4.3,
    18: 0.54, 19: 6.1e-07, 20: 1.1e-11, 21: 4.3e-06, 22: 5.2e-06, 23: 1000.0,
    24: 640.0, 25: 4180.0, 26: 9.81, 27: 0.04, 28: 0.08, 29: 0.16,
    30: 0.32, 31: 0.64, 32: 0.7, 33: 0.7, 34: 0.27, 35: 0.94,
    36: 28.96, 37: 1.28, 38: 18.0, 39: 8314.0, 40: 5.0, 41: 275.0,
    42: 82.0, 43: -1.0, 44: 0.1, 45: 23.0, 46: 144.0, 47: 216.6,
    48: 5.7, 49: 6.2, 50: 3.5, 51: 2.8, 52: 1.2, 53: 1.0,
    54: 0.0, 55: 52.2, 56: 0.87, 57: 1.0, 58: 0.0, 59: 0.35,
    60: 3e-05, 61: 0.02, 62: 0.0, 63: 0.85, 64: 2600.0, 65: 0.13,
    66: 0.13, 67: 0.15, 68: 0.57, 69: 0.57, 70: 0.0, 71: 1.05,
    72: 840.0, 73: 0.004, 74: 0.67, 75: 200.0, 76: 0.35, 77: 0.35,
    78: 0.18, 79: 0.75, 80: 0.75, 81: 0.15, 82: 1800.0, 83: 0.00035,
    84: 0.0005, 85: 0.67, 86: 200.0, 87: 0.35, 88: 0.35, 89: 0.01,
    90: 0.01, 91: 0.7, 92: 1800.0, 93: 0.00035, 94: 0.0005, 95: 1.0,
    96: 2300.0, 97: 0.5, 98: 0.65, 99: 1.7, 100: 880.0, 101: 0.02,
    102: 1730000.0, 103: 0.85, 104: 0.88, 105: 0.051, 106: 0.048749999999999995, 107: 1.3375,
    127: 31.65, 128: 2.3, 129: 210.0, 130: 1.7, 131: 0.67, 132: 37000,
    133: 298.15, 134: 710, 135: 220000, 136: 0.7, 137: 0.385, 138: 0.03,
    139: 0.044, 140: 4.6, 141: 3.0, 142: 2.66e-05, 143: 3e-06, 145: 3000000,
    146: 0.27, 147: 0.28, 148: 0.3, 149: 2850000, 150: 2.0, 151: 1.16e-07,
    152: 3.47e-07, 153: 1.47e-07, 154: 0.328, 155: 0.095, 156: 0.074, 157: 20000.0,
    158: 1000.0, 159: 24.5, 160: 15, 161: 34, 162: 10, 163: 1035,
    164: 1250, 165: 0, 166: 1.655, 167: 0.035, 168: 0.033800000000000004, 170: 0,
    172: 116, 173: 0, 174: 0.31, 175: 0.02, 176: 0.95, 177: 0.95,
    178: 0.95, 179: 0.0, 180: 0.0, 181: 0.05, 182: 0.88, 183: 0.88,
    184: 10.0, 185: 2.3, 186: 0.63, 187: 5.2, 188: 0, 189: 0.5,
    190: 0.5, 191: 10, 192: 0, 193: 0, 194: 0, 195: 0,
    196: 0, 197: 0, 198: 0, 199: 1, 200: 1.4, 201: 1.4,
    202: 0.54, 203: 1.88, 204: 0.9, 205: 0.25, 206: 0.0627, 207: 1e-06,
}

def _derived(p, legacy):
    """Fills the derived entries.  `p` is float32; W() widens a Python scalar the way the active NumPy would."""
    f32 = np.float32
    # numpy 1.26: python float (x) float32 scalar -> float64 ; numpy 2: -> float32
    W = np.float64 if legacy else f32
    pi = W(math.pi)

    def store(i, v):
        p[i] = f32(v)

    store(108, W(130.0) * p[46])                                   # pBoil
    store(109, W(5.0) * p[46])                                     # phiExtCo2
    steel_water = ((p[105] * p[105] - p[106] * p[106]) * p[12] * p[24] + p[106] * p[106] * p[13] * p[25])
    store(110, W(0.25) * pi * p[107] * steel_water)                # capPipe
    expo = np.exp(p[26] * p[36] * p[54] / (p[39] * W(293.15)))
    store(111, p[9] * expo)                                        # rhoAir
    store(112, p[48] * p[111] * p[23])                             # capAir
    store(113, p[101] * p[96] * p[100])                            # capFlr
    for k, layer in enumerate((27, 28, 29, 30, 31)):               # capSo1..5
        store(114 + k, p[layer] * p[102])
    store(119, p[83] * p[75] * p[82])                              # capThScr
    store(120, (p[49] - p[48]) * p[111] * p[23])                   # capTop
    store(121, p[93] * p[86] * p[92])                              # capBlScr
    store(122, p[48])                                              # capCo2Air
    store(123, p[49] - p[48])                                      # capCo2Top
    store(124, pi * p[107] * p[105])                               # aPipe
    store(125, 1 - W(0.49) * pi * p[107] * p[105])                 # fCanFlr
    store(126, 101325 * pow((1 - W(2.5577e-5) * p[54]), 5.25588))  # pressure
    store(144, p[141] / p[142])                                    # cLeafMax
    store(169, pi * p[166] * p[167])                               # aGroPipe
    gro = ((p[167] * p[167] - p[168] * p[168]) * p[12] * p[24] + p[168] * p[168] * p[13] * p[25])
    store(171, W(0.25) * pi * p[166] * gro)                        # capGroPipe


def init_default_params(nparams=NUM_PARAMS, legacy_promotion=True):
    """Same contract as the reference's init_default_params(nparams): float32 array of length nparams."""
    if nparams != NUM_PARAMS:
        raise ValueError(f"GreenLight has {NUM_PARAMS} parameters, got nparams={nparams}")
    p = np.zeros(nparams, dtype=np.float32)
    for i, v in _BASE.items():
        p[i] = v
    _derived(p, legacy_promotion)
    return p


def param_index(name):
    return PARAM_NAMES.index(name)
