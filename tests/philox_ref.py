"""numpy restatement of csrc/glg_philox.h (Philox4x32-10 + the 34-draw noise layout) for the GPU parity tests."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & MASK, p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def u01(a, b):
    return ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0


def noise34(seed, env_id, step_ctr, scale):
    k0, k1 = seed & MASK, (seed >> 32) & MASK
    e0, e1 = env_id & MASK, (env_id >> 32) & MASK
    out = np.zeros(34)
    for b in range(17):
        r = philox4x32_10(b, step_ctr & MASK, e0, e1, k0, k1)
        out[2 * b] = -0.5 * scale + scale * u01(r[0], r[1])
        out[2 * b + 1] = -0.5 * scale + scale * u01(r[2], r[3])
    return out


def rand_below(seed, env_id, step_ctr, n):
    r = philox4x32_10(64, step_ctr & MASK, env_id & MASK, (env_id >> 32) & MASK, seed & MASK, (seed >> 32) & MASK)
    return (r[0] * n) >> 32
