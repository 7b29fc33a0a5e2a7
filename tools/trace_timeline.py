"""Per-warp timeline of kernel C's evaluation loop from a -DGLG_TRACE build: python tools/trace_timeline.py <lib.so> [B]
Prints, averaged over 30 evaluations of CTA 0: when each warp woke up and when it arrived at its barrier, in cycles after the
owners' previous arrive (warps 0..3 = owners, 4.. = group warps)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "greenlight-gym2_b200"))
from glgym import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
import numpy as np, torch
from glgym.vec_env import GreenLightVecEnv
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
env = GreenLightVecEnv(B, n_sub=600, integrator="fixed", role_warps=2); env.reset_tensor()
A = torch.rand(B, 6, device="cuda") * 2 - 1
for _ in range(3): env.step_tensor(A)
torch.cuda.synchronize()
buf = np.zeros((32, 16, 2), dtype=np.int64)
L = C.CDLL(_lib.LIB_PATH)
assert L.glg_debug_trace(C.c_void_p(buf.ctypes.data)) == 0
t = buf.astype(np.float64)
# reference time of evaluation e: the last owner arrive of evaluation e-1
ref = t[:-1, 0:4, 1].max(axis=1)  # [31]
wake = t[1:, :, 0] - ref[:, None]
arr = t[1:, :, 1] - ref[:, None]
per = np.diff(ref).mean()
print(f"{os.path.basename(sys.argv[1])}: {per:.0f} cycles per evaluation (CTA 0)")
for w in range(16):
    print(f"  warp {w:2d} {'owner' if w < 4 else 'group %2d' % (w - 4)}: wake {wake[1:, w].mean():7.0f}  arrive {arr[1:, w].mean():7.0f}  busy {(arr[1:, w] - wake[1:, w]).mean():6.0f}")
