"""Standalone MSM timing over the resident table: python tools/msm_probe.py LOG [g2=1]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nim-groth16_b200"))
import numpy as np, torch
import g16b200 as g
from g16b200 import _lib, encoding as E
log_n = int(sys.argv[1]); g2 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lib = _lib.load()
n = 1 << log_n
dl = E.random_fr_std(n, seed=5)
pts = g.fixed_base_g2(dl) if g2 else g.fixed_base_g1(dl)
sc = E.random_fr_std(n, seed=4)
if os.environ.get("PROBE_SKEWED"):       # 40 % zeros, 20 % ones, 20 % below 2^16, 20 % uniform
    rng = np.random.Generator(np.random.PCG64(8))
    cls = rng.integers(0, 5, size=n)
    sc[cls <= 1] = 0
    sc[cls == 2] = np.array([1, 0, 0, 0], np.uint64)
    small = cls == 3
    sc[small, 1:] = 0
    sc[small, 0] &= np.uint64(0xFFFF)
d_pts = torch.from_numpy(pts.view(np.int64).copy()).to("cuda")
d_sc = torch.from_numpy(sc.view(np.int64).copy()).to("cuda")
res = torch.zeros(64, dtype=torch.int64, device="cuda")
plan = C.c_void_p()
_lib.check(lib.g16_msm_plan_create(g2, n, 0, C.byref(plan)))
_lib.check(lib.g16_msm_plan_profile(plan, 1))
_lib.check(lib.g16_msm_plan_build_table(plan, d_pts.data_ptr(), n, None))
a, t, p = C.c_float(), C.c_float(), C.c_uint64()
for i in range(4):
    _lib.check(lib.g16_msm_dev_table(plan, d_sc.data_ptr(), 1, n, res.data_ptr(), None))
    _lib.check(lib.g16_msm_plan_last_profile(plan, C.byref(a), C.byref(t), C.byref(p)))
print("G%d 2^%d MINB=%s: total %.3f ms accumulate %.3f ms pairs %d" % (g2 + 1, log_n, os.environ.get("G16_G2_MINB", "default"), t.value, a.value, p.value), flush=True)
