"""NTT and quotient timing over sizes: python tools/ntt_probe2.py [LOG ...]   (A/B: G16_NTT_RADIX2=1)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nim-groth16_b200"))
import numpy as np, torch
import g16b200 as g
from g16b200 import _lib, encoding as E
lib = _lib.load()
ops, ms = C.c_double(), C.c_float()
_lib.check(lib.g16_bench_int_pipe(3, C.byref(ops), C.byref(ms)))
peak = ops.value
for lg in [int(a) for a in sys.argv[1:]] or [16, 20, 22, 24]:
    n = 1 << lg
    x = torch.from_numpy(E.random_fr_std(n, 6).view(np.int64)).to("cuda")
    y = torch.empty_like(x)
    abc = torch.empty((3 * n, 4), dtype=torch.int64, device="cuda")
    _lib.check(lib.g16_ntt_prepare(lg))
    st = torch.cuda.Stream(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    res = {}
    for name in ("fwd", "inv", "quotient"):
        ts = []
        for i in range(6):
            if name == "quotient":
                abc[:n].copy_(x); abc[n:2 * n].copy_(x)
            torch.cuda.synchronize()
            e0.record(st)
            if name == "quotient":
                _lib.check(lib.g16_quotient_dev(abc.data_ptr(), y.data_ptr(), lg, 1, st.cuda_stream))
            else:
                _lib.check(lib.g16_ntt_fr_dev(x.data_ptr(), y.data_ptr(), x.data_ptr(), lg, int(name == "inv"), st.cuda_stream))
            e1.record(st); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[name] = min(ts[2:])
    mm = n / 2 * lg
    print("2^%d: fwd %.4f ms (%.2f Gelem/s, %.0f %% of modmul peak)  inv %.4f ms  quotient %.4f ms (%.0f %%)" %
          (lg, res["fwd"], n / res["fwd"] / 1e6, 100 * mm / (res["fwd"] * 1e-3) / peak, res["inv"], res["quotient"],
           100 * (6 * mm + 5 * n) / (res["quotient"] * 1e-3) / peak), flush=True)
