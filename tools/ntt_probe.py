"""Forward NTT timing: python tools/ntt_probe.py LOG"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nim-groth16_b200"))
import numpy as np, torch
import g16b200 as g
from g16b200 import _lib, encoding as E
lg = int(sys.argv[1]); n = 1 << lg
lib = _lib.load()
x = torch.from_numpy(E.random_fr_std(n, 6).view(np.int64)).to("cuda")
y = torch.empty_like(x); wk = torch.empty_like(x)
_lib.check(lib.g16_ntt_prepare(lg))
st = torch.cuda.Stream(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
for i in range(4):
    wk.copy_(x); torch.cuda.synchronize()
    e0.record(st)
    _lib.check(lib.g16_ntt_fr_dev(wk.data_ptr(), y.data_ptr(), wk.data_ptr(), lg, 0, st.cuda_stream))
    e1.record(st); torch.cuda.synchronize()
print("ntt 2^%d: %.4f ms" % (lg, e0.elapsed_time(e1)))
