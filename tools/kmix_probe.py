import ctypes as C, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "nim-groth16_b200"))
from g16b200 import _lib
lib = _lib.load()
for kind in (3, 13, 10, 11, 12):
    ops, ms = C.c_double(), C.c_float()
    _lib.check(lib.g16_bench_int_pipe(kind, C.byref(ops), C.byref(ms)))
    print(kind, "%.4g/s" % ops.value, "%.3f ms" % ms.value)
