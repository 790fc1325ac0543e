"""FP64-pipe co-issue probe: DFMA Montgomery multiplies (field_fp64.cuh) next to IMAD Montgomery multiplies."""
import ctypes as C, json, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "nim-groth16_b200"))
from g16b200 import _lib
lib = _lib.load()
rc = lib.g16_selftest(7, 4096)
print("selftest mismatches:", rc, lib.g16_last_error().decode() if rc else "")
out = {}
for kind, name in [(3, "imad_all_warps"), (4, "dfma_odd_warps"), (5, "imad_even_warps"), (6, "mix"), (7, "dfma_all_warps"), (8, "mix_6imad_2dfma"), (9, "mix_2imad_6dfma")]:
    ops, ms = C.c_double(), C.c_float()
    _lib.check(lib.g16_bench_int_pipe(kind, C.byref(ops), C.byref(ms)))
    out[name] = {"modmul_per_s": ops.value, "ms": ms.value}
out["mix"]["modmul_per_s"] *= 2
print(json.dumps(out, indent=1))
