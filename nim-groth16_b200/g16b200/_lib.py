"""ctypes binding of libg16b200.so (include/g16b200.h).  There is no CPU fallback: importing works
without a GPU (so the symbol table can be checked), every compute call fails loudly without one."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("G16B200_LIB") or os.path.join(os.path.dirname(_HERE), "libg16b200.so")

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)


class G16Error(AssertionError):
    """Raised for a non-zero status; mirrors the reference's AssertionDefect (e.g. msm.nim:97)."""


class ZkeyView(C.Structure):
    _fields_ = [("nvars", C.c_uint32), ("npubs", C.c_uint32), ("log_domain", C.c_uint32), ("flavour", C.c_uint32),
                ("coeff_format", C.c_uint32), ("mem_kind", C.c_uint32), ("flags", C.c_uint32),
                ("reserved0", C.c_uint32), ("ncoeffs", C.c_uint64),
                ("coeffs", C.c_void_p), ("points_a1", C.c_void_p), ("points_b1", C.c_void_p),
                ("points_b2", C.c_void_p), ("points_c1", C.c_void_p), ("points_h1", C.c_void_p),
                ("alpha1", C.c_uint64 * 8), ("beta1", C.c_uint64 * 8), ("beta2", C.c_uint64 * 16),
                ("delta1", C.c_uint64 * 8), ("delta2", C.c_uint64 * 16)]


class ProofRaw(C.Structure):
    _fields_ = [("pi_a", C.c_uint64 * 8), ("pi_b", C.c_uint64 * 16), ("pi_c", C.c_uint64 * 8)]


class Stats(C.Structure):
    _fields_ = [("ms_h2d", C.c_float), ("ms_abc", C.c_float), ("ms_quotient", C.c_float),
                ("ms_sort_witness", C.c_float), ("ms_msm_g1_witness", C.c_float), ("ms_msm_b2", C.c_float),
                ("ms_msm_h", C.c_float), ("reserved_ms", C.c_float), ("ms_assemble", C.c_float),
                ("ms_total", C.c_float), ("kernel_launches", C.c_uint32), ("reserved", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("reserved")}


class R1csView(C.Structure):
    _fields_ = [("nvars", C.c_uint32), ("npubs", C.c_uint32), ("neqs", C.c_uint32), ("flavour", C.c_uint32),
                ("nnz", C.c_uint64 * 3), ("rows", C.c_void_p * 3), ("cols", C.c_void_p * 3),
                ("vals", C.c_void_p * 3)]


class Toxic(C.Structure):
    _fields_ = [("alpha", C.c_uint64 * 4), ("beta", C.c_uint64 * 4), ("gamma", C.c_uint64 * 4),
                ("delta", C.c_uint64 * 4), ("tau", C.c_uint64 * 4)]


class SetupOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("points_a1", "points_b1", "points_b2", "points_c1", "points_h1",
                                          "points_ic", "spec", "dlog_a", "dlog_b", "dlog_k", "dlog_h", "dlog_ic")]


PARTIALS_BYTES = 4 * 64 + 128 + 16      # sizeof(g16_partials): five affine sums and the mask tag
ZKEY_TRUSTED, ZKEY_ONE_SHOT = 1, 2       # g16_zkey_view.flags

# every symbol include/g16b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "g16_last_error": (C.c_char_p, []),
    "g16_version": (C.c_int, []),
    "g16_set_device": (C.c_int, [C.c_int]),
    "g16_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "g16_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "g16_host_unregister": (C.c_int, [C.c_void_p]),
    "g16_release_cached_memory": (C.c_int, []),
    "g16_msm_g1": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "g16_msm_g2": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "g16_ntt_fr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "g16_quotient": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "g16_build_abc": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int, C.c_size_t, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "g16_ctx_create": (C.c_int, [C.POINTER(ZkeyView), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "g16_ctx_clone": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "g16_ctx_destroy": (None, [C.c_void_p]),
    "g16_prove": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(ProofRaw),
                            C.POINTER(Stats)]),
    "g16_prove_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(ProofRaw),
                                C.POINTER(Stats)]),
    "g16_prove_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "g16_prove_wait": (C.c_int, [C.c_void_p, C.POINTER(ProofRaw), C.POINTER(Stats)]),
    "g16_prove_partials_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "g16_prove_partials_wait": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "g16_prove_finish_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "g16_shard_plan": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64)]),
    "g16_ctx_order_stream": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "g16_ctx_last_witness_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "g16_ctx_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_uint64)]),
    "g16_ctx_set_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "g16_prove_partials": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(Stats)]),
    "g16_ctx_last_partials": (C.c_int, [C.c_void_p, C.c_void_p]),
    "g16_prove_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(ProofRaw)]),
    "g16_msm_plan_create": (C.c_int, [C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "g16_msm_plan_destroy": (None, [C.c_void_p]),
    "g16_msm_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "g16_msm_plan_build_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "g16_msm_dev_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]),
    "g16_msm_result_to_affine": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "g16_msm_plan_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "g16_msm_plan_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "g16_msm_plan_last_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                            C.POINTER(C.c_uint64)]),
    "g16_ctx_timer_start": (C.c_int, [C.c_void_p]),
    "g16_ctx_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "g16_ntt_fr_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "g16_quotient_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "g16_ntt_prepare": (C.c_int, [C.c_int]),
    "g16_fake_setup": (C.c_int, [C.POINTER(R1csView), C.POINTER(Toxic), C.POINTER(C.c_uint32),
                                 C.POINTER(SetupOut)]),
    "g16_fixed_base_g1": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "g16_fixed_base_g2": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "g16_selftest": (C.c_int, [C.c_uint32, C.c_uint32]),
    "g16_bench_int_pipe": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]),
    "g16_glv_decompose": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "g16_kernel_launch_count": (C.c_uint64, []),
}

_lib = None


def load():
    """Loads the CUDA library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libg16b200.so is missing at %s: run __graft_entry__.build() or `make -C nim-groth16_b200`; "
                           "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        raise G16Error("g16b200 error %d: %s" % (status, load().g16_last_error().decode("utf-8", "replace")))
