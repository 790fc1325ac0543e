"""Host-side mirror of groth16/bn128/msm.nim and the generator `**` of groth16/bn128/curves.nim.
All compute happens in libg16b200.so on the GPU."""
from __future__ import annotations

import numpy as np

from . import _lib
from .encoding import FORM_MONT, FORM_STD


def _c(a: np.ndarray, cols: int) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim != 2 or a.shape[1] != cols:
        raise _lib.G16Error("expected an (n, %d) uint64 array" % cols)
    return a


def msm_multi_threaded_g1(nthreads_hint: int, coeffs: np.ndarray, points: np.ndarray,
                          form: int = FORM_MONT) -> np.ndarray:
    """msmMultiThreadedG1 (msm.nim:89-124).  coeffs: (N,4) Fr limbs (Montgomery by default, as the
    reference's seq[Fr]); points: (N,8) affine Montgomery.  Returns the affine sum, (8,) limbs.
    `nthreads_hint` is accepted for signature parity and ignored (no host threads are involved)."""
    coeffs, points = _c(coeffs, 4), _c(points, 8)
    if coeffs.shape[0] != points.shape[0]:
        raise _lib.G16Error("incompatible sequence lengths")            # msm.nim:97
    out = np.zeros(8, dtype=np.uint64)
    lib = _lib.load()
    _lib.check(lib.g16_msm_g1(coeffs.ctypes.data, form, points.ctypes.data, coeffs.shape[0], out.ctypes.data))
    return out


def msm_multi_threaded_g2(nthreads_hint: int, coeffs: np.ndarray, points: np.ndarray,
                          form: int = FORM_MONT) -> np.ndarray:
    """msmMultiThreadedG2 (msm.nim:128-158).  points: (N,16)."""
    coeffs, points = _c(coeffs, 4), _c(points, 16)
    if coeffs.shape[0] != points.shape[0]:
        raise _lib.G16Error("incompatible sequence lengths")
    out = np.zeros(16, dtype=np.uint64)
    lib = _lib.load()
    _lib.check(lib.g16_msm_g2(coeffs.ctypes.data, form, points.ctypes.data, coeffs.shape[0], out.ctypes.data))
    return out


def msm_g1(coeffs, points, form: int = FORM_MONT):      # msm.nim:202
    return msm_multi_threaded_g1(0, coeffs, points, form)


def msm_g2(coeffs, points, form: int = FORM_MONT):      # msm.nim:203
    return msm_multi_threaded_g2(0, coeffs, points, form)


def fixed_base_g1(scalars_std: np.ndarray) -> np.ndarray:
    """[k ** gen1 for k in scalars] (curves.nim:182-188 on the generator of curves.nim:123)."""
    s = _c(scalars_std, 4)
    out = np.zeros((s.shape[0], 8), dtype=np.uint64)
    _lib.check(_lib.load().g16_fixed_base_g1(s.ctypes.data, s.shape[0], out.ctypes.data))
    return out


def fixed_base_g2(scalars_std: np.ndarray) -> np.ndarray:
    s = _c(scalars_std, 4)
    out = np.zeros((s.shape[0], 16), dtype=np.uint64)
    _lib.check(_lib.load().g16_fixed_base_g2(s.ctypes.data, s.shape[0], out.ctypes.data))
    return out
