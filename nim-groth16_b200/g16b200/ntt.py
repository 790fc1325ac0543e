"""Host-side mirror of groth16/math/domain.nim and groth16/math/ntt.nim (compute on the GPU)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _lib


@dataclass
class Domain:                      # domain.nim:16-21 (the field constants live on the device)
    domainSize: int
    logDomainSize: int


def create_domain(size: int) -> Domain:
    """createDomain (domain.nim:28-46)."""
    log2 = max(0, (size - 1).bit_length())
    if size < 1 or (1 << log2) != size:
        raise _lib.G16Error("domain must have a power-of-two size")     # domain.nim:30
    return Domain(size, log2)


def _run(src: np.ndarray, D: Domain, inverse: int) -> np.ndarray:
    src = np.ascontiguousarray(src, dtype=np.uint64)
    if src.ndim != 2 or src.shape[1] != 4 or src.shape[0] != D.domainSize:
        raise _lib.G16Error("input must have the same size as the domain")   # ntt.nim:57
    out = np.empty_like(src)
    _lib.check(_lib.load().g16_ntt_fr(src.ctypes.data, out.ctypes.data, D.logDomainSize, inverse))
    return out


def forward_ntt(src: np.ndarray, D: Domain) -> np.ndarray:
    """forwardNTT (ntt.nim:55-77): (n,4) Montgomery limbs, natural order in and out."""
    return _run(src, D, 0)


def inverse_ntt(src: np.ndarray, D: Domain) -> np.ndarray:
    """inverseNTT (ntt.nim:139-161), including the 1/n factor."""
    return _run(src, D, 1)
