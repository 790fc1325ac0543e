"""Byte layouts of the boundary (groth16/bn128/io.nim:103-153) as numpy arrays.

Fr/Fp element = 4 little-endian uint64 limbs; Fr vectors are (n, 4) uint64 arrays, G1 arrays (n, 8),
G2 arrays (n, 16).  "mont" = Montgomery residue with R = 2^256, "std" = the plain integer."""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np

P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47   # fields.nim:36
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001   # fields.nim:37
MONT = 1 << 256
_RINV_R = pow(MONT, -1, R)
_RINV_P = pow(MONT, -1, P)

FORM_MONT, FORM_STD = 0, 1


def ints_to_limbs(xs: Iterable[int]) -> np.ndarray:
    data = b"".join(int(x).to_bytes(32, "little") for x in xs)
    return np.frombuffer(data, dtype="<u8").reshape(-1, 4).copy()


def limbs_to_ints(a: np.ndarray) -> List[int]:
    b = np.ascontiguousarray(a, dtype="<u8").tobytes()
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def fr_std(xs: Iterable[int]) -> np.ndarray:
    return ints_to_limbs(x % R for x in xs)


def fr_mont(xs: Iterable[int]) -> np.ndarray:
    return ints_to_limbs(x * MONT % R for x in xs)


def fr_from_std(a: np.ndarray) -> List[int]:
    return limbs_to_ints(a)


def fr_from_mont(a: np.ndarray) -> List[int]:
    return [x * _RINV_R % R for x in limbs_to_ints(a)]


def g1_array(points: Sequence[Tuple[int, int]]) -> np.ndarray:
    flat = []
    for (x, y) in points:
        flat += [x * MONT % P, y * MONT % P]
    return ints_to_limbs(flat).reshape(-1, 8)


def g2_array(points) -> np.ndarray:
    flat = []
    for (x, y) in points:
        flat += [x[0] * MONT % P, x[1] * MONT % P, y[0] * MONT % P, y[1] * MONT % P]
    return ints_to_limbs(flat).reshape(-1, 16)


def g1_from_array(a: np.ndarray):
    v = [x * _RINV_P % P for x in limbs_to_ints(np.ascontiguousarray(a).reshape(-1, 4))]
    return [(v[2 * i], v[2 * i + 1]) for i in range(len(v) // 2)]


def g2_from_array(a: np.ndarray):
    v = [x * _RINV_P % P for x in limbs_to_ints(np.ascontiguousarray(a).reshape(-1, 4))]
    return [((v[4 * i], v[4 * i + 1]), (v[4 * i + 2], v[4 * i + 3])) for i in range(len(v) // 4)]


COEFF_DTYPE = np.dtype([("m", "<u4"), ("row", "<u4"), ("col", "<u4"), ("val", "<u8", (4,))])   # zkey.nim:169-188
assert COEFF_DTYPE.itemsize == 44


def random_fr_std(n: int, seed: int) -> np.ndarray:
    """n pseudo-random standard-form scalars below 2^253 (< r), reproducible."""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 61) - 1)
    return a
