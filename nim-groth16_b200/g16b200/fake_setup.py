"""Fake circuit-specific trusted setup (groth16/fake_setup.nim) on the GPU: random or explicit toxic
waste -> an in-memory ZKey in the .zkey byte layout.  Fixture generator for tests and benchmarks."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _lib
from .encoding import COEFF_DTYPE, MONT, R, ints_to_limbs
from .zkey_types import R1CS, SNARKJS, ZKey


@dataclass
class ToxicWaste:                 # fake_setup.nim:24-30
    alpha: int
    beta: int
    gamma: int
    delta: int
    tau: int


def random_toxic_waste(seed: Optional[int] = None) -> ToxicWaste:
    """randomToxicWaste (fake_setup.nim:32-42); like rnd.nim this is NOT a cryptographic source."""
    rng = np.random.Generator(np.random.PCG64(seed))
    vals = [int.from_bytes(rng.bytes(40), "little") % R for _ in range(5)]
    return ToxicWaste(*vals)


@dataclass
class SetupScalars:
    """Discrete logs of the zkey points (what makes closed-form checks possible)."""
    a: np.ndarray
    b: np.ndarray
    k: np.ndarray
    h: np.ndarray
    ic: np.ndarray


def r1cs_to_coeffs(r1cs: R1CS) -> np.ndarray:
    """r1csToCoeffs (fake_setup.nim:46-66): A and B entries row by row, then the npub+1 dummy rows
    A[n+i][i] = 1; values R^2-encoded as they sit in a .zkey (io.nim:134-139)."""
    n = r1cs.nConstr
    p = r1cs.nPubIn + r1cs.nPubOut
    na, nb = len(r1cs.rows[0]), len(r1cs.rows[1])
    out = np.zeros(na + nb + p + 1, dtype=COEFF_DTYPE)
    r2 = MONT * MONT % R

    def enc(vals: np.ndarray) -> np.ndarray:
        b = np.ascontiguousarray(vals, dtype="<u8").tobytes()
        ints = (int.from_bytes(b[i:i + 32], "little") * r2 % R for i in range(0, len(b), 32))
        return ints_to_limbs(ints) if len(b) else np.zeros((0, 4), np.uint64)

    m = np.concatenate([np.zeros(na, np.uint32), np.ones(nb, np.uint32)])
    rows = np.concatenate([r1cs.rows[0], r1cs.rows[1]]).astype(np.uint32)
    cols = np.concatenate([r1cs.cols[0], r1cs.cols[1]]).astype(np.uint32)
    vals = np.concatenate([enc(r1cs.vals[0]), enc(r1cs.vals[1])]) if na + nb else np.zeros((0, 4), np.uint64)
    order = np.lexsort((m, rows))                     # row-major, A before B inside a row
    out["m"][:na + nb] = m[order]
    out["row"][:na + nb] = rows[order]
    out["col"][:na + nb] = cols[order]
    out["val"][:na + nb] = vals[order]
    one = ints_to_limbs([r2])[0]
    for i in range(n, n + p + 1):                     # fake_setup.nim:61-63
        j = na + nb + (i - n)
        out["m"][j], out["row"][j], out["col"][j], out["val"][j] = 0, i, i - n, one
    return out


def fake_circuit_setup(r1cs: R1CS, toxic: ToxicWaste, flavour: int = SNARKJS,
                       want_scalars: bool = False) -> Tuple[ZKey, Optional[SetupScalars]]:
    """fakeCircuitSetup (fake_setup.nim:201-326)."""
    lib = _lib.load()
    nvars = r1cs.nWires
    npubs = r1cs.nPubIn + r1cs.nPubOut
    neqs = r1cs.nConstr
    logn = max(1, (neqs + npubs + 1 - 1).bit_length())          # fake_setup.nim:205
    n = 1 << logn
    nk = nvars - npubs - 1
    view = _lib.R1csView()
    view.nvars, view.npubs, view.neqs, view.flavour = nvars, npubs, neqs, flavour
    keep = []
    for m in range(3):
        rr = np.ascontiguousarray(r1cs.rows[m], dtype=np.uint32)
        cc = np.ascontiguousarray(r1cs.cols[m], dtype=np.uint32)
        vv = np.ascontiguousarray(r1cs.vals[m], dtype=np.uint64).reshape(-1, 4)
        keep += [rr, cc, vv]
        view.nnz[m] = rr.shape[0]
        view.rows[m] = rr.ctypes.data if rr.size else None
        view.cols[m] = cc.ctypes.data if cc.size else None
        view.vals[m] = vv.ctypes.data if vv.size else None
    tox = _lib.Toxic()
    for name in ("alpha", "beta", "gamma", "delta", "tau"):
        v = getattr(toxic, name) % R
        arr = getattr(tox, name)
        for i in range(4):
            arr[i] = (v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF
    z = lambda rows, cols: np.zeros((rows, cols), dtype=np.uint64)
    a1, b1, b2, c1, h1, ic = z(nvars, 8), z(nvars, 8), z(nvars, 16), z(nk, 8), z(n, 8), z(npubs + 1, 8)
    spec = np.zeros(3 * 8 + 3 * 16, dtype=np.uint64)
    out = _lib.SetupOut()
    out.points_a1, out.points_b1, out.points_b2 = a1.ctypes.data, b1.ctypes.data, b2.ctypes.data
    out.points_c1 = c1.ctypes.data if nk else None
    out.points_h1, out.points_ic, out.spec = h1.ctypes.data, ic.ctypes.data, spec.ctypes.data
    sc = None
    if want_scalars:
        sc = SetupScalars(z(nvars, 4), z(nvars, 4), z(nk, 4), z(n, 4), z(npubs + 1, 4))
        out.dlog_a, out.dlog_b, out.dlog_h, out.dlog_ic = (sc.a.ctypes.data, sc.b.ctypes.data, sc.h.ctypes.data,
                                                           sc.ic.ctypes.data)
        out.dlog_k = sc.k.ctypes.data if nk else None
    logd = C.c_uint32(0)
    _lib.check(lib.g16_fake_setup(C.byref(view), C.byref(tox), C.byref(logd), C.byref(out)))
    assert logd.value == logn
    zk = ZKey(nvars=nvars, npubs=npubs, domainSize=n, logDomainSize=logn, flavour=flavour,
              alpha1=spec[0:8].copy(), beta1=spec[8:16].copy(), delta1=spec[16:24].copy(),
              beta2=spec[24:40].copy(), gamma2=spec[40:56].copy(), delta2=spec[56:72].copy(),
              pointsIC=ic, pointsA1=a1, pointsB1=b1, pointsB2=b2, pointsC1=c1, pointsH1=h1,
              coeffs=r1cs_to_coeffs(r1cs))
    return zk, sc


def create_fake_circuit_setup(r1cs: R1CS, flavour: int = SNARKJS) -> ZKey:
    """createFakeCircuitSetup (fake_setup.nim:330-332)."""
    return fake_circuit_setup(r1cs, random_toxic_waste(), flavour)[0]


def synthetic_chain_circuit(neqs: int, seed: int = 3) -> Tuple[R1CS, np.ndarray]:
    """The synthetic benchmark circuit of SURVEY.md 8d: nPubOut = 1, nPubIn = 0; wires 0:'1', 1:out,
    2:x0, 3..:x_{j+1}; constraint j < neqs-1: (x_j + c_j) * x_j = x_{j+1}; last: x_last * 1 = out.
    Returns (r1cs, witness as (nvars,4) standard-form limbs).  Every witness value is full width."""
    rng = np.random.Generator(np.random.PCG64(seed))
    raw = rng.bytes(32 * neqs)
    cs = [int.from_bytes(raw[32 * i:32 * i + 32], "little") % R for i in range(neqs)]
    x = cs[-1]
    wit = [1, 0, x]
    for j in range(neqs - 1):
        x = (x + cs[j]) * x % R
        wit.append(x)
    wit[1] = wit[neqs + 1]
    nv = neqs + 2
    j = np.arange(neqs - 1, dtype=np.uint32)
    one = ints_to_limbs([1])
    rows_a = np.concatenate([j, j, np.asarray([neqs - 1], np.uint32)])
    cols_a = np.concatenate([j + 2, np.zeros(neqs - 1, np.uint32), np.asarray([neqs + 1], np.uint32)])
    vals_a = np.concatenate([np.repeat(one, neqs - 1, 0), ints_to_limbs(cs[:neqs - 1]) if neqs > 1 else
                             np.zeros((0, 4), np.uint64), one])
    oa = np.lexsort((cols_a, rows_a))
    rows_b = np.concatenate([j, np.asarray([neqs - 1], np.uint32)])
    cols_b = np.concatenate([j + 2, np.asarray([0], np.uint32)])
    vals_b = np.repeat(one, neqs, 0)
    rows_c = np.concatenate([j, np.asarray([neqs - 1], np.uint32)])
    cols_c = np.concatenate([j + 3, np.asarray([1], np.uint32)])
    vals_c = np.repeat(one, neqs, 0)
    r1cs = R1CS(nWires=nv, nPubOut=1, nPubIn=0, nPrivIn=1, nConstr=neqs,
                rows=(rows_a[oa], rows_b, rows_c), cols=(cols_a[oa], cols_b, cols_c),
                vals=(vals_a[oa], vals_b, vals_c))
    return r1cs, ints_to_limbs(wit)
