"""ctypes wrapper of oracle/libg16oracle.so (g16_oracle_cpu.cpp) -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.
Numpy limb arrays in the boundary layout: scalars standard form (n,4) uint64, points Montgomery."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OraZkey(C.Structure):
    _fields_ = [("nvars", C.c_uint32), ("npubs", C.c_uint32), ("log_n", C.c_uint32), ("flavour", C.c_uint32),
                ("ncoeffs", C.c_uint64), ("coeffs", C.c_void_p), ("a1", C.c_void_p), ("b1", C.c_void_p),
                ("b2", C.c_void_p), ("c1", C.c_void_p), ("h1", C.c_void_p), ("alpha1", C.c_void_p),
                ("beta1", C.c_void_p), ("beta2", C.c_void_p), ("delta1", C.c_void_p), ("delta2", C.c_void_p)]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "libg16oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", HERE])
        _LIB = C.CDLL(path)
        _LIB.ora_prove.restype = C.c_int
        _LIB.ora_build_abc.restype = C.c_int
    return _LIB


def _a(x, cols):
    x = np.ascontiguousarray(x, dtype=np.uint64)
    return x.reshape(-1, cols) if cols else x


def ncpu():
    return os.cpu_count() or 1


def msm_g1(scalars_std, points, nthreads=0):
    s, p = _a(scalars_std, 4), _a(points, 8)
    out = np.zeros(8, np.uint64)
    lib().ora_msm_g1(C.c_void_p(s.ctypes.data), C.c_void_p(p.ctypes.data), C.c_size_t(s.shape[0]), nthreads, ncpu(),
                     C.c_void_p(out.ctypes.data))
    return out


def msm_g2(scalars_std, points, nthreads=0):
    s, p = _a(scalars_std, 4), _a(points, 16)
    out = np.zeros(16, np.uint64)
    lib().ora_msm_g2(C.c_void_p(s.ctypes.data), C.c_void_p(p.ctypes.data), C.c_size_t(s.shape[0]), nthreads, ncpu(),
                     C.c_void_p(out.ctypes.data))
    return out


def ntt(x_mont, inverse=False):
    x = _a(x_mont, 4)
    out = np.empty_like(x)
    lib().ora_ntt(C.c_void_p(x.ctypes.data), C.c_void_p(out.ctypes.data), x.shape[0].bit_length() - 1, int(inverse))
    return out


def build_abc(coeffs44, witness_std, log_n):
    co = np.ascontiguousarray(coeffs44)
    w = _a(witness_std, 4)
    n = 1 << log_n
    az, bz, cz = (np.zeros((n, 4), np.uint64) for _ in range(3))
    rc = lib().ora_build_abc(C.c_void_p(co.ctypes.data), C.c_size_t(co.shape[0]), C.c_void_p(w.ctypes.data), log_n,
                             C.c_void_p(az.ctypes.data), C.c_void_p(bz.ctypes.data), C.c_void_p(cz.ctypes.data))
    if rc:
        raise AssertionError("fatal error")
    return az, bz, cz


def quotient(az, bz, cz, flavour, nthreads=3):
    az, bz, cz = _a(az, 4), _a(bz, 4), _a(cz, 4)
    qs = np.empty_like(az)
    lib().ora_quotient(C.c_void_p(az.ctypes.data), C.c_void_p(bz.ctypes.data), C.c_void_p(cz.ctypes.data),
                       az.shape[0].bit_length() - 1, flavour, nthreads, C.c_void_p(qs.ctypes.data))
    return qs


def prove(zk, witness_std, r: int, s: int, nthreads=0):
    """zk: any object with the g16b200.zkey_types.ZKey fields.  Returns (pi_a, pi_b, pi_c, phase_seconds)."""
    keep = []

    def ptr(a):
        a = np.ascontiguousarray(a)
        keep.append(a)
        return a.ctypes.data

    z = OraZkey()
    z.nvars, z.npubs, z.log_n, z.flavour = zk.nvars, zk.npubs, zk.logDomainSize, zk.flavour
    z.ncoeffs = zk.coeffs.shape[0]
    z.coeffs = ptr(zk.coeffs)
    z.a1, z.b1, z.b2, z.c1, z.h1 = ptr(zk.pointsA1), ptr(zk.pointsB1), ptr(zk.pointsB2), ptr(zk.pointsC1), ptr(zk.pointsH1)
    z.alpha1, z.beta1, z.beta2 = ptr(zk.alpha1), ptr(zk.beta1), ptr(zk.beta2)
    z.delta1, z.delta2 = ptr(zk.delta1), ptr(zk.delta2)
    w = _a(witness_std, 4)
    lim = lambda x: np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
    rr, ss = lim(r), lim(s)
    proof = np.zeros(32, np.uint64)
    phases = (C.c_double * 6)()
    nt = nthreads if nthreads > 0 else ncpu()
    rc = lib().ora_prove(C.byref(z), C.c_void_p(w.ctypes.data), C.c_void_p(rr.ctypes.data),
                         C.c_void_p(ss.ctypes.data), nt, ncpu(), C.c_void_p(proof.ctypes.data), phases)
    if rc:
        raise AssertionError("fatal error")
    return proof[0:8].copy(), proof[8:24].copy(), proof[24:32].copy(), list(phases)
