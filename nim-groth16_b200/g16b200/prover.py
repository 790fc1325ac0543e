"""Host-side mirror of groth16/prover.nim: same entry points and argument meaning, compute on the GPU.

`ProverContext` keeps a zkey resident in HBM (the coarse boundary of include/g16b200.h);
`generate_proof_with_mask` & co. are the drop-in procs (prover.nim:215, 308, 312)."""
from __future__ import annotations

import ctypes as C
import secrets
from typing import Optional

import numpy as np

from . import _lib
from .encoding import COEFF_DTYPE, FORM_MONT, FORM_STD, R
from .zkey_types import JENS_GROTH, SNARKJS, Mask, Proof, Witness, ZKey

MEM_HOST, MEM_DEVICE = 0, 1
COEFF_PACKED44_R2, COEFF_STRUCT48_MONT = 0, 1


def _limbs4(x: int):
    x %= R
    return (C.c_uint64 * 4)(*[(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)])


def build_abc(zkey: ZKey, witness: np.ndarray, witness_form: int = FORM_STD):
    """buildABC (prover.nim:56-73) -> (Az, Bz, Cz), each (n,4) Montgomery limbs."""
    n = zkey.domainSize
    w = np.ascontiguousarray(witness, dtype=np.uint64).reshape(-1, 4)
    co = np.ascontiguousarray(zkey.coeffs, dtype=COEFF_DTYPE)
    az, bz, cz = (np.zeros((n, 4), dtype=np.uint64) for _ in range(3))
    _lib.check(_lib.load().g16_build_abc(co.ctypes.data if co.size else None, co.shape[0], COEFF_PACKED44_R2,
                                         w.ctypes.data, witness_form, w.shape[0], zkey.logDomainSize,
                                         az.ctypes.data, bz.ctypes.data, cz.ctypes.data))
    return az, bz, cz


def _quotient(az: np.ndarray, bz: np.ndarray, flavour: int) -> np.ndarray:
    az = np.ascontiguousarray(az, dtype=np.uint64).reshape(-1, 4)
    bz = np.ascontiguousarray(bz, dtype=np.uint64).reshape(-1, 4)
    n = az.shape[0]
    if bz.shape[0] != n or n < 2 or n & (n - 1):
        raise _lib.G16Error("incompatible vector lengths / domain must be a power of two >= 2")
    qs = np.zeros((n, 4), dtype=np.uint64)
    _lib.check(_lib.load().g16_quotient(az.ctypes.data, bz.ctypes.data, n.bit_length() - 1, flavour, qs.ctypes.data))
    return qs


def compute_snarkjs_scalar_coeffs(nthreads: int, az: np.ndarray, bz: np.ndarray) -> np.ndarray:
    """computeSnarkjsScalarCoeffs (prover.nim:158-181); Cz = Az o Bz is derived as in prover.nim:69-71."""
    return _quotient(az, bz, SNARKJS)


def compute_quotient_pointwise(nthreads: int, az: np.ndarray, bz: np.ndarray) -> np.ndarray:
    """computeQuotientPointwise (prover.nim:118-148): coefficients of Q = (A*B - C)/Z."""
    return _quotient(az, bz, JENS_GROTH)


class ProverContext:
    """A zkey resident on one GPU (g16_ctx).  shard_index/shard_count select the point range of every MSM
    this GPU owns (msm.nim:107-115 chunking across devices)."""

    def __init__(self, zkey: ZKey, shard_index: int = 0, shard_count: int = 1, *, trusted: bool = False,
                 one_shot: bool = False, devices: int = 0):
        """`devices` = N > 0: the whole key over N devices of this process (g16_ctx_create with shard_count = -N,
        first device `shard_index`); prove / submit / wait then work as on one GPU.  `trusted`: skip the on-curve
        validation of the points (io.nim:228-236).  `one_shot`: no window tables (cli_main.nim:193-210 usage)."""
        lib = _lib.load()
        self.zkey = zkey
        if devices > 0:
            shard_count = -devices
        self.shard_index, self.shard_count = shard_index, shard_count
        v = _lib.ZkeyView()
        v.nvars, v.npubs, v.log_domain, v.flavour = zkey.nvars, zkey.npubs, zkey.logDomainSize, zkey.flavour
        v.coeff_format, v.mem_kind = COEFF_PACKED44_R2, MEM_HOST
        v.flags = (_lib.ZKEY_TRUSTED if trusted else 0) | (_lib.ZKEY_ONE_SHOT if one_shot else 0)
        self._keep = []

        def ptr(a, cols):
            a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, cols)
            self._keep.append(a)
            return a.ctypes.data if a.size else None

        # the length asserts of prover.nim:270-276
        if not (zkey.pointsA1.shape[0] == zkey.pointsB1.shape[0] == zkey.pointsB2.shape[0] == zkey.nvars):
            raise _lib.G16Error("witness.len != pts.pointsA1/B1/B2.len")
        if zkey.pointsH1.shape[0] != zkey.domainSize:
            raise _lib.G16Error("hdr.domainSize != pts.pointsH1.len")
        if zkey.pointsC1.shape[0] != zkey.nvars - zkey.npubs - 1:
            raise _lib.G16Error("nvars - npubs - 1 != pts.pointsC1.len")
        co = np.ascontiguousarray(zkey.coeffs, dtype=COEFF_DTYPE)
        self._keep.append(co)
        v.ncoeffs = co.shape[0]
        v.coeffs = co.ctypes.data if co.size else None
        v.points_a1, v.points_b1, v.points_b2 = ptr(zkey.pointsA1, 8), ptr(zkey.pointsB1, 8), ptr(zkey.pointsB2, 16)
        v.points_c1, v.points_h1 = ptr(zkey.pointsC1, 8), ptr(zkey.pointsH1, 8)
        for name, cnt in (("alpha1", 8), ("beta1", 8), ("beta2", 16), ("delta1", 8), ("delta2", 16)):
            arr = getattr(v, name)
            src = np.asarray(getattr(zkey, name), dtype=np.uint64).reshape(-1)
            for i in range(cnt):
                arr[i] = int(src[i])
        self._h = C.c_void_p()
        _lib.check(lib.g16_ctx_create(C.byref(v), shard_index, shard_count, C.byref(self._h)))
        self._keep = []          # uploaded; host copies no longer needed
        self.last_stats: Optional[dict] = None

    def clone(self) -> "ProverContext":
        """g16_ctx_clone: another proof slot sharing this context's resident key (for proofs in flight)."""
        other = ProverContext.__new__(ProverContext)
        other.zkey, other.shard_index, other.shard_count = self.zkey, self.shard_index, self.shard_count
        other._keep, other.last_stats = [], None
        other._h = C.c_void_p()
        _lib.check(_lib.load().g16_ctx_clone(self._h, C.byref(other._h)))
        return other

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.load().g16_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _proof(self, raw: _lib.ProofRaw, wit: np.ndarray, form: int) -> Proof:
        npubs = self.zkey.npubs
        pub = np.array(wit[:npubs + 1], dtype=np.uint64, copy=True)               # prover.nim:239-240
        if form == FORM_MONT:
            from .encoding import fr_from_mont, fr_std
            pub = fr_std(fr_from_mont(pub))
        return Proof(publicIO=pub, pi_a=np.frombuffer(bytes(raw.pi_a), dtype="<u8").copy(),
                     pi_b=np.frombuffer(bytes(raw.pi_b), dtype="<u8").copy(),
                     pi_c=np.frombuffer(bytes(raw.pi_c), dtype="<u8").copy())

    def prove(self, witness: np.ndarray, mask: Mask, witness_form: int = FORM_STD) -> Proof:
        """generateProofWithMask (prover.nim:215-304) against the resident key; host witness in."""
        w = np.ascontiguousarray(witness, dtype=np.uint64).reshape(-1, 4)
        if w.shape[0] != self.zkey.nvars:
            raise _lib.G16Error("wrong witness length")                               # prover.nim:236
        raw, st = _lib.ProofRaw(), _lib.Stats()
        _lib.check(_lib.load().g16_prove(self._h, w.ctypes.data, witness_form, _limbs4(mask.r), _limbs4(mask.s),
                                         C.byref(raw), C.byref(st)))
        self.last_stats = st.as_dict()
        return self._proof(raw, w, witness_form)

    def prove_ptr(self, witness_host_ptr: int, mask: Mask, witness_form: int = FORM_STD):
        """Same, from a raw host pointer (e.g. a pinned torch tensor); returns (ProofRaw, stats)."""
        raw, st = _lib.ProofRaw(), _lib.Stats()
        _lib.check(_lib.load().g16_prove(self._h, witness_host_ptr, witness_form, _limbs4(mask.r), _limbs4(mask.s),
                                         C.byref(raw), C.byref(st)))
        self.last_stats = st.as_dict()
        return raw, self.last_stats

    def prove_dev(self, witness_std_dev_ptr: int, mask: Mask):
        """Witness already in device memory (standard form)."""
        raw, st = _lib.ProofRaw(), _lib.Stats()
        _lib.check(_lib.load().g16_prove_dev(self._h, witness_std_dev_ptr, _limbs4(mask.r), _limbs4(mask.s),
                                             C.byref(raw), C.byref(st)))
        self.last_stats = st.as_dict()
        return raw, self.last_stats

    def submit(self, witness_ptr: int, mask: Mask, witness_form: int = FORM_STD, mem_kind: int = MEM_HOST):
        """g16_prove_submit: enqueue a proof and return; the witness buffer must outlive wait()."""
        _lib.check(_lib.load().g16_prove_submit(self._h, witness_ptr, witness_form, mem_kind, _limbs4(mask.r),
                                                _limbs4(mask.s)))

    def wait(self):
        """g16_prove_wait -> (ProofRaw, stats)."""
        raw, st = _lib.ProofRaw(), _lib.Stats()
        _lib.check(_lib.load().g16_prove_wait(self._h, C.byref(raw), C.byref(st)))
        self.last_stats = st.as_dict()
        return raw, self.last_stats

    def set_mask(self, mask: Mask):
        """g16_ctx_set_mask: announce r, s before the partial sums, so that this rank folds s*A_k + r*B1_k into its
        record and prove_finish needs no scalar multiplication.  All ranks of a proof call it, or none."""
        _lib.check(_lib.load().g16_ctx_set_mask(self._h, _limbs4(mask.r), _limbs4(mask.s)))

    def prove_partials(self, witness_ptr: int, witness_form: int, mem_kind: int, partials_dev_ptr: int):
        st = _lib.Stats()
        _lib.check(_lib.load().g16_prove_partials(self._h, witness_ptr, witness_form, mem_kind, partials_dev_ptr,
                                                  C.byref(st)))
        self.last_stats = st.as_dict()
        return self.last_stats

    def layout(self):
        """(has_window_tables, device_bytes): g16_ctx_layout."""
        t, b = C.c_int(), C.c_uint64()
        _lib.check(_lib.load().g16_ctx_layout(self._h, C.byref(t), C.byref(b)))
        return bool(t.value), int(b.value)

    def last_witness_bytes(self) -> int:
        """Bytes of witness the last prove / partials call copied to the device(s) of this context."""
        n = C.c_uint64()
        _lib.check(_lib.load().g16_ctx_last_witness_bytes(self._h, C.byref(n)))
        return int(n.value)

    def prove_finish(self, gathered_dev_ptr: int, count: int, mask: Mask) -> _lib.ProofRaw:
        raw = _lib.ProofRaw()
        _lib.check(_lib.load().g16_prove_finish(self._h, gathered_dev_ptr, count, _limbs4(mask.r), _limbs4(mask.s),
                                                C.byref(raw)))
        return raw


def generate_proof_with_mask(nthreads: int, print_timings: bool, zkey: ZKey, wtns: Witness, mask: Mask,
                             ctx: Optional[ProverContext] = None) -> Proof:
    """generateProofWithMask (prover.nim:215-304).  `nthreads` is kept for signature parity."""
    if zkey.curve != wtns.curve:
        raise _lib.G16Error("zkey.header.curve != wtns.curve")                        # prover.nim:224
    if zkey.nvars != wtns.nvars:
        raise _lib.G16Error("wrong witness length")                                   # prover.nim:236
    own = ctx is None
    ctx = ctx or ProverContext(zkey)
    try:
        proof = ctx.prove(wtns.values, mask, FORM_STD)
        if print_timings:                                                             # prover.nim:244-297 labels
            s = ctx.last_stats
            # the reference's labels (prover.nim:244-297); pi_A, rho and the C half of pi_C are one fused pass here
            for label, ms in (("building 'ABC'", s["ms_abc"]), ("computing the quotient (FFTs)", s["ms_quotient"]),
                              ("computing pi_A + rho + pi_C/C (3x G1 MSM, fused)",
                               s["ms_sort_witness"] + s["ms_msm_g1_witness"]),
                              ("computing pi_B (G2 MSM)", s["ms_msm_b2"]),
                              ("computing pi_C/H (G1 MSM)", s["ms_msm_h"])):
                print("%s took %.4f seconds" % (label, ms / 1e3))
        return proof
    finally:
        if own:
            ctx.close()


def generate_proof_with_trivial_mask(nthreads: int, print_timings: bool, zkey: ZKey, wtns: Witness,
                                     ctx: Optional[ProverContext] = None) -> Proof:
    """generateProofWithTrivialMask (prover.nim:308-310)."""
    return generate_proof_with_mask(nthreads, print_timings, zkey, wtns, Mask(0, 0), ctx)


def generate_proof(nthreads: int, print_timings: bool, zkey: ZKey, wtns: Witness,
                   ctx: Optional[ProverContext] = None) -> Proof:
    """generateProof (prover.nim:312-319) with random masks (here from the OS CSPRNG)."""
    mask = Mask(secrets.randbelow(R), secrets.randbelow(R))
    return generate_proof_with_mask(nthreads, print_timings, zkey, wtns, mask, ctx)
