"""Readers and writers of the circom/snarkjs binary formats, zero-copy with numpy.

Mirrors groth16/files/container.nim:75-93, zkey.nim:114-248, witness.nim:36-76 and r1cs.nim:84-176.  The
.zkey point sections and the .wtns value section are already in the byte layout the GPU library
consumes, so parsing is `np.frombuffer` on the section, not a per-element loop.
The reference has no writers; the writers here produce files its parsers accept."""
from __future__ import annotations

import struct
from typing import Dict, Tuple

import numpy as np

from . import _lib
from .encoding import COEFF_DTYPE, P, R
from .zkey_types import R1CS, SNARKJS, Witness, ZKey


def parse_container(data, magic: bytes, version: int) -> Dict[int, memoryview]:
    """container.nim:75-93: magic, version, nsections, then (id u32, len u64, payload)*."""
    mv = memoryview(data)
    if bytes(mv[0:4]) != magic:
        raise _lib.G16Error("not a `%s` file" % magic.decode())
    ver, nsec = struct.unpack_from("<II", mv, 4)
    if ver != version:
        raise _lib.G16Error("not a version %d `%s` file" % (version, magic.decode()))
    pos, out = 12, {}
    for _ in range(nsec):
        sid, slen = struct.unpack_from("<IQ", mv, pos)
        pos += 12
        out[sid] = mv[pos:pos + slen]
        pos += slen
    return out


def _container(magic: bytes, version: int, sections) -> bytes:
    parts = [magic, struct.pack("<II", version, len(sections))]
    for sid, payload in sections:
        payload = bytes(payload)
        parts.append(struct.pack("<IQ", sid, len(payload)))
        parts.append(payload)
    return b"".join(parts)


def _u64(mv, cols) -> np.ndarray:
    return np.frombuffer(mv, dtype="<u8").reshape(-1, cols)


def parse_zkey_bytes(data) -> ZKey:
    """parseZKey (zkey.nim:241-246); flavour is Snarkjs as in zkey.nim:129."""
    sec = parse_container(data, b"zkey", 1)
    if struct.unpack("<I", sec[1])[0] != 1:
        raise _lib.G16Error("expecting `.zkey` file for a Groth16 prover")             # zkey.nim:110
    s2 = sec[2]
    n8p = struct.unpack_from("<I", s2, 0)[0]
    if n8p != 32 or int.from_bytes(s2[4:36], "little") != P:
        raise _lib.G16Error("expecting the alt-bn128 curve")                            # zkey.nim:134
    n8r = struct.unpack_from("<I", s2, 36)[0]
    if n8r != 32 or int.from_bytes(s2[40:72], "little") != R:
        raise _lib.G16Error("expecting the alt-bn128 curve")
    if len(s2) != 2 * 4 + 32 + 32 + 3 * 4 + 3 * 64 + 3 * 128:
        raise _lib.G16Error("unexpected section length")                                # zkey.nim:122
    nvars, npubs, dom = struct.unpack_from("<III", s2, 72)
    logd = max(0, (dom - 1).bit_length())
    if (1 << logd) != dom:
        raise _lib.G16Error("domain size should be a power of two")                     # zkey.nim:143
    spec = np.frombuffer(s2, dtype="<u8", offset=84)
    alpha1, beta1 = spec[0:8], spec[8:16]
    beta2, gamma2 = spec[16:32], spec[32:48]
    delta1, delta2 = spec[48:56], spec[56:72]
    ncoeffs = struct.unpack_from("<I", sec[4], 0)[0]
    if len(sec[4]) != 4 + ncoeffs * 44:
        raise _lib.G16Error("unexpected section length")                                # zkey.nim:171
    coeffs = np.frombuffer(sec[4], dtype=COEFF_DTYPE, offset=4, count=ncoeffs)
    for sid, cnt, size in ((3, npubs + 1, 64), (5, nvars, 64), (6, nvars, 64), (7, nvars, 128),
                           (8, nvars - npubs - 1, 64), (9, dom, 64)):
        if len(sec[sid]) != cnt * size:
            raise _lib.G16Error("unexpected section length")                            # zkey.nim:198-224
    return ZKey(nvars=nvars, npubs=npubs, domainSize=dom, logDomainSize=logd, flavour=SNARKJS,
                alpha1=alpha1, beta1=beta1, beta2=beta2, gamma2=gamma2, delta1=delta1, delta2=delta2,
                pointsIC=_u64(sec[3], 8), pointsA1=_u64(sec[5], 8), pointsB1=_u64(sec[6], 8),
                pointsB2=_u64(sec[7], 16), pointsC1=_u64(sec[8], 8), pointsH1=_u64(sec[9], 8), coeffs=coeffs)


def parse_zkey(fname: str) -> ZKey:
    return parse_zkey_bytes(np.fromfile(fname, dtype=np.uint8).data)


def write_zkey_bytes(zk: ZKey) -> bytes:
    """Sections 1..9 of zkey.nim:1-92."""
    le = lambda a: np.ascontiguousarray(a, dtype="<u8").tobytes()
    s2 = (struct.pack("<I", 32) + P.to_bytes(32, "little") + struct.pack("<I", 32) + R.to_bytes(32, "little")
          + struct.pack("<III", zk.nvars, zk.npubs, zk.domainSize)
          + le(zk.alpha1) + le(zk.beta1) + le(zk.beta2) + le(zk.gamma2) + le(zk.delta1) + le(zk.delta2))
    co = np.ascontiguousarray(zk.coeffs, dtype=COEFF_DTYPE)
    s4 = struct.pack("<I", co.shape[0]) + co.tobytes()
    return _container(b"zkey", 1, [(1, struct.pack("<I", 1)), (2, s2), (3, le(zk.pointsIC)), (4, s4),
                                   (5, le(zk.pointsA1)), (6, le(zk.pointsB1)), (7, le(zk.pointsB2)),
                                   (8, le(zk.pointsC1)), (9, le(zk.pointsH1))])


def write_zkey(fname: str, zk: ZKey) -> None:
    with open(fname, "wb") as f:
        f.write(write_zkey_bytes(zk))


def parse_witness_bytes(data) -> Witness:
    """parseWitness (witness.nim:71-76)."""
    sec = parse_container(data, b"wtns", 2)
    s1 = sec[1]
    n8r = struct.unpack_from("<I", s1, 0)[0]
    if n8r != 32:
        raise _lib.G16Error("expecting 256 bit prime")                                  # witness.nim:46
    if int.from_bytes(s1[4:36], "little") != R:
        raise _lib.G16Error("expecting the alt-bn128 curve")                            # witness.nim:47
    if len(s1) != 4 + 32 + 4:
        raise _lib.G16Error("unexpected section length")                                # witness.nim:44
    nvars = struct.unpack_from("<I", s1, 36)[0]
    if len(sec[2]) != 32 * nvars:
        raise _lib.G16Error("unexpected section length")                                # witness.nim:59
    return Witness(values=_u64(sec[2], 4))


def parse_witness(fname: str) -> Witness:
    return parse_witness_bytes(np.fromfile(fname, dtype=np.uint8).data)


def write_witness_bytes(w: Witness) -> bytes:
    s1 = struct.pack("<I", 32) + R.to_bytes(32, "little") + struct.pack("<I", w.nvars)
    return _container(b"wtns", 2, [(1, s1), (2, np.ascontiguousarray(w.values, dtype="<u8").tobytes())])


def write_witness(fname: str, w: Witness) -> None:
    with open(fname, "wb") as f:
        f.write(write_witness_bytes(w))


def parse_r1cs_bytes(data) -> R1CS:
    """parseR1CS (r1cs.nim:170-176); constraints become COO triplets per matrix."""
    sec = parse_container(data, b"r1cs", 1)
    s1 = sec[1]
    n8r = struct.unpack_from("<I", s1, 0)[0]
    if int.from_bytes(s1[4:4 + n8r], "little") != R:
        raise _lib.G16Error("expecting the alt-bn128 curve")                            # r1cs.nim:95
    if len(s1) != 4 + n8r + 16 + 8 + 4:
        raise _lib.G16Error("unexpected section length")                                # r1cs.nim:93
    nWires, nPubOut, nPubIn, nPrivIn, nLabels, nConstr = struct.unpack_from("<IIIIQI", s1, 4 + n8r)
    s2 = sec[2]
    rows = ([], [], [])
    cols = ([], [], [])
    vals = ([], [], [])
    pos = 0
    for i in range(nConstr):
        for m in range(3):
            nterms = struct.unpack_from("<I", s2, pos)[0]
            pos += 4
            for _ in range(nterms):
                cols[m].append(struct.unpack_from("<I", s2, pos)[0])
                vals[m].append(bytes(s2[pos + 4:pos + 36]))
                rows[m].append(i)
                pos += 36
    mk = lambda l: np.asarray(l, dtype=np.uint32)
    mv = lambda l: np.frombuffer(b"".join(l), dtype="<u8").reshape(-1, 4).copy() if l else np.zeros((0, 4), np.uint64)
    labels = np.frombuffer(sec[3], dtype="<u8").copy() if 3 in sec else None
    return R1CS(nWires=nWires, nPubOut=nPubOut, nPubIn=nPubIn, nPrivIn=nPrivIn, nConstr=nConstr,
                rows=tuple(mk(r) for r in rows), cols=tuple(mk(c) for c in cols), vals=tuple(mv(v) for v in vals),
                nLabels=nLabels, wireToLabel=labels)


def parse_r1cs(fname: str) -> R1CS:
    return parse_r1cs_bytes(np.fromfile(fname, dtype=np.uint8).data)


def write_r1cs_bytes(r: R1CS) -> bytes:
    """r1cs.nim:1-50; constraints are emitted row by row from the COO arrays (rows must be sorted)."""
    s1 = (struct.pack("<I", 32) + R.to_bytes(32, "little")
          + struct.pack("<IIIIQI", r.nWires, r.nPubOut, r.nPubIn, r.nPrivIn, r.nLabels, r.nConstr))
    ptr = [np.searchsorted(r.rows[m], np.arange(r.nConstr + 1)) for m in range(3)]
    parts = []
    for i in range(r.nConstr):
        for m in range(3):
            a, b = int(ptr[m][i]), int(ptr[m][i + 1])
            parts.append(struct.pack("<I", b - a))
            for t in range(a, b):
                parts.append(struct.pack("<I", int(r.cols[m][t])) + np.ascontiguousarray(r.vals[m][t], "<u8").tobytes())
    s3 = np.arange(r.nWires, dtype="<u8").tobytes()
    return _container(b"r1cs", 1, [(1, s1), (2, b"".join(parts)), (3, s3)])
