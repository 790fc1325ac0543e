"""Data model mirroring groth16/zkey_types.nim and groth16/files/witness.nim / r1cs.nim, with numpy payloads
in the boundary's byte layout (points and H/A/B/C arrays exactly as they sit in a .zkey)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

JENS_GROTH = 0      # Flavour.JensGroth   zkey_types.nim:11
SNARKJS = 1         # Flavour.Snarkjs     zkey_types.nim:12


@dataclass
class ZKey:                           # zkey_types.nim:54-60
    nvars: int
    npubs: int
    domainSize: int
    logDomainSize: int
    flavour: int
    alpha1: np.ndarray                # (8,)   SpecPoints, zkey_types.nim:24-31
    beta1: np.ndarray
    beta2: np.ndarray                 # (16,)
    gamma2: np.ndarray
    delta1: np.ndarray
    delta2: np.ndarray
    pointsIC: np.ndarray              # (npubs+1, 8)   VerifierPoints
    pointsA1: np.ndarray              # (nvars, 8)     ProverPoints, zkey_types.nim:34-40
    pointsB1: np.ndarray
    pointsB2: np.ndarray              # (nvars, 16)
    pointsC1: np.ndarray              # (nvars-npubs-1, 8)
    pointsH1: np.ndarray              # (domainSize, 8)
    coeffs: np.ndarray                # COEFF_DTYPE records, values R^2-encoded as on disk (zkey.nim:169-188)
    curve: str = "bn128"


@dataclass
class Witness:                        # witness.nim:28-32
    values: np.ndarray                # (nvars, 4) standard form (witness.nim:14)
    curve: str = "bn128"

    @property
    def nvars(self) -> int:
        return int(self.values.shape[0])


@dataclass
class R1CS:                           # r1cs.nim:64-80, matrices as COO arrays (row, col, standard-form value)
    nWires: int
    nPubOut: int
    nPubIn: int
    nPrivIn: int
    nConstr: int
    rows: tuple                       # (A, B, C) uint32 arrays
    cols: tuple
    vals: tuple                       # (nnz, 4) uint64 arrays
    nLabels: int = 0
    wireToLabel: Optional[np.ndarray] = None


@dataclass
class Proof:                          # prover.nim:38-43
    publicIO: np.ndarray              # (npubs+1, 4) standard form
    pi_a: np.ndarray                  # (8,)  affine Montgomery
    pi_b: np.ndarray                  # (16,)
    pi_c: np.ndarray                  # (8,)
    curve: str = "bn128"


@dataclass
class Mask:                           # prover.nim:211-213 (integers)
    r: int = 0
    s: int = 0
