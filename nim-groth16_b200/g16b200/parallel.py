"""Multi-GPU proving: one process per GPU (torch.distributed), each owning a contiguous point range of
every MSM -- the chunking of msm.nim:107-115 lifted from CPU threads to devices.  The only exchange
is an all-gather of one 384-byte record of partial sums per rank (msm.nim:117-119).  The blinding scalars are
announced before the partial sums (g16_ctx_set_mask), so that the two MSM-dependent scalar multiplications of
prover.nim:298-299 are done per rank on its own partial sums, next to its MSMs, and not after the exchange."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _lib
from .encoding import FORM_STD
from .prover import MEM_DEVICE, MEM_HOST, ProverContext
from .zkey_types import Mask, Proof, ZKey


def shard_range(n: int, k: int, g: int):
    """[N*k/G, N*(k+1)/G) with the last shard taking the remainder (msm.nim:107-111)."""
    lo = (n * k) // g
    hi = n if k == g - 1 else (n * (k + 1)) // g
    return lo, hi


def shard_ranges(nvars: int, domain_size: int, k: int, g: int):
    """g16_shard_ranges: (v_lo, v_hi, h_lo, h_hi) a ProverContext(zkey, k, g) owns -- contiguous ranges of the
    witness-indexed arrays and of the H array.  From four ranks up the H array (with buildABC and the quotient) goes
    to the first ranks only, which get a smaller share of the witness arrays; G16_SHARD_POLICY=uniform restores
    shard_range() for every array."""
    import ctypes as C
    out = (C.c_uint64 * 4)()
    _lib.check(_lib.load().g16_shard_ranges(nvars, domain_size, k, g, out))
    return tuple(int(x) for x in out)


def gather_partials(local: "torch.Tensor", group=None) -> "torch.Tensor":   # noqa: F821
    """All-gather of the per-rank partial-sum records (uint8[384]) -> uint8[world, 384]."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    flat = torch.empty(world * local.numel(), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(flat, local.contiguous().view(-1), group=group)
    return flat.view(world, local.numel())


class ShardedProver:
    """A prover context per rank; prove() returns the full proof on every rank."""

    def __init__(self, zkey: ZKey, rank: int, world: int, device: Optional[int] = None,
                 share: Optional["ShardedProver"] = None):
        import torch
        self.rank, self.world = rank, world
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        _lib.check(_lib.load().g16_set_device(self.device.index))
        # `share`: another slot over the same resident shard (g16_ctx_clone) for proofs in flight
        self.ctx = share.ctx.clone() if share is not None else ProverContext(zkey, rank, world)
        self.partials = torch.zeros(_lib.PARTIALS_BYTES, dtype=torch.uint8, device=self.device)

    def prove_raw(self, witness_ptr: int, mem_kind: int, mask: Mask, group=None):
        import torch
        self.ctx.set_mask(mask)              # every rank: its share of s*pi_a + r*rho is computed next to its MSMs
        self.ctx.prove_partials(witness_ptr, FORM_STD, mem_kind, self.partials.data_ptr())
        if self.world > 1:
            allp = gather_partials(self.partials, group)
            torch.cuda.current_stream().synchronize()
        else:
            allp = self.partials.view(1, -1)
        return self.ctx.prove_finish(allp.data_ptr(), self.world, mask)

    # asynchronous halves, for overlapping consecutive proofs (one ShardedProver per proof in flight)
    def partials_submit(self, witness_ptr: int, mem_kind: int, mask: Optional[Mask] = None):
        """`mask`: announce r, s now (g16_ctx_set_mask) -- the same mask must then be given to complete(), and
        every rank must do the same."""
        if mask is not None:
            self.ctx.set_mask(mask)
        _lib.check(_lib.load().g16_prove_partials_submit(self.ctx._h, witness_ptr, FORM_STD, mem_kind,
                                                         self.partials.data_ptr()))

    def complete(self, mask: Mask, group=None):
        """Waits for this rank's partial sums, all-gathers them, assembles; returns the raw proof."""
        import ctypes as C
        import torch
        lib = _lib.load()
        _lib.check(lib.g16_prove_partials_wait(self.ctx._h, None))
        if self.world > 1:
            allp = gather_partials(self.partials, group)
            torch.cuda.current_stream().synchronize()
        else:
            allp = self.partials.view(1, -1)
        from .prover import _limbs4
        _lib.check(lib.g16_prove_finish_submit(self.ctx._h, allp.data_ptr(), self.world, _limbs4(mask.r),
                                               _limbs4(mask.s)))
        raw = _lib.ProofRaw()
        _lib.check(lib.g16_prove_wait(self.ctx._h, C.byref(raw), None))
        return raw

    def prove(self, witness: np.ndarray, mask: Mask, group=None) -> Proof:
        w = np.ascontiguousarray(witness, dtype=np.uint64).reshape(-1, 4)
        raw = self.prove_raw(w.ctypes.data, MEM_HOST, mask, group)
        return self.ctx._proof(raw, w, FORM_STD)

    def close(self):
        self.ctx.close()
