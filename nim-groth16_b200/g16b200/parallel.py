"""Multi-GPU proving: one process per GPU (torch.distributed), each owning a contiguous point range of
every MSM -- the chunking of msm.nim:107-115 lifted from CPU threads to devices.  The only exchange
is an all-gather of one 400-byte record of partial sums per rank (msm.nim:117-119).  The blinding scalars are
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


PLAN_FIELDS = ("a1_lo", "a1_hi", "b1_lo", "b1_hi", "c1_lo", "c1_hi", "b2_lo", "b2_hi", "h_lo", "h_hi")


def shard_plan(nvars: int, npubs: int, domain_size: int, k: int, g: int) -> dict:
    """g16_shard_plan: the contiguous point range of each of the five MSMs that ProverContext(zkey, k, g) owns
    (witness indices for A1, B1, C1, B2; domain indices for H).  The default policy places whole MSMs (or a tail /
    head of one) per rank by a cost model; the ranks with H points run buildABC and the quotient;
    G16_SHARD_POLICY=uniform restores shard_range() for every array."""
    import ctypes as C
    out = (C.c_uint64 * 10)()
    _lib.check(_lib.load().g16_shard_plan(nvars, npubs, domain_size, k, g, out))
    return dict(zip(PLAN_FIELDS, (int(x) for x in out)))


def gather_partials(local: "torch.Tensor", group=None) -> "torch.Tensor":   # noqa: F821
    """All-gather of the per-rank partial-sum records (uint8[400]) -> uint8[world, 400]."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    flat = torch.empty(world * local.numel(), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(flat, local.contiguous().view(-1), group=group)
    return flat.view(world, local.numel())


class ShardedProver:
    """A prover context per rank; prove() returns the full proof on every rank."""

    def __init__(self, zkey: ZKey, rank: int, world: int, device: Optional[int] = None,
                 share: Optional["ShardedProver"] = None, trusted: bool = False):
        import torch
        self.rank, self.world = rank, world
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        _lib.check(_lib.load().g16_set_device(self.device.index))
        # `share`: another slot over the same resident shard (g16_ctx_clone) for proofs in flight
        self.ctx = share.ctx.clone() if share is not None else ProverContext(zkey, rank, world, trusted=trusted)
        self.partials = torch.zeros(_lib.PARTIALS_BYTES, dtype=torch.uint8, device=self.device)

    def prove_raw(self, witness_ptr: int, mem_kind: int, mask: Mask, group=None):
        import torch
        self.partials_submit(witness_ptr, mem_kind, mask)
        return self.complete(mask, group)

    # asynchronous halves, for overlapping consecutive proofs (one ShardedProver per proof in flight)
    def partials_submit(self, witness_ptr: int, mem_kind: int, mask: Optional[Mask] = None):
        """`mask`: announce r, s now (g16_ctx_set_mask) -- the same mask must then be given to complete(), and
        every rank must do the same."""
        if mask is not None:
            self.ctx.set_mask(mask)
        _lib.check(_lib.load().g16_prove_partials_submit(self.ctx._h, witness_ptr, FORM_STD, mem_kind,
                                                         self.partials.data_ptr()))

    def partials_submit_host(self, witness_pinned: "torch.Tensor", mask: Optional[Mask] = None, group=None):   # noqa: F821
        """Host witness (a pinned torch tensor of nvars x 4 int64 limbs, standard form), travelling ONCE over PCIe: this
        rank uploads the rank-th of `world` slices over its own link, the slices are all-gathered over NVLink (NCCL), and
        the partial sums read the gathered device copy -- instead of every rank pulling the intervals it needs (up to
        the whole witness) through the host's PCIe lanes at the same time.  Ordered on the device like
        exchange_submit(): no host synchronisation.

        `group`: give this all-gather its OWN process group (dist.new_group()).  Collectives of one group execute in
        issue order on one NCCL stream; the record all-gather of the previous proof waits for that proof's MSMs, so on
        the same group the witness of the next proof would queue behind it and the proofs would no longer overlap."""
        import torch
        import torch.distributed as dist
        lib = _lib.load()
        n = witness_pinned.shape[0]
        per = -(-n // self.world)
        if getattr(self, "_wbuf", None) is None or self._wbuf.shape[0] != per * self.world:
            self._wbuf = torch.zeros((per * self.world, 4), dtype=torch.int64, device=self.device)
            self._wstream = torch.cuda.Stream(device=self.device)
        lo, hi = min(n, per * self.rank), min(n, per * (self.rank + 1))
        mine = self._wbuf[per * self.rank: per * (self.rank + 1)]
        with torch.cuda.stream(self._wstream):                  # not the stream the record all-gathers are ordered on
            if hi > lo:
                mine[: hi - lo].copy_(witness_pinned[lo:hi], non_blocking=True)
            if self.world > 1:
                dist.all_gather_into_tensor(self._wbuf.view(-1), mine.reshape(-1), group=group)
        _lib.check(lib.g16_ctx_order_stream(self.ctx._h, self._wstream.cuda_stream, 1))   # the context waits for the gather
        self.h2d_bytes = (hi - lo) * 32
        self.partials_submit(self._wbuf.data_ptr(), MEM_DEVICE, mask)

    def exchange_submit(self, mask: Mask, group=None):
        """Enqueues the all-gather of the partial records and the assembly behind this rank's partial sums, ordered
        on the device (g16_ctx_order_stream): the NCCL stream waits for the record, the context waits for the
        all-gather -- no host synchronisation between partials_submit() and wait()."""
        import torch
        lib = _lib.load()
        st = torch.cuda.current_stream(self.device)
        if self.world > 1:
            _lib.check(lib.g16_ctx_order_stream(self.ctx._h, st.cuda_stream, 0))
            self._gathered = gather_partials(self.partials, group)       # kept alive until wait()
            _lib.check(lib.g16_ctx_order_stream(self.ctx._h, st.cuda_stream, 1))
            allp = self._gathered
        else:
            allp = self.partials.view(1, -1)
        from .prover import _limbs4
        _lib.check(lib.g16_prove_finish_submit(self.ctx._h, allp.data_ptr(), self.world, _limbs4(mask.r),
                                               _limbs4(mask.s)))

    def wait(self):
        """The proof of the last exchange_submit(); fills ctx.last_stats (this rank's phases)."""
        return self.ctx.wait()[0]

    def complete(self, mask: Mask, group=None):
        """exchange_submit + wait: returns the raw proof."""
        self.exchange_submit(mask, group)
        return self.wait()

    def prove(self, witness: np.ndarray, mask: Mask, group=None) -> Proof:
        w = np.ascontiguousarray(witness, dtype=np.uint64).reshape(-1, 4)
        raw = self.prove_raw(w.ctypes.data, MEM_HOST, mask, group)
        return self.ctx._proof(raw, w, FORM_STD)

    def close(self):
        self.ctx.close()
