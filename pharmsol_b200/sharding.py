"""Multi-GPU psi: support-point COLUMNS shard across ranks.

Every (subject, support point) pair is independent, so the computation itself needs no exchange.
psi is column-major (likelihood/matrix.rs:60), hence a contiguous block of columns is one contiguous
slab of the output: rank r evaluates columns [r*shard, (r+1)*shard) straight into its slab of the
full matrix and ONE all-gather (NCCL over NVLink/NVSwitch; in place, no repack) reassembles psi on
every rank.  The flattened population is small and replicated on every GPU.

One process per GPU; ``torch.distributed`` is plumbing only (process group, the all-gather, and the
device memory the slabs live in).  The host-side logic (partition, slab views, error reduction) is
backend-agnostic and covered by world_size-2 gloo tests on CPU with a stub column evaluator.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class ColumnPartition:
    """Equal contiguous column blocks; the last block may be partly padding."""
    nspp: int
    world: int

    @property
    def shard(self) -> int:
        return (self.nspp + self.world - 1) // self.world if self.world > 0 else 0

    @property
    def padded(self) -> int:
        return self.shard * self.world

    def range(self, rank: int):
        lo = min(rank * self.shard, self.nspp)
        hi = min(lo + self.shard, self.nspp)
        return lo, hi

    def counts(self):
        return [self.range(r)[1] - self.range(r)[0] for r in range(self.world)]


def pack_error(code: int, pair: int) -> int:
    """(pair << 8 | code) so that MIN over ranks yields the first failing pair (matrix.rs:96-104);
    no error = int64 max."""
    return (int(pair) << 8) | (int(code) & 0xFF) if code else np.iinfo(np.int64).max


def unpack_error(word: int):
    if word == np.iinfo(np.int64).max:
        return 0, -1
    return int(word) & 0xFF, int(word) >> 8


class ShardedPsi:
    """Column-sharded psi across the ranks of the default process group.

    ``evaluate(first_col, ncols, slab)`` is the per-rank column evaluator: it must fill ``slab`` — a
    (ncols, nsub) C-order tensor view, i.e. the column-major block psi[:, first_col:first_col+ncols] —
    and return ``(error_code, first_failing_global_pair)``.
    """

    def __init__(self, nsub: int, nspp: int, device, dtype=None, group=None, peer_stores=False):
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.part = ColumnPartition(int(nspp), self.world)
        self.nsub, self.nspp = int(nsub), int(nspp)
        self.device = device
        # column-major psi == C-order (columns, nsub); padded to world * shard columns
        # Fused all-gather: allocate psi in symmetric memory so every rank's full matrix is mapped into every
        # process (NVLink / NVSwitch peer pointers); the psi kernel then stores each result to all ranks directly
        # and `gather` degenerates to a device barrier.  Falls back to the NCCL all-gather if the rendezvous fails.
        self.symm, self.peer_ptrs = None, None
        self.full = None
        if peer_stores and self.world > 1 and torch.device(device).type == "cuda":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self.full = symm_mem.empty((self.part.padded, self.nsub), dtype=dtype or torch.float64, device=device)
                self.symm = symm_mem.rendezvous(self.full, group if group is not None else dist.group.WORLD)
                self.peer_ptrs = [int(q) for q in self.symm.buffer_ptrs]
            except Exception as e:      # noqa: BLE001 - any failure means "no peer mapping available"
                self.symm, self.peer_ptrs, self.full = None, None, None
                self.peer_error = repr(e)
        if self.full is None:
            self.full = torch.empty((self.part.padded, self.nsub), dtype=dtype or torch.float64, device=device)
        self._err = torch.zeros(1, dtype=torch.int64, device=device)

    @property
    def local_range(self):
        return self.part.range(self.rank)

    def local_slab(self):
        """This rank's block of columns as a view into the full matrix (shard x nsub, padding included)."""
        s = self.part.shard
        return self.full[self.rank * s:(self.rank + 1) * s]

    def gather(self):
        """In-place all-gather of the slabs (each rank's input IS its slice of the output); with peer stores the
        kernel has already written every rank's matrix and only a device barrier over the ranks remains."""
        if self.world > 1:
            if self.peer_ptrs is not None:
                self.symm.barrier()
            else:
                self.dist.all_gather_into_tensor(self.full, self.local_slab(), group=self.group)
        return self.full

    def reduce_error(self, code: int, pair: int):
        import torch
        self._err.fill_(pack_error(code, pair))
        if self.world > 1:
            self.dist.all_reduce(self._err, op=self.dist.ReduceOp.MIN, group=self.group)
        return unpack_error(int(self._err.item()))

    def run(self, evaluate):
        lo, hi = self.local_range
        slab = self.local_slab()
        code, pair = evaluate(lo, hi - lo, slab[: hi - lo]) if hi > lo else (0, -1)
        self.gather()
        return self.reduce_error(code, pair)

    def matrix(self):
        """(nsub, nspp) F-order view semantics: returns the (nspp, nsub) C-order tensor transposed."""
        return self.full[: self.nspp].t()
