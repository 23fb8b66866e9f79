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
    """Equal contiguous column blocks; the last block may be partly padding.

    ``tail_fraction`` > 0 splits the columns into two PHASES, each sharded over all ranks: phase 0 = the first
    ``~(1 - tail_fraction)`` of the columns (a multiple of ``world``, so it has no padding), phase 1 = the rest.  Every
    phase is one contiguous block of the full matrix and every rank's share of it one contiguous slab inside that block, so
    the all-gather of a phase stays in place and can run while the next phase is computed (closed-form models, whose
    all-gather is not negligible next to the kernel)."""
    nspp: int
    world: int
    tail_fraction: float = 0.0

    def phase_list(self):
        """[(first column, columns, shard)] per phase."""
        w = max(self.world, 1)
        head = 0
        if self.tail_fraction > 0.0 and self.world > 1:
            head = int(self.nspp * (1.0 - self.tail_fraction)) // w * w
            if head <= 0 or head >= self.nspp:
                head = 0
        out = []
        if head:
            out.append((0, head, head // w))
        rest = self.nspp - head
        out.append((head, rest, (rest + w - 1) // w if self.world > 0 else 0))
        return out

    @property
    def shard(self) -> int:
        """Columns per rank (single phase) / in the last phase."""
        return self.phase_list()[-1][2]

    @property
    def padded(self) -> int:
        start, _, shard = self.phase_list()[-1]
        return start + shard * self.world

    def range(self, rank: int, phase: int = -1):
        start, count, shard = self.phase_list()[phase]
        lo = min(start + rank * shard, start + count)
        hi = min(lo + shard, start + count)
        return lo, hi

    def ranges(self, rank: int):
        return [self.range(rank, p) for p in range(len(self.phase_list()))]

    def counts(self):
        return [sum(hi - lo for lo, hi in self.ranges(r)) for r in range(self.world)]


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

    def __init__(self, nsub: int, nspp: int, device, dtype=None, group=None, peer_stores=False, tail_fraction=0.0):
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        # the fused path has no collective to overlap: phases only make sense for the NCCL / gloo all-gather
        self.part = ColumnPartition(int(nspp), self.world, 0.0 if peer_stores else float(tail_fraction))
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
        """(lo, hi) of this rank's columns when there is one phase (the common case)."""
        return self.part.range(self.rank)

    @property
    def local_ranges(self):
        return self.part.ranges(self.rank)

    @property
    def nphases(self):
        return len(self.part.phase_list())

    def local_slab(self, phase: int = -1):
        """This rank's block of columns of one phase as a view into the full matrix (shard x nsub, padding included)."""
        start, _, s = self.part.phase_list()[phase]
        return self.full[start + self.rank * s:start + (self.rank + 1) * s]

    def gather_phase(self, phase: int, async_op: bool = False):
        """In-place all-gather of one phase's slabs; returns the work handle when asynchronous."""
        if self.world <= 1 or self.peer_ptrs is not None:
            return None
        start, _, s = self.part.phase_list()[phase]
        return self.dist.all_gather_into_tensor(self.full[start:start + s * self.world], self.local_slab(phase), group=self.group, async_op=async_op)

    def gather(self):
        """In-place all-gather of the slabs (each rank's input IS its slice of the output); with peer stores the
        kernel has already written every rank's matrix and only a device barrier over the ranks remains."""
        if self.world > 1:
            if self.peer_ptrs is not None:
                self.symm.barrier()
            else:
                for p in range(self.nphases):
                    self.gather_phase(p)
        return self.full

    def reduce_error(self, code: int, pair: int):
        import torch
        self._err.fill_(pack_error(code, pair))
        if self.world > 1:
            self.dist.all_reduce(self._err, op=self.dist.ReduceOp.MIN, group=self.group)
        return unpack_error(int(self._err.item()))

    def run(self, evaluate):
        """Evaluate phase after phase; the all-gather of a phase runs while the next phase is evaluated."""
        works, first = [], (0, -1)
        for p, (lo, hi) in enumerate(self.local_ranges):
            code, pair = evaluate(lo, hi - lo, self.local_slab(p)[: hi - lo]) if hi > lo else (0, -1)
            if code and (first[0] == 0 or pair < first[1]):
                first = (code, pair)
            if self.peer_ptrs is None:
                works.append(self.gather_phase(p, async_op=True))
        if self.peer_ptrs is not None:
            self.gather()
        for w in works:
            if w is not None:
                w.wait()
        return self.reduce_error(*first)

    def matrix(self):
        """(nsub, nspp) F-order view semantics: returns the (nspp, nsub) C-order tensor transposed."""
        return self.full[: self.nspp].t()
