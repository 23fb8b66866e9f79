"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and 3, uneven) `gloo` process groups
with a stub column evaluator standing in for the CUDA launch.  Checks the column partition, the
in-place slab all-gather (psi is column-major so column blocks are contiguous) and the first-error
reduction over ranks (likelihood/matrix.rs:96-104 semantics)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pharmsol_b200.sharding import ColumnPartition, ShardedPsi, pack_error, unpack_error


def test_partition_covers_all_columns():
    for nspp in (0, 1, 7, 8, 1000, 20001):
        for world in (1, 2, 3, 8):
            p = ColumnPartition(nspp, world)
            ranges = [p.range(r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == nspp
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert sum(p.counts()) == nspp and p.padded >= nspp and p.padded - nspp < max(world, 1)


def test_phased_partition_keeps_every_phase_contiguous_and_in_place():
    """Two phases (7/8 + 1/8 of the columns, each sharded over all ranks): every column has exactly one owner, the head
    phase has no padding, and each rank's share of a phase is one contiguous slab inside the phase's block."""
    for nspp in (0, 1, 7, 64, 1001, 50000):
        for world in (1, 2, 3, 8):
            for tail in (0.125, 0.5):
                p = ColumnPartition(nspp, world, tail)
                phases = p.phase_list()
                assert 1 <= len(phases) <= 2 and phases[0][0] == 0 and sum(c for _, c, _ in phases) == nspp
                if len(phases) == 2:
                    assert phases[0][1] % world == 0 and phases[0][1] == phases[0][2] * world and phases[1][0] == phases[0][1]
                owner = np.full(nspp, -1)
                for r in range(world):
                    for ph, (lo, hi) in enumerate(p.ranges(r)):
                        assert np.all(owner[lo:hi] == -1)
                        owner[lo:hi] = r
                        start, count, shard = phases[ph]
                        assert lo == min(start + r * shard, start + count)
                assert np.all(owner >= 0) and sum(p.counts()) == nspp and nspp <= p.padded < nspp + max(world, 1)


def test_error_word_orders_by_pair_then_code():
    assert unpack_error(pack_error(0, 5)) == (0, -1)
    assert unpack_error(pack_error(7, 123456789012)) == (7, 123456789012)
    assert pack_error(12, 10) < pack_error(1, 11) < pack_error(0, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _truth(nsub, nspp):
    i = np.arange(nsub)[:, None]
    j = np.arange(nspp)[None, :]
    return -(i * 1000.0 + j) - 0.25


def _worker(rank, world, port, nsub, nspp, fail_at, out, tail=0.0):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sp = ShardedPsi(nsub, nspp, torch.device("cpu"), tail_fraction=tail)
        assert sp.nphases == len(ColumnPartition(nspp, world, tail).phase_list()) == (2 if tail else 1)
        truth = torch.from_numpy(_truth(nsub, nspp))

        def evaluate(first_col, ncols, slab):
            assert slab.shape == (ncols, nsub) and slab.is_contiguous()
            slab.copy_(truth[:, first_col:first_col + ncols].t())
            if fail_at is not None and first_col <= fail_at[1] < first_col + ncols:
                return 12, fail_at[0] + fail_at[1] * nsub
            return 0, -1

        code, pair = sp.run(evaluate)
        m = sp.matrix()
        ok = bool(torch.equal(m, truth)) and m.shape == (nsub, nspp) and m.t().is_contiguous()
        out.put((rank, ok, code, pair))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nspp,fail_at,tail", [(2, 10, None, 0.0), (2, 11, (3, 7), 0.0), (3, 8, (0, 2), 0.0),
                                                     (2, 37, None, 0.125), (2, 37, (1, 35), 0.125), (3, 100, (4, 5), 0.25)])
def test_sharded_assembly_gloo(world, nspp, fail_at, tail):
    """tail > 0: two phases, the all-gather of the first is asynchronous and overlaps the evaluation of the second."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 5, nspp, fail_at, q, tail)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, code, pair in res:
        assert ok, f"rank {rank}: gathered psi differs"
        if fail_at is None:
            assert (code, pair) == (0, -1)
        else:
            assert (code, pair) == (12, fail_at[0] + fail_at[1] * 5)   # every rank learns the first failing pair
