"""N>1 path on CPU: two gloo ranks shard a global batch of independent solves (no data-path
collective), and the gathered result equals the single-process result.  The per-rank "device" is the
CPU oracle here (test infrastructure); the GPU twin of this test is bench.py --gpus 2."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from idto_b200 import problems
from idto_b200.sharding import shard_slice


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _solve_shard(q0, v0, qg, iters):
    from oracle import oracle
    m, dt, prob, params, _ = problems.spinner(max_iterations=iters)
    out = []
    for b in range(q0.shape[0]):
        o = oracle.Oracle(m, dt, prob, params)
        o.reset_initial_conditions(q0[b], v0[b])
        o.set_q(qg[b])
        o.solve(iters)
        out.append(o.solution()[0])
    return np.array(out)


def _worker(rank, world, port, gb, iters, ret):
    import torch.distributed as dist
    from idto_b200.sharding import gather_batches, max_over_ranks
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, dt, prob, params, _ = problems.spinner()
    q0, v0, qg = problems.perturbed_batch(m, prob, gb)
    sl = shard_slice(gb, rank, world)
    local = _solve_shard(q0[sl], v0[sl], qg[sl], iters)
    full = gather_batches(local, gb)
    tmax = max_over_ranks(1.0 + rank)
    if rank == 0:
        ret["q"] = full
        ret["tmax"] = tmax
    dist.barrier()
    dist.destroy_process_group()


def test_shard_slices_cover_batch():
    for gb, w in ((64, 8), (5, 2), (7, 3), (3, 4)):
        idx = np.concatenate([np.arange(gb)[shard_slice(gb, r, w)] for r in range(w)])
        assert np.array_equal(idx, np.arange(gb))


def test_two_gloo_ranks_match_single_process():
    gb, iters, world = 5, 3, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), gb, iters, ret), nprocs=world, join=True)
    m, dt, prob, params, _ = problems.spinner()
    q0, v0, qg = problems.perturbed_batch(m, prob, gb)
    ref = _solve_shard(q0, v0, qg, iters)
    assert ret["tmax"] == 2.0
    assert np.array_equal(np.asarray(ret["q"]), ref)
