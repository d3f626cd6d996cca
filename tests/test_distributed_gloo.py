"""
world_size-2 gloo tests (CPU) of the multi-GPU host logic: chain sharding, gathers, and the
reduction of diagnostic partial sums.  The sampling kernels themselves need no collective.
"""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from geosss_b200 import diagnostics, distributed
    from oracle import ess as E
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, n, d = 37, 400, 3  # odd chain count: blocks of 19 and 18
        gen = np.random.default_rng(5)
        full = np.zeros((N, n, d))
        noise = gen.standard_normal(full.shape)
        for t in range(1, n):
            full[:, t] = 0.6 * full[:, t - 1] + noise[:, t]
        off, cnt = distributed.shard_range(N, rank, world)
        local = torch.as_tensor(full[off:off + cnt])
        # gather restores global chain order
        got = distributed.gather_chains(local, N)
        ok_gather = bool(np.array_equal(got.numpy(), full))
        # moments over all ranks == numpy over the full array
        mean, var = distributed.reduce_moments(local)
        flat = full.reshape(-1, d)
        ok_mom = bool(np.allclose(mean.numpy(), flat.mean(0)) and np.allclose(var.numpy(), flat.var(0)))
        # ESS partial sums are additive over shards: reduce, finish, compare with the oracle on all chains
        L = 150
        tr = full[off:off + cnt, :, 0]
        ac = E.autocov_multichain(tr, L)
        part = np.concatenate([ac.sum(0), [tr.mean(1).sum(), (tr.mean(1) ** 2).sum()]])
        tot = distributed.reduce_sum(part).numpy()
        ess, rel = diagnostics.ess_from_partials(tot, n, N)
        want = E.ess_multichain(full[:, :, 0], L)
        ok_ess = bool(abs(ess - want) <= 1e-9 * want)
        # initial-state sharding hands every rank its block and offset
        x0 = np.arange(N * d, dtype=float).reshape(N, d)
        loc, o = distributed.shard_initial_states(x0, rank, world)
        ok_shard = bool(o == off and np.array_equal(loc, x0[off:off + cnt]))
        q.put((rank, ok_gather, ok_mom, ok_ess, ok_shard))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_every_chain_once():
    sys.path.insert(0, ROOT)
    from geosss_b200 import distributed
    for n in (1, 7, 8, 65536, 2**20 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [distributed.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (o1, c1), (o2, _) in zip(spans, spans[1:]):
                assert o1 + c1 == o2
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


@pytest.mark.timeout(120)
def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for r in res:
        assert all(r[1:]), r
