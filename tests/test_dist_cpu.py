"""N>1 host logic on CPU: 2 `gloo` ranks each step their env shard (host build of the kernel body, in-kernel Philox keyed
by the GLOBAL env id via env_offset) and must reproduce the single-process run exactly; timing is reduced as max over ranks."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from airgym_b200 import _capi
from airgym_b200.dist_utils import max_over_ranks, shard

N_TOTAL, STEPS, SEED = 301, 12, 77


def _run_shard(offset, count):
    from tests.hostsim.driver import HostEnv

    P = _capi.default_params("hovering", "rate")
    P.flags &= ~_capi.FLAG_MUTATE_ACTIONS
    he = HostEnv(P, count)
    g = np.random.default_rng(5)
    acts = g.uniform(-1, 1, size=(STEPS, N_TOTAL, 4)).astype(np.float32)
    for t in range(STEPS):
        if t == 6:
            he.progress[:] = P.max_episode_length - 2  # force a reset wave → post-step Philox reset draws
        he.step(np.ascontiguousarray(acts[t, offset:offset + count]), seed=SEED, step=t, env_offset=offset)
    return he.state.copy(), he.obs.copy(), he.reward.copy(), he.reset.copy()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    offset, count = shard(N_TOTAL, rank, world)
    state, obs, rew, reset = _run_shard(offset, count)
    ms = max_over_ranks(10.0 + rank, dist)  # rank-dependent "time": everyone must see the max
    gathered = [None] * world
    dist.all_gather_object(gathered, (offset, state, obs, rew, reset))
    if rank == 0:
        q.put((ms, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_partitions_cover_the_axis():
    for n, w in ((65536, 8), (301, 2), (7, 3), (5, 8)):
        spans = [shard(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (o1, c1), (o2, _) in zip(spans, spans[1:]):
            assert o1 + c1 == o2


def test_two_gloo_ranks_equal_single_process(built):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ms, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 11.0
    full = _run_shard(0, N_TOTAL)
    for offset, state, obs, rew, reset in gathered:
        n = state.shape[0]
        assert np.array_equal(state, full[0][offset:offset + n])
        assert np.array_equal(obs, full[1][offset:offset + n])
        assert np.array_equal(rew, full[2][offset:offset + n], equal_nan=True)
        assert np.array_equal(reset, full[3][offset:offset + n])
    assert sum(g[1].shape[0] for g in gathered) == N_TOTAL
