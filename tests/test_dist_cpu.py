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


# ---- moment merging of the sharded PPO update (airgym_b200/lib/core/moments.py) vs the oracle on the concatenated batch ---------
def _moments_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from airgym_b200.lib.core.moments import batch_sums, moments_from_sums, sums_from_moments
    from airgym_b200.lib.core.running_mean_std import RunningMeanStd

    g = torch.Generator().manual_seed(9)
    full = torch.randn(2 * 1000, 18, generator=g) * 3.0 + 1.5   # every rank can rebuild the global batch
    adv_full = torch.randn(2 * 1000, generator=g) * 0.7 - 0.2
    mine, adv = full[rank * 1000:(rank + 1) * 1000], adv_full[rank * 1000:(rank + 1) * 1000]
    rms = RunningMeanStd((18,))
    for _ in range(2):  # two updates: the merge with the running statistics is exercised too
        s = batch_sums(mine)
        dist.all_reduce(s)
        mean, var = moments_from_sums(s, 2000)
        rms.update_from_moments(mean, var, 2000)
    # cached (mean, var, n) of a shard -> sums -> merged moments (the image-statistics path)
    v, m = torch.var_mean(mine.double(), dim=0)
    s2 = sums_from_moments(m, v, 1000)
    dist.all_reduce(s2)
    mean2, var2 = moments_from_sums(s2, 2000)
    sa = batch_sums(adv)
    dist.all_reduce(sa)
    am, av = moments_from_sums(sa, 2000)
    adv_n = (adv - am.float()) / (torch.sqrt(av).float() + 1e-8)
    gathered = [None] * world
    dist.all_gather_object(gathered, adv_n)
    if rank == 0:
        q.put((rms.running_mean, rms.running_var, rms.count, mean2, var2, torch.cat(gathered), full, adv_full))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_merge_moments_like_the_oracle_on_the_global_batch():
    from oracle import ppo as O

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_moments_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    mean, var, count, mean2, var2, adv_n, full, adv_full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m, v, c = torch.zeros(18, dtype=torch.float64), torch.ones(18, dtype=torch.float64), torch.ones((), dtype=torch.float64)
    m32, v32, c32 = m.clone(), v.clone(), c.clone()
    for _ in range(2):
        m, v, c = O.rms_update(m, v, c, full.double())  # the reference's update on the GLOBAL batch (running_mean_std.py:45-60)
        m32, v32, c32 = O.rms_update(m32, v32, c32, full)  # ... with its fp32 batch moments
    assert float(c) == float(count) == 4001.0
    assert float((mean - m).abs().max()) < 1e-12 and float((var - v).abs().max()) < 1e-11
    assert float((mean - m32).abs().max()) < 1e-6 and float((var - v32).abs().max()) < 1e-5
    v_ref, m_ref = torch.var_mean(full.double(), dim=0)
    assert float((mean2 - m_ref).abs().max()) < 1e-12 and float((var2 - v_ref).abs().max()) < 1e-11
    ref = (adv_full - adv_full.mean()) / (adv_full.std() + 1e-8)   # a2c_continuous.py:160-164 on the global batch
    assert float((adv_n - ref).abs().max()) < 2e-6
