"""Worker of tests/test_gpu_multi.py — launched by torchrun with one rank per GPU; rank 0 writes a JSON report."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from airgym_b200.comm import PeerComm  # noqa: E402
from airgym_b200.lib.config import default_ppo_config, scale_minibatch  # noqa: E402
from airgym_b200.lib.agent.a2c_continuous import A2CAgent  # noqa: E402
from airgym_b200.lib.utils import tr_helpers  # noqa: E402

N_TOTAL, H, SEED = 4096, 8, 11


def make_agent(num_envs, minibatch, multi_gpu, comm_kind="peer", graph=True, task="hovering"):
    cfg = scale_minibatch(default_ppo_config(task), num_envs)
    c = cfg["params"]["config"]
    c.update(horizon_length=H, minibatch_size=minibatch, multi_gpu=multi_gpu, multi_gpu_comm=comm_kind, use_cuda_graph=graph,
             print_stats=False, write_summaries=False, train_dir="/tmp/agx_mgpu", save_frequency=0, save_best_after=10**9,
             device=f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}")
    c["env_config"].update(ctl_mode="rate", num_envs=num_envs, seed=SEED)
    c["reward_shaper"] = tr_helpers.DefaultRewardsShaper(**c["reward_shaper"])
    torch.manual_seed(SEED)
    return A2CAgent("run", cfg["params"])


def run_epochs(agent, epochs, noise):
    agent.noise_table = noise
    agent.env_reset()
    if agent.multi_gpu and agent.world_size > 1:
        dist.broadcast(agent.flat_params, 0)
    for _ in range(epochs):
        agent.train_epoch()
    torch.cuda.synchronize()


def main():
    out_path = sys.argv[1]
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    report = {"world": world}

    # ---- (a) the collective itself across processes (CUDA IPC + NVLink)
    comm = PeerComm(rank, world, 1 << 20, dev)
    ok = True
    for call in range(6):
        g = torch.Generator(device=dev).manual_seed(100 + call)
        allr = [torch.randn(18129, device=dev, generator=g) * (r + 1) for r in range(world)]  # every rank can build every rank's data
        mine = allr[rank].clone()
        comm.all_reduce(mine)
        want = allr[0].clone()
        for r in range(1, world):
            want += allr[r]
        ok &= bool(torch.equal(mine, want))
        d = torch.full((36,), float(rank + 1), device=dev, dtype=torch.float64)
        comm.all_reduce(d)
        ok &= bool((d == world * (world + 1) / 2).all())
    seq, err = comm.status()
    report["collective_exact"] = ok and err == 0 and seq == 12
    comm.close()

    # ---- (b) one minibatch per mini-epoch: `world` ranks x N/world envs must reproduce the 1-rank run over N envs
    n_local = N_TOTAL // world
    gen = torch.Generator(device="cpu").manual_seed(5)
    noise_all = torch.randn(H, N_TOTAL, 4, generator=gen)
    agent = make_agent(n_local, n_local * H, True)
    init = agent.flat_params.clone()
    dist.broadcast(init, 0)
    agent.flat_params.copy_(init)
    run_epochs(agent, 2, noise_all[:, rank * n_local:(rank + 1) * n_local].to(dev).contiguous())
    gathered = [torch.empty_like(agent.flat_params) for _ in range(world)]
    dist.all_gather(gathered, agent.flat_params)
    report["replicas_identical_single_mb"] = all(bool(torch.equal(g, gathered[0])) for g in gathered)
    rms = agent.model.running_mean_std
    sharded = {"params": agent.flat_params.clone(), "obs_mean": rms.running_mean.clone(), "obs_var": rms.running_var.clone(),
               "val_mean": agent.value_mean_std.running_mean.clone(), "lr": float(agent.lr_dev), "count": float(rms.count)}
    agent.comm.close()
    del agent
    if rank == 0:
        ref = make_agent(N_TOTAL, N_TOTAL * H, False)
        ref.flat_params.copy_(init)
        run_epochs(ref, 2, noise_all.to(dev).contiguous())
        rr = ref.model.running_mean_std
        scale = float(ref.flat_params.abs().max())
        report["single_mb_param_err"] = float((ref.flat_params - sharded["params"]).abs().max()) / scale
        report["single_mb_update_size"] = float((ref.flat_params - init).abs().max()) / scale
        report["obs_mean_err"] = float((rr.running_mean - sharded["obs_mean"]).abs().max())
        report["obs_var_err"] = float((rr.running_var - sharded["obs_var"]).abs().max())
        report["val_mean_err"] = float((ref.value_mean_std.running_mean - sharded["val_mean"]).abs().max())
        report["count_equal"] = float(rr.count) == sharded["count"]
        report["lr_equal"] = float(ref.lr_dev) == sharded["lr"]
        del ref
    dist.barrier()

    # ---- (c) the production shape (48 minibatches, graphs): peer-memory path vs captured NCCL vs eager NCCL
    finals = {}
    for kind, graph in (("peer", True), ("nccl", True), ("nccl", False)):
        agent = make_agent(n_local, n_local * H // 8, True, comm_kind=kind, graph=graph)
        agent.flat_params.copy_(init)
        run_epochs(agent, 4, noise_all[:, rank * n_local:(rank + 1) * n_local].to(dev).contiguous())
        gathered = [torch.empty_like(agent.flat_params) for _ in range(world)]
        dist.all_gather(gathered, agent.flat_params)
        report[f"replicas_identical_{kind}_{'graph' if graph else 'eager'}"] = all(bool(torch.equal(g, gathered[0])) for g in gathered)
        finals[(kind, graph)] = agent.flat_params.clone()
        report[f"finite_{kind}_{'graph' if graph else 'eager'}"] = bool(torch.isfinite(agent.flat_params).all())
        if agent.comm is not None:
            agent.comm.close()
        del agent
        dist.barrier()
    scale = float(finals[("nccl", False)].abs().max())
    report["peer_vs_nccl_eager"] = float((finals[("peer", True)] - finals[("nccl", False)]).abs().max()) / scale
    report["nccl_graph_vs_eager"] = float((finals[("nccl", True)] - finals[("nccl", False)]).abs().max()) / scale
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(report, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
