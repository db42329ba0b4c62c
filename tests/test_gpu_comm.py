"""Peer-memory all-reduce (include/agx.h multi-GPU section) on ONE GPU: `world` communicators whose regions live in the same
device memory, every rank's collective issued on its own stream so the kernels run concurrently and talk through the same
flags / slots a multi-GPU run uses through NVLink.  The real IPC path (2 processes, 2 GPUs) is tests/test_gpu_multi.py."""
import ctypes as C

import pytest
import torch

from airgym_b200 import _capi
from airgym_b200.comm import make_local_group

pytestmark = pytest.mark.gpu


def _free(comms):
    torch.cuda.synchronize()
    for c in comms:
        c._lib.agx_comm_free(c._own)
        c._own = None


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("dtype,n", [(torch.float32, 18129), (torch.float64, 36), (torch.float32, 3), (torch.float64, 50880)])
def test_allreduce_matches_sum_and_is_identical_on_every_rank(built, world, dtype, n):
    comms = make_local_group(world, n * 8, "cuda:0")
    streams = [torch.cuda.Stream() for _ in range(world)]
    g = torch.Generator(device="cuda").manual_seed(world * 1000 + n)
    for call in range(5):  # both slot parities, call counter advancing
        data = [torch.randn(n, device="cuda", dtype=dtype, generator=g) for _ in range(world)]
        want = torch.stack(data).double().sum(0)
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                comms[r].all_reduce(data[r])
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(data[r], data[0]), f"rank {r} differs from rank 0 (call {call})"
        tol = 1e-5 if dtype == torch.float32 else 1e-12
        assert float((data[0].double() - want).abs().max()) <= tol * max(1.0, float(want.abs().max()))
    for c in comms:
        seq, err = c.status()
        assert seq == 5 and err == 0
    _free(comms)


def test_allreduce_replays_from_cuda_graphs(built):
    world, n = 2, 4096
    comms = make_local_group(world, n * 4, "cuda:0")
    bufs = [torch.zeros(n, device="cuda") for _ in range(world)]
    src = [torch.full((n,), float(r + 1), device="cuda") for r in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    graphs = []
    for r in range(world):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=streams[r]):
            bufs[r].copy_(src[r])
            comms[r].all_reduce(bufs[r])
            comms[r].all_reduce(bufs[r])  # two collectives per replay: the call counter is read from the region, not baked in
        graphs.append(g)
    for it in range(4):
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                graphs[r].replay()
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(bufs[r], torch.full((n,), 6.0, device="cuda")), (it, r)  # (1 + 2) then (3 + 3)
    assert comms[0].status() == (8, 0)
    _free(comms)


def test_fused_allreduce_adam_equals_adam_on_the_summed_gradient(built):
    lib = _capi.load()
    world, n, extra = 2, 18121 + 3, _capi.AGX_PPO_STATS
    comms = make_local_group(world, (n + extra) * 4, "cuda:0")
    hp = _capi.AgxPpoHyper()
    hp.e_clip, hp.critic_coef, hp.entropy_coef, hp.bounds_loss_coef = 0.2, 2.0, 0.0, 1e-4
    hp.kl_threshold, hp.grad_norm, hp.beta1, hp.beta2, hp.eps, hp.weight_decay, hp.adaptive_lr = 0.008, 1.5, 0.9, 0.999, 1e-8, 0.0, 1
    torch.manual_seed(0)
    p0 = torch.randn(n, device="cuda")
    mk = lambda: dict(p=p0.clone(), m=torch.zeros(n, device="cuda"), v=torch.zeros(n, device="cuda"), lr=torch.full((1,), 3e-4, device="cuda"),
                      step=torch.zeros(1, device="cuda", dtype=torch.int64), norm=torch.zeros(1, device="cuda"))
    ranks, ref = [mk() for _ in range(world)], mk()
    streams = [torch.cuda.Stream() for _ in range(world)]
    ptr = lambda t: t.data_ptr()
    for it in range(6):
        gs = [torch.randn(n + extra, device="cuda") * (3.0 if it % 2 else 0.01) for _ in range(world)]
        for r in range(world):
            gs[r][n + 4] = 0.002 + 0.03 * it + 0.001 * r  # the KL slot: drives the learning-rate rule
        total = gs[0] + gs[1]  # the kernel adds the slots in rank order: same fp32 sum
        torch.cuda.synchronize()
        for r in range(world):
            s = ranks[r]
            with torch.cuda.stream(streams[r]):
                _capi.check(lib.agx_adam_step_allreduce(C.byref(hp), C.byref(comms[r].c), n, extra, ptr(s["p"]), ptr(gs[r]), ptr(s["m"]), ptr(s["v"]),
                                                        ptr(s["lr"]), ptr(s["step"]), ptr(gs[r][n + 4:n + 5]), 0.5, ptr(s["norm"]),
                                                        C.c_void_p(streams[r].cuda_stream)), "agx_adam_step_allreduce")
        _capi.check(lib.agx_adam_step(C.byref(hp), n, ptr(ref["p"]), ptr(total), ptr(ref["m"]), ptr(ref["v"]), ptr(ref["lr"]), ptr(ref["step"]),
                                      ptr(total[n + 4:n + 5]), 0.5, ptr(ref["norm"]), None), "agx_adam_step")
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(gs[r], total), "the summed gradients + statistics are written back in place"
            for k in ("p", "m", "v", "lr", "step", "norm"):
                assert torch.equal(ranks[r][k], ref[k]), (it, r, k)
    assert float(ref["lr"]) != pytest.approx(3e-4)  # the rule moved it
    _free(comms)


def test_bad_arguments_fail_loudly(built):
    lib = _capi.load()
    comms = make_local_group(2, 1024, "cuda:0")
    big = torch.zeros(1024, device="cuda")
    with pytest.raises(ValueError):
        comms[0].all_reduce(big)  # 4096 B > 1024 B slot
    assert lib.agx_comm_allreduce(C.byref(comms[0].c), big.data_ptr(), 1024, _capi.AGX_F32, None) == -1
    assert lib.agx_comm_allreduce(C.byref(comms[0].c), big.data_ptr(), 16, 7, None) == -1
    bad = _capi.AgxComm()
    bad.rank, bad.world, bad.slot_bytes = 0, 2, 1024  # regions unmapped
    assert lib.agx_comm_allreduce(C.byref(bad), big.data_ptr(), 16, _capi.AGX_F32, None) == -1
    assert lib.agx_comm_region_bytes(9, 1024) == -1 and lib.agx_comm_region_bytes(2, 1000) == 256 + 4 * 2 * 1024
    _free(comms)
