"""Multi-GPU PPO (reference lib/agent/a2c_base.py:293-309, a2c_continuous.py:112-123,160-164,188-192) on 2 GPUs:
skipped on a single-GPU box.  torchrun launches tests/mgpu_worker.py with one rank per GPU; the worker reports
(a) the peer-memory all-reduce across processes is exact and identical on both ranks,
(b) 2 ranks x N/2 envs reproduce the 1-rank run over N envs when each mini-epoch is one minibatch (global advantage and
    RunningMeanStd moments + summed gradients make the two runs equal up to fp32 summation order),
(c) with the production minibatching the replicas stay bitwise identical, for the peer-memory path and for NCCL (captured
    in the graphs, and eager)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_ppo(built, tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "report.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py"), str(out)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-6000:]
    rep = json.loads(out.read_text())
    print(rep)
    assert rep["collective_exact"]
    assert rep["replicas_identical_single_mb"] and rep["count_equal"]
    assert rep["single_mb_update_size"] > 1e-3, "the parameters must have moved for the comparison to mean anything"
    assert rep["single_mb_param_err"] < 2e-4 * max(1.0, rep["single_mb_update_size"] / 1e-2), rep
    # the two runs diverge at fp32 summation-order level after the first update, so epoch 2 sees slightly different observations
    assert rep["obs_mean_err"] < 1e-6 and rep["obs_var_err"] < 1e-6 and rep["val_mean_err"] < 1e-5
    for k in ("peer_graph", "nccl_graph", "nccl_eager"):
        assert rep[f"replicas_identical_{k}"] and rep[f"finite_{k}"], k
    assert rep["peer_vs_nccl_eager"] < 5e-3 and rep["nccl_graph_vs_eager"] < 5e-3, rep
