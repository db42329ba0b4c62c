"""Player / checkpoint compatibility: the reference's shipped policy checkpoint
(trained/planning_cnn_rate.pth, stripped to its model state dict by tests/golden/make_golden_ckpt.py) loads key for key into
the B200 model and plays in the B200 Planning env through the reference's `--play` entry point."""
import copy
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
CKPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "planning_cnn_rate_model.pth")

PARAMS = {
    "algo": {"name": "a2c_continuous"}, "model": {"name": "continuous_a2c_logstd"},
    "network": {"name": "actor_critic", "separate": False, "space": {"continuous": {"fixed_sigma": True}},
                "mlp": {"units": [64, 128, 64], "activation": "elu"}, "cnn": {"output_dim": 30}},
    "config": {"env_name": "planning", "env_config": {"use_image": True, "ctl_mode": "rate", "seed": 1}, "name": "ppo_planning",
               "normalize_input": True, "normalize_value": True, "num_actors": 64, "clip_actions": True, "reward_shaper": {"scale_value": 0.1},
               "player": {"games_num": 64, "deterministic": True, "print_stats": False, "max_steps": 60}},
}


def test_reference_checkpoint_loads_and_plays():
    from airgym_b200.lib.torch_runner import Runner

    r = Runner()
    r.load({"params": copy.deepcopy(PARAMS)})
    av_reward, av_steps = r.run({"play": True, "train": False, "checkpoint": CKPT})
    model = r.player.model
    ck = torch.load(CKPT, weights_only=False)["model"]
    assert set(model.state_dict().keys()) == set(ck.keys())  # the reference's .pth key layout, key for key
    for k, v in ck.items():
        assert torch.allclose(model.state_dict()[k].cpu().double(), v.double(), rtol=1e-6, atol=1e-7), k
    assert r.player.games_played >= 64 and av_steps > 1 and torch.isfinite(torch.tensor(av_reward))
    # the model's forward is the reference's: [obs16 | cnn(norm(image))] -> norm -> MLP -> mu (a2c_continuous_logstd_model.py:139-150)
    obs = r.player.env.reset()
    with torch.no_grad():
        res = model({"is_train": False, "obs": obs})
        img = model.running_mean_std.running_mean_std["image"](obs["image"])
        x = torch.cat((obs["observation"], model.actor_cnn(img)), -1)
        h = model.actor_mlp(model.running_mean_std.running_mean_std["observation"](x))
    # the model fuses the image normalisation into the encoder kernel ((x - mean) * rsqrt instead of a division): rounding-level differences
    assert torch.allclose(res["mus"], model.mu(h), atol=1e-4)


def test_pretrained_mlp_checkpoint_initialises_cnn_policy(tmp_path):
    """players.py:387-428: an MLP-only checkpoint (obs 46 wide) fills everything but the CNN."""
    from airgym_b200.lib.agent.players import PpoPlayerContinuous

    ck = torch.load(CKPT, weights_only=False)
    w = {k: v for k, v in ck["model"].items() if "cnn" not in k and "image" not in k}
    w = {k.replace("running_mean_std.running_mean_std.observation.", "running_mean_std."): v for k, v in w.items()}
    fn = str(tmp_path / "mlp_only.pth")
    torch.save({"model": w}, fn)
    p = copy.deepcopy(PARAMS)
    pl = PpoPlayerContinuous(p)
    before = pl.model.actor_cnn.fc.weight.clone()
    pl.restore(fn)
    assert torch.equal(pl.model.actor_cnn.fc.weight, before)
    assert torch.allclose(pl.model.mu.weight.cpu(), ck["model"]["mu.weight"])
    assert torch.allclose(pl.model.running_mean_std.running_mean_std["observation"].running_mean.cpu(),
                          ck["model"]["running_mean_std.running_mean_std.observation.running_mean"])


@pytest.mark.parametrize("task,fused", [("avoid", True), ("planning", True), ("planning", False)])
def test_ppo_trains_on_image_tasks(task, fused):
    """ppo_avoid / ppo_planning.yaml shape (CNN 30 + MLP, dict observations) through Runner.run: a few epochs run, statistics stay
    finite, the trunk weights move, the CNN's do not (the reference's no_grad normalisation gives the encoder no gradient)."""
    from airgym_b200.lib.config import default_ppo_config, scale_minibatch
    from airgym_b200.lib.torch_runner import Runner

    cfg = scale_minibatch(default_ppo_config(task), 256)
    c = cfg["params"]["config"]
    c.update(max_epochs=3, print_stats=False, save_frequency=0, save_best_after=10**9, train_dir="/tmp/agx_runs", horizon_length=8,
             fused_mlp=fused)  # fused_mlp False: torch fp32 MLP + autograd on the cached trunk input (model.heads takes the tensor)
    c["minibatch_size"] = 256 * 8 // 4
    c["env_config"].update(ctl_mode="rate", num_envs=256, seed=1)
    cfg["params"]["seed"] = 1
    r = Runner()
    r.load(cfg)
    r.run({"train": True})
    ag = r.agent
    assert ag.has_cnn and ag.obs_shape == (46,)
    h = ag.history
    assert len(h) == 3 and all(torch.isfinite(torch.tensor([x["a_loss"], x["c_loss"], x["kl"]])).all() for x in h)
    rms = ag.model.running_mean_std.running_mean_std
    assert float(rms["image"].count) == 1 + 3 * 8 * 256 and float(rms["observation"].count) == 1 + 3 * 8 * 256
    sd = ag.model.state_dict()
    assert set(k for k in sd if "cnn" in k) == {k for k in torch.load(CKPT, weights_only=False)["model"] if "cnn" in k}
