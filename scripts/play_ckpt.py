#!/usr/bin/env python
"""Play the reference's shipped planning policy (trained/planning_cnn_rate.pth, stripped copy in tests/golden) in the B200
Planning env next to a randomly initialised policy of the same architecture:  python scripts/play_ckpt.py [--envs 2048]"""
import argparse, copy, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PARAMS = {
    "algo": {"name": "a2c_continuous"}, "model": {"name": "continuous_a2c_logstd"},
    "network": {"name": "actor_critic", "separate": False,
                "space": {"continuous": {"fixed_sigma": True}},
                "mlp": {"units": [64, 128, 64], "activation": "elu"}, "cnn": {"output_dim": 30}},
    "config": {"env_name": "planning", "env_config": {"use_image": True, "ctl_mode": "rate", "seed": 1}, "name": "ppo_planning",
               "normalize_input": True, "normalize_value": True, "num_actors": 2048, "clip_actions": True,
               "player": {"games_num": 4096, "deterministic": True, "print_stats": False, "max_steps": 1700}},
}

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=2048)
    ap.add_argument("--ckpt", default=os.path.join(ROOT, "tests", "golden", "planning_cnn_rate_model.pth"))
    a = ap.parse_args()
    import torch
    from airgym_b200.lib.agent.players import PpoPlayerContinuous
    out = {}
    for name in ("trained", "random"):
        p = copy.deepcopy(PARAMS)
        p["config"]["num_actors"] = a.envs
        p["config"]["player"]["games_num"] = 2 * a.envs
        torch.manual_seed(0)
        pl = PpoPlayerContinuous(p)
        if name == "trained":
            pl.restore(a.ckpt)
        r, s = pl.run()
        e = pl.vec_env.env
        out[name] = {"av_reward": round(r, 2), "av_steps": round(s, 1), "games": pl.games_played}
    print(json.dumps(out))

if __name__ == "__main__":
    main()
