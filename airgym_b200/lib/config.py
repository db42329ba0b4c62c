"""Default PPO configurations with the schema and values of the reference's yamls (scripts/config/ppo_hovering.yaml:1-73;
ppo_tracking / ppo_balloon / ppo_avoid / ppo_planning.yaml differ in the fields default_ppo_config sets).  The reference's yaml files load unchanged through
`scripts/runner.py --config <file>`; these defaults exist so the package runs without them."""
import copy

_BASE = {
    "params": {
        "algo": {"name": "a2c_continuous"},
        "model": {"name": "continuous_a2c_logstd"},
        "load_checkpoint": False,
        "network": {
            "name": "actor_critic", "separate": False,
            "space": {"continuous": {"mu_activation": "None", "sigma_activation": "None", "mu_init": {"name": "default"},
                                     "sigma_init": {"name": "const_initializer", "val": 0}, "fixed_sigma": True}},
            "mlp": {"units": [64, 128, 64], "d2rl": False, "activation": "elu", "initializer": {"name": "default", "scale": 2}},
        },
        "config": {
            "env_name": "hovering", "env_config": {"use_image": False}, "name": "ppo_hovering",
            "reward_shaper": {"scale_value": 0.1},
            "normalize_advantage": True, "gamma": 0.99, "tau": 0.95, "ppo": True, "learning_rate": 3e-4,
            "lr_schedule": "adaptive", "kl_threshold": 0.008, "save_best_after": 10, "score_to_win": 100000,
            "grad_norm": 1.5, "entropy_coef": 0, "truncate_grads": True, "e_clip": 0.2, "clip_value": False,
            "num_actors": 4096, "horizon_length": 24, "minibatch_size": 2048, "mini_epochs": 5, "critic_coef": 2,
            "normalize_input": True, "bounds_loss_coef": 0.0001, "max_epochs": 200, "normalize_value": True,
            "use_diagnostics": True, "value_bootstrap": True, "use_smooth_clamp": False, "save_frequency": 100,
            "player": {"deterministic": True, "games_num": 100000, "use_vecenv": True},
        },
    }
}


def default_ppo_config(task: str = "hovering"):
    cfg = copy.deepcopy(_BASE)
    c = cfg["params"]["config"]
    c["env_name"], c["name"] = task, f"ppo_{task}"
    if task == "tracking":
        c["max_epochs"] = 300
    elif task == "balloon":  # ppo_balloon.yaml
        c.update(num_actors=64, horizon_length=32)
    elif task == "avoid":  # ppo_avoid.yaml:23-57
        cfg["params"]["network"]["cnn"] = {"output_dim": 30}
        c["env_config"]["use_image"] = True
        c.update(horizon_length=64, max_epochs=20000)
    elif task == "planning":  # ppo_planning.yaml:31-66
        cfg["params"]["network"]["cnn"] = {"output_dim": 30}
        c["env_config"]["use_image"] = True
        c.update(max_epochs=1000, save_frequency=20)
    return cfg


def scale_minibatch(config, num_actors):
    """Keep the reference's 48 minibatches per mini-epoch (98 304 / 2048) when num_actors changes."""
    c = config["params"]["config"]
    ratio = (c["num_actors"] * c["horizon_length"]) // c["minibatch_size"]
    c["num_actors"] = num_actors
    c["minibatch_size"] = max(1, (num_actors * c["horizon_length"]) // max(ratio, 1))
    return config
