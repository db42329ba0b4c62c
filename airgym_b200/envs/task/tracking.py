"""Tracking — host-side mirror of airgym/envs/task/tracking.py (lemniscate reference, 48-dim obs)."""
from ..base.hovering import Hovering


class Tracking(Hovering):
    TASK = "tracking"
    REWARD_KEYS = (  # tracking.py:283-292
        "dist_norm", "dist_reward", "yaw_reward", "spin_reward", "continous_action_reward", "thrust_reward",
        "effort_reward", "ups_reward", "reward",
    )
