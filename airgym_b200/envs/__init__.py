"""Task registry contents — same task names as the reference (airgym/envs/__init__.py:5-87).
Tasks whose kernels are not built yet are listed in PENDING and raise on make_env."""
from ..utils.task_registry import task_registry
from .base.hovering import Hovering
from .base.hovering_config import HoveringCfg
from .task.balloon import Balloon
from .task.balloon_config import BalloonCfg
from .task.tracking import Tracking
from .task.tracking_config import TrackingCfg

TASK_CONFIGS = [
    {"name": "hovering", "config_class": HoveringCfg, "task_class": Hovering},
    {"name": "tracking", "config_class": TrackingCfg, "task_class": Tracking},
    {"name": "balloon", "config_class": BalloonCfg, "task_class": Balloon},
]
PENDING = ("customized", "avoid", "planning")


def register_tasks():
    for c in TASK_CONFIGS:
        task_registry.register(c["name"], c["task_class"], c["config_class"]())


register_tasks()
