"""Task registry contents — same task names as the reference (airgym/envs/__init__.py:5-87).
Names in PENDING raise on make_env."""
from ..utils.task_registry import task_registry
from .base.hovering import Hovering
from .base.hovering_config import HoveringCfg
from .task.avoid import Avoid
from .task.avoid_config import AvoidCfg
from .task.balloon import Balloon
from .task.balloon_config import BalloonCfg
from .task.planning import Planning
from .task.planning_config import PlanningCfg
from .task.tracking import Tracking
from .task.tracking_config import TrackingCfg

TASK_CONFIGS = [
    {"name": "hovering", "config_class": HoveringCfg, "task_class": Hovering},
    {"name": "tracking", "config_class": TrackingCfg, "task_class": Tracking},
    {"name": "balloon", "config_class": BalloonCfg, "task_class": Balloon},
    {"name": "avoid", "config_class": AvoidCfg, "task_class": Avoid},
    {"name": "planning", "config_class": PlanningCfg, "task_class": Planning},
]
PENDING = ("customized",)  # the reference's bare Customized base has an empty reward (customized.py:462-474): not a trainable task


def register_tasks():
    for c in TASK_CONFIGS:
        task_registry.register(c["name"], c["task_class"], c["config_class"]())


register_tasks()
