"""airgym/utils/__init__.py of the reference exports these three names."""
from .helpers import class_to_dict, get_args  # noqa: F401
from .task_registry import task_registry  # noqa: F401
