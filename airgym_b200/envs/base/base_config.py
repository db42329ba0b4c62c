"""Nested-class configs, instantiated recursively — same contract as the reference's BaseConfig
(airgym/envs/base/base_config.py:33-54): after construction every inner class attribute is an instance."""
import inspect


class BaseConfig:
    def __init__(self) -> None:
        _instantiate_members(self)


def _instantiate_members(obj):
    for key in dir(obj):
        if key == "__class__":
            continue
        member = getattr(obj, key)
        if inspect.isclass(member):
            inst = member()
            setattr(obj, key, inst)
            _instantiate_members(inst)
