"""The reference's import paths resolve to the B200 implementation (north_star: "keeps AirGym's task/env API surface — airgym.envs
registry ... lib.torch_runner PPO entry point"): scripts/runner.py:14-16 and airgym/utils/__init__.py of the reference."""
import sys


def test_reference_import_paths_are_aliases_of_the_package(built):
    from airgym.envs import task_registry
    from airgym.utils import class_to_dict, get_args, task_registry as tr2
    from airgym.utils.helpers import get_args as ga2
    from lib.torch_runner import Runner
    from lib.utils.isaacgym_utils import RLGPUAlgoObserver

    import airgym_b200.envs
    import airgym_b200.lib.torch_runner
    from airgym.envs.base.hovering import Hovering
    from airgym.envs.base.hovering_config import HoveringCfg
    from airgym.envs.task.planning import Planning
    from lib.agent.a2c_continuous import A2CAgent
    from lib.core.running_mean_std import RunningMeanStd

    assert sys.modules["airgym.envs"] is airgym_b200.envs and task_registry is tr2 and get_args is ga2
    assert Runner is airgym_b200.lib.torch_runner.Runner
    assert Hovering is airgym_b200.envs.base.hovering.Hovering and issubclass(Planning, Hovering)
    assert set(task_registry.get_registered_tasks()) >= {"hovering", "tracking", "balloon", "avoid", "planning"}
    assert callable(class_to_dict) and HoveringCfg.env.num_observations == 18 and A2CAgent and RunningMeanStd
    obs = RLGPUAlgoObserver()
    obs.process_infos({"x": 1.0, "item_reward_info": {}}, None)
    assert obs.direct_info == {"x": 1.0}
    try:
        import airgym.does_not_exist  # noqa: F401
    except ImportError:
        pass
    else:
        raise AssertionError("unknown submodules must not resolve")
