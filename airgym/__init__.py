"""`airgym` — the reference's package name (airgym/__init__.py, airgym/envs, airgym/utils) as an alias of `airgym_b200`:
`from airgym.envs import task_registry`, `from airgym.utils.helpers import get_args`, `from airgym.envs.base.hovering import
Hovering` resolve to the B200 implementation."""
from airgym_b200._alias import install as _install

_install("airgym", "airgym_b200")
from airgym_b200 import *  # noqa: E402,F401,F403
