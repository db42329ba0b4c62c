"""`lib` — the reference's top-level PPO package (lib/torch_runner.py, lib/agent, lib/core, lib/model, lib/network, lib/utils)
as an alias of `airgym_b200.lib`: `from lib.torch_runner import Runner` resolves to the B200 trainer."""
from airgym_b200._alias import install as _install

_install("lib", "airgym_b200.lib")
