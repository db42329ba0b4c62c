"""Import aliases so that code written against the reference's module paths runs unchanged:
`airgym.*` → `airgym_b200.*` (airgym/envs, airgym/utils of the reference) and `lib.*` → `airgym_b200.lib.*` (the reference's
top-level `lib/` package: lib.torch_runner, lib.agent.*, lib.core.*, lib.network.*).  The alias modules ARE the real modules
(`sys.modules['airgym.envs'] is sys.modules['airgym_b200.envs']`), so registries and class identities are shared."""
import importlib
import importlib.abc
import importlib.util
import sys


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, prefix, target):
        self.prefix, self.target = prefix, target

    def _real(self, fullname):
        return self.target + fullname[len(self.prefix):]

    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(self.prefix + "."):
            return None
        try:
            real = importlib.util.find_spec(self._real(fullname))
        except (ImportError, ValueError):
            return None
        if real is None:
            return None
        return importlib.util.spec_from_loader(fullname, self, is_package=real.submodule_search_locations is not None)

    def create_module(self, spec):
        return importlib.import_module(self._real(spec.name))

    def exec_module(self, module):
        pass


def install(prefix, target):
    if not any(isinstance(f, _AliasFinder) and f.prefix == prefix for f in sys.meta_path):
        sys.meta_path.insert(0, _AliasFinder(prefix, target))
