"""Drop-in alias: ``import umbrella...`` resolves to ``umbrella_b200...`` (SURVEY.md 8b -- the reference's
front-ends import ``umbrella.speculation.auto_engine.AutoEngine`` etc.).  No code lives here."""
import importlib
import importlib.abc
import importlib.util
import sys

import umbrella_b200 as _impl


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith("umbrella."):
            return None
        real = "umbrella_b200." + fullname[len("umbrella."):]
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except ModuleNotFoundError:
            return None
        return importlib.util.spec_from_loader(fullname, self)

    def create_module(self, spec):
        real = importlib.import_module("umbrella_b200." + spec.name[len("umbrella."):])
        return real

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _AliasFinder())
__path__ = _impl.__path__
__version__ = getattr(_impl, "__version__", "0")
