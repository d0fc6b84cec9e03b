"""Minimal `type=`-string registries with the reference's names (pointcept/utils/registry.py:212-316,
pointcept/models/builder.py:10-16, pointcept/models/losses/builder.py:10) so a SceneSplat config dict
builds this package's classes unchanged, plus `register_into_pointcept()` which installs them into a
live Pointcept's registries (force=True)."""
from __future__ import annotations


class Registry:
    def __init__(self, name):
        self.name = name
        self._module_dict = {}

    def get(self, key):
        return self._module_dict.get(key)

    def _register(self, cls, name=None, force=False):
        name = name or cls.__name__
        if not force and name in self._module_dict:
            raise KeyError(f"{name} is already registered in {self.name}")
        self._module_dict[name] = cls

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module

        def deco(cls):
            self._register(cls, name, force)
            return cls

        return deco

    def build(self, cfg):
        cfg = dict(cfg)
        typ = cfg.pop("type")
        cls = self.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        return cls(**cfg)


MODELS = Registry("models")
MODULES = Registry("modules")
LOSSES = Registry("losses")
TRANSFORMS = Registry("transforms")


def build_model(cfg):
    return MODELS.build(cfg)


def register_into_pointcept():
    """Replace the reference's registrations with this package's classes (same names)."""
    from pointcept.datasets.transform import TRANSFORMS as P_TRANSFORMS
    from pointcept.models.builder import MODELS as P_MODELS
    from pointcept.models.losses.builder import LOSSES as P_LOSSES

    for src, dst in ((MODELS, P_MODELS), (LOSSES, P_LOSSES), (TRANSFORMS, P_TRANSFORMS)):
        for name, cls in src._module_dict.items():
            dst.register_module(name=name, force=True, module=cls)
