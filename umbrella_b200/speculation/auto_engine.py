"""AutoEngine: ``engine`` key of a config -> engine object (the call the reference's front-ends make:
``AutoEngine.from_config(device, **json_config)``, umbrella/speculation/auto_engine.py:11-22).

Error behaviour kept from the reference: an unknown engine name is a ``ValueError`` listing the known ones;
missing ``model`` / ``draft_model`` keys are assertion failures; the default engine is the dynamic one.
"""
from __future__ import annotations

from importlib import import_module

# engine name -> (module inside this package, class); imported on first use
_REGISTRY = {
    "static": ("static_speculation_engine", "StaticSpeculationEngine"),
    "dynamic": ("dynamic_speculation_engine", "DynamicSpeculationEngine"),
}
# keys of configs/*.json that belong to the chat / API front-ends, not to an engine (app/chatbot.py:17-19)
_FRONT_END_KEYS = ("generation_length", "max_turns", "template")


def engine_class(name: str):
    if name not in _REGISTRY:
        raise ValueError(f"Engine type '{name}' is not supported. Supported types: {sorted(_REGISTRY)}")
    module, cls = _REGISTRY[name]
    return getattr(import_module(f"{__package__}.{module}"), cls)


class AutoEngine:
    @staticmethod
    def from_config(device: str, **config):
        cls = engine_class(config.pop("engine", "dynamic"))
        names = {role: config.pop(key, None) for role, key in (("draft_model_name", "draft_model"),
                                                                ("target_model_name", "model"))}
        for role, value in names.items():
            assert value is not None, f"config lacks the {role.replace('_name', '').replace('_', ' ')}"
        for key in _FRONT_END_KEYS:
            config.pop(key, None)
        return cls(device=device, **names, **config)
