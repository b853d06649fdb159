"""Plugin registry mirroring ofasys/configure/config_store.py:22-131 (register_config / ConfigStore)
and the BaseDataclass of ofasys/configure/configs.py:34-94 -- only what the hot path's classes use."""
import dataclasses
from dataclasses import dataclass
from typing import Callable, Dict, Optional


@dataclass
class BaseDataclass:
    _name: Optional[str] = None

    @classmethod
    def from_namespace(cls, args):
        if isinstance(args, cls):
            return args
        cfg = cls()
        for k in cfg.__dataclass_fields__:
            if not k.startswith("_") and hasattr(args, k):
                setattr(cfg, k, getattr(args, k))
        return cfg


@dataclass
class ConfigNode:
    target: object
    config: Optional[object] = None
    is_active: bool = False


class ConfigStore:
    """Singleton registry keyed `<group>.<name>` (groups: ofasys.model, ofasys.adaptor, ...)."""

    _inst = None

    def __new__(cls):
        if cls._inst is None:
            cls._inst = super().__new__(cls)
            cls._inst.repo = {}
        return cls._inst

    def store(self, group: str, name: str, obj: type, dc: Optional[type] = None) -> None:
        assert group and name
        assert dc is None or dataclasses.is_dataclass(dc)
        self.repo[f"{group}.{name}"] = ConfigNode(obj, dc() if dc is not None else None)

    def get(self, group: str, name: Optional[str] = None):
        if name is None:
            return [n for p, n in self.repo.items() if p.rsplit(".", 1)[0] == group and n.is_active]
        return self.repo[f"{group}.{name}"]

    def contain(self, group: str, name: str) -> bool:
        return f"{group}.{name}" in self.repo

    def names(self, group: str):
        return [p.rsplit(".", 1)[1] for p in self.repo if p.rsplit(".", 1)[0] == group]


def register_config(group: str, name: str, dataclass: Optional[type] = None) -> Callable:
    def _register(cls):
        ConfigStore().store(group, name, cls, dataclass)
        return cls

    return _register
