"""Name -> class registry: the plug-in API the caption path sits behind.

Same contract as the reference's builders/registry.py:8-90: ``register()`` works as a decorator or
a call, registering under ``obj.__name__``; duplicate names raise AssertionError; ``get`` raises
KeyError for unknown names; iteration yields (name, object) pairs.
"""

from __future__ import annotations

from typing import Any, Dict, Iterator, Optional, Tuple


class Registry:
    def __init__(self, name: str) -> None:
        self._name = name
        self._obj_map: Dict[str, Any] = {}

    def _do_register(self, name: str, obj: Any) -> None:
        assert name not in self._obj_map, (
            "An object named '{}' was already registered in '{}' registry!".format(name, self._name))
        self._obj_map[name] = obj

    def register(self, obj: Optional[Any] = None) -> Any:
        if obj is None:
            def deco(func_or_class: Any) -> Any:
                self._do_register(func_or_class.__name__, func_or_class)
                return func_or_class
            return deco
        self._do_register(obj.__name__, obj)
        return None

    def get(self, name: str) -> Any:
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError("No object named '{}' found in '{}' registry!".format(name, self._name))
        return ret

    def __contains__(self, name: str) -> bool:
        return name in self._obj_map

    def __iter__(self) -> Iterator[Tuple[str, Any]]:
        return iter(self._obj_map.items())

    def __len__(self) -> int:
        return len(self._obj_map)

    def __repr__(self) -> str:
        rows = "\n".join("  {:<48} {}".format(k, v) for k, v in self._obj_map.items())
        return "Registry of {}:\n{}".format(self._name, rows)

    __str__ = __repr__
