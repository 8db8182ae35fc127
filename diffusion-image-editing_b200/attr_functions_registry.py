"""Name -> strategy registry (drop-in for src/attr_functions_registry.py)."""
from typing import Any, Dict, Optional, Type, Union

from attr_functions import (AnyGANAttrFunc, AttrFunc, MultiColorAttrFunc, NetAttrFunc, SingleColorAttrFunc)

Strategy = Union[Type[AttrFunc], AttrFunc]


class AttrFuncRegistry:
    """Holds strategy classes (instantiated on ``get`` with ``params``) or ready instances."""

    def __init__(self) -> None:
        self._registry: Dict[str, Strategy] = {}

    def register(self, strategy: Strategy) -> None:
        key = strategy.__name__ if isinstance(strategy, type) else strategy.name
        self._registry[key] = strategy

    def get(self, name: str, params: Optional[Dict[str, Any]] = None) -> AttrFunc:
        entry = self._registry.get(name)
        if entry is None:
            raise ValueError(f"No strategy registered with name: {name}")
        if isinstance(entry, type):
            return entry(**params) if params else entry()
        return entry

    def get_attribute_functions(self) -> list:
        return list(self._registry.keys())


def create_attr_func_registry() -> AttrFuncRegistry:
    reg = AttrFuncRegistry()
    for cls in (SingleColorAttrFunc, MultiColorAttrFunc, NetAttrFunc, AnyGANAttrFunc):
        reg.register(cls)
    return reg
