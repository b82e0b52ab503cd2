"""Importing this package registers every class the YAML configs can name: each submodule decorates its classes
with the registry of its plug-in point when it is imported."""

from importlib import import_module as _import

for _submodule in ("modules", "standard_transformer", "meshed_memory_transformer", "object_relation_transformer",
                   "camo_transformer", "dual_collaborative_transformer"):
    _import(f"{__name__}.{_submodule}")

from .camo_transformer import CamoTransformer  # noqa: E402,F401
from .dual_collaborative_transformer import DualCollaborativeTransformer  # noqa: E402,F401
from .meshed_memory_transformer import MeshedMemoryTransformer  # noqa: E402,F401
from .modules import *  # noqa: E402,F401,F403
from .object_relation_transformer import ObjectRelationTransformer  # noqa: E402,F401
from .standard_transformer import StandardTransformerUsingGrid, StandardTransformerUsingRegion  # noqa: E402,F401
