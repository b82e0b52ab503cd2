"""Importing this package registers every class the YAML configs can name."""

from .modules import *  # noqa: F401,F403
from .standard_transformer import StandardTransformerUsingGrid, StandardTransformerUsingRegion  # noqa: F401
from .meshed_memory_transformer import MeshedMemoryTransformer  # noqa: F401
from .object_relation_transformer import ObjectRelationTransformer  # noqa: F401
