"""Meshed-memory transformer (reference: models/meshed_memory_transformer.py:9-40): memory-augmented
multi-level encoder + meshed decoder, selected entirely by the YAML's ENCODER/DECODER nodes."""

from __future__ import annotations

from ..builders.model_builder import META_ARCHITECTURE
from .standard_transformer import _SingleStreamTransformer


@META_ARCHITECTURE.register()
class MeshedMemoryTransformer(_SingleStreamTransformer):
    feature_field = "region_features"
