"""CamoTransformer (reference: models/camo_transformer.py:10-45): region features -> FeatureEmbedding ->
CrossAttentionMultiLevelEncoder -> Decoder.  The whole-path engine does not cover this encoder; ``beam_search`` runs on
the registered modules (the module-level CUDA path)."""

from __future__ import annotations

from ..builders.model_builder import META_ARCHITECTURE
from .standard_transformer import _SingleStreamTransformer


@META_ARCHITECTURE.register()
class CamoTransformer(_SingleStreamTransformer):
    feature_field = "region_features"
