"""Object-relation transformer (reference: models/object_relation_transformer.py:10-43).

The reference's ``encoder_forward`` wraps its tensors in an ``Instance`` that ``GeometricEncoder``
cannot unpack (TypeError as shipped); this class passes features, boxes and
mask as the encoder's signature asks -- the same one-line fix the oracle documents.
"""

from __future__ import annotations

from ..builders.model_builder import META_ARCHITECTURE
from .standard_transformer import _SingleStreamTransformer


@META_ARCHITECTURE.register()
class ObjectRelationTransformer(_SingleStreamTransformer):
    feature_field = "region_features"

    def engine_inputs(self, input_features):
        return input_features.region_features, input_features.region_boxes

    def encoder_forward(self, input_features):
        region_features, region_padding_mask = self.vision_embedding(input_features.region_features)
        encoder_features = self.encoder(features=region_features, boxes=input_features.region_boxes,
                                        padding_mask=region_padding_mask)
        return encoder_features, region_padding_mask
