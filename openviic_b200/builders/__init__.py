"""Registries and ``build_*`` functions selected by the YAML ``ARCHITECTURE`` strings.

Importing ``openviic_b200.models`` fills the registries (the reference gets the same effect from the
star-imports in its builders/__init__.py:1-2).
"""

from .registry import Registry  # noqa: F401
from .attention_builder import META_ATTENTION, build_attention  # noqa: F401
from .encoder_builder import META_ENCODER, build_encoder  # noqa: F401
from .decoder_builder import META_DECODER, build_decoder  # noqa: F401
from .vision_embedding_builder import META_VISION_EMBEDDING, build_vision_embedding  # noqa: F401
from .text_embedding_builder import META_TEXT_EMBEDDING, build_text_embedding  # noqa: F401
from .model_builder import META_ARCHITECTURE, build_model  # noqa: F401
