"""Reference import path ``builders.vision_embedding_builder``; defined in ``builders/__init__.py``."""

from . import META_VISION_EMBEDDING, build_vision_embedding  # noqa: F401
