"""Reference import path ``builders.text_embedding_builder``; defined in ``builders/__init__.py``."""

from . import META_TEXT_EMBEDDING, build_text_embedding  # noqa: F401
