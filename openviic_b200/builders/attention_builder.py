"""Reference import path ``builders.attention_builder``; defined in ``builders/__init__.py``."""

from . import META_ATTENTION, build_attention  # noqa: F401
