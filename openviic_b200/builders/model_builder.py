"""Reference import path ``builders.model_builder``; defined in ``builders/__init__.py``."""

from . import META_ARCHITECTURE, build_model  # noqa: F401
