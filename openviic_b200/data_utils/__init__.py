"""Host glue either side of the caption path: vocabulary (text <-> ids) and
the collate that turns per-image feature rows into the pinned batch the engine's H2D copy reads."""

from .dataset import DictionaryDataset, FeatureDataset  # noqa: F401
from .utils import FeatureBatcher, collate_fn, get_tokenizer, preprocess_caption  # noqa: F401
from .vocab import Vocab  # noqa: F401
