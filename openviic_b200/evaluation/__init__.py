"""Caption metrics on the caller side of the path (reference: evaluation/__init__.py).  Only CIDEr is built: it is
the self-critical reward (trainers/vi_trainer.py:137-145) and the score the evaluation loop selects models by;
BLEU / METEOR / ROUGE need the Java tokenizer and METEOR jars the reference shells out to."""

from .cider import Cider  # noqa: F401
