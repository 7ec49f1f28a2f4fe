"""Stand-in for the two `river.stats` accumulators the reference uses (river is not in the offline
wheelhouse).  Ours, not reference code: mkb/evaluation/evaluation.py:187-199 and
mkb/compose/pipeline.py:189,242-244 only call ``update(x)`` and ``get()``."""
from . import stats  # noqa: F401
