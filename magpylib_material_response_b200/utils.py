"""Per-stage wall-clock timer mirroring the reference's ``timelog`` hook points
(/root/reference/src/magpylib_material_response/utils.py:79-107; wrapped stages listed in
SURVEY.md section 5).  The watcher thread / loguru sinks are out of scope: stages are timed,
kept in ``LAST_TIMINGS`` and logged through loguru only when it is importable and enabled.
"""

from __future__ import annotations

import time
from contextlib import contextmanager

try:
    from loguru import logger

    logger.disable("magpylib_material_response_b200")
except ImportError:  # pragma: no cover
    logger = None

DEFAULT_MIN_LOG_TIME = 1.0
LAST_TIMINGS: dict[str, float] = {}


@contextmanager
def timelog(msg, min_log_time=None):
    """Time a block; record it under ``msg`` in LAST_TIMINGS; re-raise failures."""
    if min_log_time is None:
        min_log_time = DEFAULT_MIN_LOG_TIME
    start = time.perf_counter()
    try:
        yield
    except Exception:
        if logger is not None:
            logger.opt(depth=2).error("Failed: {msg}", msg=msg)
        raise
    finally:
        LAST_TIMINGS[msg] = time.perf_counter() - start
    elapsed = LAST_TIMINGS[msg]
    if logger is not None and elapsed > min_log_time:
        logger.opt(depth=2).info("Completed: {msg} in {t:.3f} s", msg=msg, t=elapsed)
