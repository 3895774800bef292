"""Import helper for the live reference (build container only; /root/reference is absent on the
GPU box). SURVEY.md §8c recipe: trunking must be imported before capture (circular import)."""
from __future__ import annotations

import logging
import os
import sys

REFERENCE_BACKEND = "/root/reference/backend"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_BACKEND, "wavecapsdr"))


def load():
    """Make `wavecapsdr` importable and silence its INFO diagnostics."""
    if not available():
        raise RuntimeError("reference not present")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    sys.dont_write_bytecode = True
    if REFERENCE_BACKEND not in sys.path:
        sys.path.insert(0, REFERENCE_BACKEND)
    logging.disable(logging.CRITICAL)
    import wavecapsdr.trunking  # noqa: F401  (must precede wavecapsdr.capture)
    import wavecapsdr.capture  # noqa: F401
    return sys.modules["wavecapsdr"]
