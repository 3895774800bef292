"""GPU soft sync detector with the call surface of `wavecapsdr.decoders.p25_framer.P25P1SoftSyncDetector`
(decoders/p25_framer.py:125-231): sliding 24-symbol correlation against the P25 sync word 0x5575F5FF77FF.
`process_batch` is one launch (csrc/p25.cu `wc_c4fm_sync_scores`); the message assembler / NID BCH / TSBK trellis
above it stay the reference's Python."""
from __future__ import annotations

import numpy as np

from ..dsp.p25.c4fm import _SoftSyncDetector


class P25P1SoftSyncDetector:
    SYNC_PATTERN = 0x5575F5FF77FF

    def __init__(self) -> None:
        self._det = _SoftSyncDetector()
        self.SYNC_PATTERN_SYMBOLS = np.array(
            [3.0 if ((self.SYNC_PATTERN >> ((23 - i) * 2)) & 3) == 1 else -3.0 for i in range(24)], dtype=np.float32)

    def reset(self) -> None:
        self._det.reset()

    def process(self, soft_symbol: float) -> float:
        return self._det.process(soft_symbol)

    def process_batch(self, soft_symbols) -> np.ndarray:
        s = np.asarray(soft_symbols, dtype=np.float32)
        if s.size == 0:
            return np.array([], dtype=np.float32)
        return self._det.process_block(s).astype(np.float32)
