#!/usr/bin/env python
"""The five timings of the reference's backend/benchmark_dsp.py (:17-173), run against this package.

Same component names, workloads and result keys, so the two outputs can be laid side by side. The per-call
helper loops (`_Interpolator.filter`, `_SoftSyncDetector.process` one value at a time) are reproduced as the
reference times them AND in the batched form a GPU is meant for.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wavecap_sdr_b200.dsp.channelizer import PolyphaseChannelizer  # noqa: E402
from wavecap_sdr_b200.dsp.p25.c4fm import C4FMDemodulator, _FMDemodulator, _Interpolator, _SoftSyncDetector  # noqa: E402


def timed(fn, iterations):
    t0 = time.perf_counter()
    for _ in range(iterations):
        fn()
    return time.perf_counter() - t0


def main():
    rng = np.random.default_rng(0)
    out = []
    n = 50000
    i = (rng.standard_normal(n) * 0.5).astype(np.float32)
    q = (rng.standard_normal(n) * 0.5).astype(np.float32)
    d = _FMDemodulator(symbol_delay=10)            # the keyword benchmark_dsp.py:27 passes
    d.demodulate(i[:1000], q[:1000])

    def fm():
        d.reset()
        d.demodulate(i, q)
    e = timed(fm, 10)
    out.append(("FM Demodulator", "samples_per_sec", n * 10 / e))

    s = rng.standard_normal(10010).astype(np.float32)
    it = _Interpolator()
    offs = np.arange(10000) % (len(s) - 8)
    it.filter(s, 3, 0.5)
    e = timed(lambda: [it.filter(s, int(o), 0.5) for o in offs[:200]], 1)
    out.append(("8-tap Interpolator (per call, as the reference loops)", "symbols_per_sec", 200 / e))
    e = timed(lambda: it.filter_batch(s, offs, np.full(10000, 0.5)), 10)
    out.append(("8-tap Interpolator (batched)", "symbols_per_sec", 10000 * 10 / e))

    sym = (rng.standard_normal(10000) * 3).astype(np.float32)
    det = _SoftSyncDetector()
    det.process(0.0)
    e = timed(lambda: [det.process(float(v)) for v in sym[:200]], 1)
    out.append(("Sync Detector (per call)", "symbols_per_sec", 200 / e))
    e = timed(lambda: det.process_block(sym), 10)
    out.append(("Sync Detector (batched)", "symbols_per_sec", 10000 * 10 / e))

    fs = 8_000_000
    x = ((rng.standard_normal(fs) + 1j * rng.standard_normal(fs)) * 0.5).astype(np.complex64)
    ch = PolyphaseChannelizer(sample_rate=fs)       # 25 kHz grid -> 320 channels (generic path)
    ch.process_array(x[:100000])

    def chan():
        ch.reset()
        ch.process_array(x)
    e = timed(chan, 3)
    out.append(("Polyphase Channelizer (8 MS/s, 320 ch, host buffers)", "samples_per_sec", fs * 3 / e))

    iq = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.5).astype(np.complex64)
    dm = C4FMDemodulator(sample_rate=50000)
    dm.demodulate(iq[:10000])

    def full():
        dm.reset()
        dm.demodulate(iq)
    e = timed(full, 5)
    out.append(("Full C4FM Demodulator", "samples_per_sec", n * 5 / e))
    for name, key, v in out:
        print(f"{name:58s} {key:16s} {v:14.0f}")


if __name__ == "__main__":
    main()
