"""Extract constant data tables shared by the oracle and the CUDA kernels from the reference.

`interp_taps_129x8.npy`: the 8-tap, 128-step MMSE fractional interpolator coefficient table of
`wavecapsdr.dsp.p25.c4fm._Interpolator.TAPS` (c4fm.py:907-2202) — the published GNU Radio
`interpolator_taps.h` table that SDRTrunk's Interpolator.java also carries. It is numeric data with no
closed form (the output of an MMSE optimisation), so bit-exact dibit parity needs the same numbers.
Run in the build container only:  python oracle/make_tables.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refenv  # noqa: E402

if __name__ == "__main__":
    refenv.load()
    from wavecapsdr.dsp.p25.c4fm import _Interpolator

    t = np.ascontiguousarray(_Interpolator.TAPS, dtype=np.float32)
    assert t.shape == (129, 8)
    dst = os.path.join(ROOT, "wavecap-sdr_b200", "dsp", "p25", "interp_taps_129x8.npy")
    np.save(dst, t)
    print("wrote", dst, t.shape, float(t.sum()))
