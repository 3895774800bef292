"""GPU AGC with the call surface of `wavecapsdr.dsp.agc` (dsp/agc.py)."""
from __future__ import annotations

import numpy as np

from . import _stages as S
from . import filters as F

SCIPY_AVAILABLE = True   # reference flags (dsp/agc.py:23-47); the GPU path always follows the
NUMBA_AVAILABLE = False  # scipy (lfilter) formulation, which is what the reference picks first


def _n(x) -> int:
    return int(x.numel()) if hasattr(x, "numel") else int(np.asarray(x).size)


def soft_clip(x):
    """tanh(1.5 x)/tanh(1.5), no head-room factor (dsp/agc.py:58-70)."""
    if _n(x) == 0:
        return F._f32_passthrough(x)
    t = S.to_device(x, np.float32)
    return S.like_input(S.elementwise(t, S.OP_SOFT_CLIP_AGC), x)


def agc_params(sample_rate: int, target_db: float = -20.0, attack_ms: float = 5.0, release_ms: float = 50.0,
               max_gain_db: float = 60.0):
    """Coefficients of dsp/agc.py:205-215 and the float32 one-pole filters of :91-98."""
    target = 10.0 ** (target_db / 20.0)
    max_gain = 10.0 ** (max_gain_db / 20.0)
    att_n = (attack_ms / 1000.0) * sample_rate
    rel_n = (release_ms / 1000.0) * sample_rate
    att = 1.0 - np.exp(-1.0 / att_n) if att_n > 0 else 1.0
    rel = 1.0 - np.exp(-1.0 / rel_n) if rel_n > 0 else 1.0

    def one_pole(c):
        b = np.array([c], dtype=np.float32)
        a = np.array([1.0, -(1.0 - c)], dtype=np.float32)
        return tuple(float(v) for v in b), tuple(float(v) for v in a)

    return one_pole(att), one_pole(rel), float(target), float(max_gain)


def agc_rows(rows, sample_rate: int, target_db: float = -20.0, attack_ms: float = 5.0, release_ms: float = 50.0,
             max_gain_db: float = 60.0):
    """apply_agc on CUDA rows [n_seq, n]: |x| -> attack one-pole -> release one-pole -> max -> gain -> tanh."""
    (ba, aa), (br, ar), target, max_gain = agc_params(sample_rate, target_db, attack_ms, release_ms, max_gain_db)
    env_a = S.lfilter(ba, aa, rows, abs_input=True)
    env_r = S.lfilter(br, ar, env_a)
    return S.agc_apply(rows, env_a, env_r, target, max_gain)


def apply_agc(x, sample_rate: int, target_db: float = -20.0, attack_ms: float = 5.0, release_ms: float = 50.0,
              max_gain_db: float = 60.0):
    """Attack/release AGC (dsp/agc.py:169-242, lfilter formulation :73-108)."""
    if _n(x) == 0:
        return F._f32_passthrough(x)
    rows = S.to_device(x, np.float32).reshape(1, -1)
    y = agc_rows(rows, sample_rate, target_db, attack_ms, release_ms, max_gain_db)
    return S.like_input(y.reshape(-1), x)


def apply_simple_agc(x, target_rms: float = 0.1, max_gain: float = 10.0):
    """Block RMS AGC (dsp/agc.py:245-285)."""
    if _n(x) == 0:
        return F._f32_passthrough(x)
    t = S.to_device(x, np.float32)
    rows = t.reshape(1, -1)
    rms = float(np.float32(np.sqrt(float(S.sumsq(rows)[0].item()) / rows.shape[1])))
    gain = min(target_rms / rms if rms > 1e-6 else max_gain, max_gain)
    y = S.elementwise(S.elementwise(t, S.OP_SCALE, float(np.float32(gain))), S.OP_SOFT_CLIP_AGC)
    return S.like_input(y, x)
