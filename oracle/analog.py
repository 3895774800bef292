"""Oracle: analog demod chain (capture.py:166-193, 298-439; dsp/fm.py; dsp/am.py; dsp/agc.py;
dsp/filters.py:41-264). numpy/scipy restatement — test infrastructure only.

dtype discipline follows the reference as executed with numpy 2.x (NEP 50 weak scalars):
float32 arrays stay float32 under Python-float scalars; scipy.signal.lfilter computes in float64
unless every operand is float32 (de-emphasis and AGC envelope coefficients are float32 arrays, so
those two recursions run in float32); each filter result is cast back to float32.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from scipy import signal

F32 = np.float32


# ---- capture.py --------------------------------------------------------------------------------

def freq_shift(iq, offset_hz, sample_rate):
    """capture.py:166-193. Phase index n is float32, the scalar -j*2*pi*off/fs is a weak Python
    complex, so exp() is evaluated on a complex64 array: theta[n] = fl32(k32 * fl32(n))."""
    if offset_hz == 0.0 or iq.size == 0:
        return iq
    off = round(offset_hz)
    n = np.arange(iq.shape[0], dtype=F32)
    ph = np.exp(-1j * 2.0 * np.pi * (off / float(sample_rate)) * n).astype(np.complex64)
    return (iq.astype(np.complex64, copy=False) * ph).astype(np.complex64)


def rssi_db(base):
    """capture.py:331-334."""
    mag = np.abs(base)
    return float(10.0 * np.log10(np.mean(mag ** 2) + 1e-10))


# ---- dsp/fm.py ---------------------------------------------------------------------------------

def soft_clip_fm(x):
    """dsp/fm.py:20-39."""
    return np.tanh(x * F32(1.5)) * F32(1.0 / np.tanh(1.5)) * F32(0.95)


def rms_normalize(x, target_rms=0.18, min_rms=1e-4):
    """dsp/fm.py:42-62."""
    if x.size == 0:
        return x
    rms = float(np.sqrt(np.mean(x ** 2)))
    return x * (target_rms / rms) if rms > min_rms else x


def quadrature_demod(iq, sample_rate):
    """dsp/fm.py:65-97."""
    if iq.size == 0:
        return np.empty(0, dtype=F32)
    x = iq.astype(np.complex64, copy=False)
    out = np.empty(x.size, dtype=F32)
    out[0] = 0.0
    out[1:] = np.angle(x[1:] * np.conj(x[:-1])) * F32(sample_rate / (2.0 * np.pi * 75000.0))
    return out


def deemphasis_filter(x, sample_rate, tau=75e-6):
    """dsp/fm.py:101-126 — float32 b, a => float32 recursion."""
    tau_q = int(tau * 1e6) * 1e-6
    alpha = 1.0 / (1.0 + (1.0 / (2.0 * np.pi * tau_q * sample_rate)))
    b = np.array([alpha], dtype=F32)
    a = np.array([1.0, -(1.0 - alpha)], dtype=F32)
    return signal.lfilter(b, a, x).astype(F32)


def lpf_audio(x, sample_rate, cutoff=15_000):
    """dsp/fm.py:130-181 — butter(5) in (b, a) form, float64 lfilter."""
    if x.size == 0:
        return x.astype(F32, copy=False)
    wn = int(cutoff) / (sample_rate / 2.0)
    if wn >= 1.0:
        return x.astype(F32, copy=False)
    b, a = signal.butter(5, wn, btype="low")
    return signal.lfilter(b, a, x).astype(F32)


def resample_poly(x, in_rate, out_rate):
    """dsp/fm.py:184-221."""
    if x.size == 0 or in_rate == out_rate:
        return x.astype(F32, copy=False)
    g = math.gcd(int(in_rate), int(out_rate))
    return signal.resample_poly(x.astype(np.float64), out_rate // g, in_rate // g).astype(F32)


# ---- dsp/filters.py ----------------------------------------------------------------------------

def _butter_apply(x, btype, wn):
    b, a = signal.butter(5, wn, btype=btype)
    return signal.lfilter(b, a, x).astype(F32)


def highpass_filter(x, sample_rate, cutoff):
    """dsp/filters.py:85-126."""
    wn = cutoff / (sample_rate / 2.0)
    if x.size == 0 or wn <= 0 or wn >= 1.0:
        return x.astype(F32, copy=False)
    return _butter_apply(x, "high", wn)


def lowpass_filter(x, sample_rate, cutoff):
    """dsp/filters.py:129-172."""
    wn = cutoff / (sample_rate / 2.0)
    if x.size == 0 or wn <= 0 or wn >= 1.0:
        return x.astype(F32, copy=False)
    return _butter_apply(x, "low", wn)


def bandpass_filter(x, sample_rate, low, high):
    """dsp/filters.py:175-221."""
    lo, hi = low / (sample_rate / 2.0), high / (sample_rate / 2.0)
    if x.size == 0 or lo <= 0 or hi >= 1.0 or lo >= hi:
        return x.astype(F32, copy=False)
    return _butter_apply(x, "band", [lo, hi])


def notch_filter(x, sample_rate, freq, q=30.0):
    """dsp/filters.py:224-264."""
    w0 = freq / (sample_rate / 2.0)
    if x.size == 0 or w0 <= 0 or w0 >= 1.0:
        return x.astype(F32, copy=False)
    b, a = signal.iirnotch(w0, q)
    return signal.lfilter(b, a, x).astype(F32)


# ---- dsp/agc.py --------------------------------------------------------------------------------

def soft_clip_agc(x):
    """dsp/agc.py:58-70."""
    return np.tanh(x * F32(1.5)) * F32(1.0 / np.tanh(1.5))


def apply_agc(x, sample_rate, target_db=-20.0, attack_ms=5.0, release_ms=50.0, max_gain_db=60.0):
    """dsp/agc.py:169-242 with the lfilter envelope of :73-108."""
    if x.size == 0:
        return x.astype(F32, copy=False)
    target = 10.0 ** (target_db / 20.0)
    max_gain = 10.0 ** (max_gain_db / 20.0)
    att_n = (attack_ms / 1000.0) * sample_rate
    rel_n = (release_ms / 1000.0) * sample_rate
    att = 1.0 - np.exp(-1.0 / att_n) if att_n > 0 else 1.0
    rel = 1.0 - np.exp(-1.0 / rel_n) if rel_n > 0 else 1.0
    mag = np.abs(x).astype(F32)
    env_a = signal.lfilter(np.array([att], dtype=F32), np.array([1.0, -(1.0 - att)], dtype=F32), mag)
    env_r = signal.lfilter(np.array([rel], dtype=F32), np.array([1.0, -(1.0 - rel)], dtype=F32), env_a)
    env = np.maximum(env_a, env_r).astype(F32)
    gain = target / np.maximum(env, 1e-6)
    np.minimum(gain, max_gain, out=gain)
    return soft_clip_agc(x * gain).astype(F32)


# ---- demodulators ------------------------------------------------------------------------------

def wbfm_demod(iq, sample_rate, audio_rate=48_000, enable_deemphasis=True, deemphasis_tau=75e-6,
               enable_mpx_filter=True, mpx_cutoff_hz=15_000, enable_highpass=False, highpass_hz=100,
               notch_frequencies=None):
    """dsp/fm.py:228-314 (noise blanker / spectral NR flags off)."""
    fm = quadrature_demod(iq, sample_rate)
    if enable_deemphasis:
        fm = deemphasis_filter(fm, sample_rate, deemphasis_tau)
    if enable_mpx_filter:
        fm = lpf_audio(fm, sample_rate, mpx_cutoff_hz)
    if enable_highpass and highpass_hz > 0:
        fm = highpass_filter(fm, sample_rate, highpass_hz)
    for f in notch_frequencies or []:
        if 0 < f < sample_rate / 2:
            fm = notch_filter(fm, sample_rate, f)
    fm = rms_normalize(fm, 0.18)
    return soft_clip_fm(resample_poly(fm, sample_rate, audio_rate))


def nbfm_demod(iq, sample_rate, audio_rate=48_000, enable_deemphasis=False, deemphasis_tau=75e-6,
               enable_highpass=False, highpass_hz=300, enable_lowpass=False, lowpass_hz=3_000,
               notch_frequencies=None):
    """dsp/fm.py:317-406."""
    fm = quadrature_demod(iq, sample_rate)
    if enable_deemphasis:
        fm = deemphasis_filter(fm, sample_rate, deemphasis_tau)
    if enable_highpass and highpass_hz > 0:
        fm = highpass_filter(fm, sample_rate, highpass_hz)
    if enable_lowpass and lowpass_hz > 0:
        fm = lowpass_filter(fm, sample_rate, lowpass_hz)
    for f in notch_frequencies or []:
        if 0 < f < sample_rate / 2:
            fm = notch_filter(fm, sample_rate, f)
    fm = rms_normalize(fm, 0.18)
    return soft_clip_fm(resample_poly(fm, sample_rate, audio_rate))


def am_freq_shift(iq, offset_hz, sample_rate):
    """dsp/am.py:23-42 — float64 time base, + sign."""
    if iq.size == 0:
        return iq
    t = np.arange(iq.shape[0], dtype=np.float64) / float(sample_rate)
    return (iq * np.exp(2j * np.pi * offset_hz * t).astype(np.complex64)).astype(np.complex64)


def am_demod(iq, sample_rate, audio_rate=48_000, enable_agc=True, enable_highpass=True, highpass_hz=100,
             enable_lowpass=True, lowpass_hz=5000, agc_target_db=-20.0, notch_frequencies=None,
             enable_noise_blanker=False, noise_blanker_threshold_db=10.0):
    """dsp/am.py:45-141."""
    if iq.size == 0:
        return np.empty(0, dtype=F32)
    audio = np.abs(iq).astype(F32)
    if enable_noise_blanker:                                  # dsp/am.py:100-101
        audio = noise_blanker(audio, noise_blanker_threshold_db, 3)
    if enable_highpass and highpass_hz > 0:
        audio = highpass_filter(audio, sample_rate, highpass_hz)
    if enable_lowpass and lowpass_hz > 0:
        audio = lowpass_filter(audio, sample_rate, lowpass_hz)
    for f in notch_frequencies or []:
        if 0 < f < sample_rate / 2:
            audio = notch_filter(audio, sample_rate, f)
    if enable_agc:
        audio = apply_agc(audio, sample_rate, target_db=agc_target_db, attack_ms=5.0, release_ms=50.0)
    audio = resample_poly(audio, sample_rate, audio_rate)
    return audio if enable_agc else soft_clip_agc(audio)


def ssb_demod(iq, sample_rate, audio_rate=48_000, mode="usb", enable_agc=True, enable_bandpass=True,
              bandpass_low=300, bandpass_high=3000, agc_target_db=-20.0, notch_frequencies=None,
              bfo_offset_hz=1500.0, enable_noise_blanker=False, noise_blanker_threshold_db=10.0):
    """dsp/am.py:144-247."""
    if iq.size == 0:
        return np.empty(0, dtype=F32)
    shifted = am_freq_shift(iq, bfo_offset_hz if mode.lower() == "usb" else -bfo_offset_hz, sample_rate)
    audio = np.real(shifted).astype(F32)
    if enable_noise_blanker:                                  # dsp/am.py:213-215
        audio = noise_blanker(audio, noise_blanker_threshold_db, 3)
    if enable_bandpass:
        audio = bandpass_filter(audio, sample_rate, bandpass_low, bandpass_high)
    for f in notch_frequencies or []:
        if 0 < f < sample_rate / 2:
            audio = notch_filter(audio, sample_rate, f)
    if enable_agc:
        audio = apply_agc(audio, sample_rate, target_db=agc_target_db, attack_ms=5.0, release_ms=50.0)
    audio = resample_poly(audio, sample_rate, audio_rate)
    return audio if enable_agc else soft_clip_agc(audio)


# ---- dsp/sam.py ----------------------------------------------------------------------------------

class CarrierRecoveryPLLOracle:
    """dsp/sam.py:25-129: type-2 PLL, proportional + integral loop filter, one Python-float (float64) update per sample."""

    def __init__(self, sample_rate, loop_bandwidth=50.0, damping=0.707):
        self.sample_rate, self.loop_bandwidth, self.damping = float(sample_rate), loop_bandwidth, damping
        omega_n = 2 * np.pi * loop_bandwidth                               # :63-65
        self.alpha = 2 * damping * omega_n / self.sample_rate
        self.beta = (omega_n ** 2) / (self.sample_rate ** 2)
        self.phase = self.frequency = self.integrator = 0.0

    def process(self, iq):
        n = len(iq)
        ci, cq = np.zeros(n, dtype=F32), np.zeros(n, dtype=F32)
        phase, integ, freq, alpha, beta = self.phase, self.integrator, self.frequency, self.alpha, self.beta
        x = np.asarray(iq, dtype=np.complex128)                            # complex64 scalar * complex128 lo -> complex128 (:96)
        for i in range(n):
            lo = complex(math.cos(phase), -math.sin(phase))                # np.exp(-1j * phase) (:93)
            mixed = complex(x[i]) * lo
            ci[i], cq[i] = mixed.real, mixed.imag
            pe = math.atan2(mixed.imag, abs(mixed.real) + 1e-10)           # :103
            integ += beta * pe                                             # :106-107
            freq = alpha * pe + integ
            phase += freq
            if phase > math.pi:                                            # :114-117
                phase -= 2 * math.pi
            elif phase < -math.pi:
                phase += 2 * math.pi
        self.phase, self.integrator, self.frequency = phase, integ, freq
        return ci, cq, freq * self.sample_rate / (2 * np.pi)


def sam_demod(iq, sample_rate, audio_rate=48_000, sideband="dsb", pll_bandwidth=50.0, pll_damping=0.707, enable_agc=True,
              enable_highpass=True, highpass_hz=100.0, enable_lowpass=True, lowpass_hz=5000.0, enable_noise_blanker=False,
              noise_blanker_threshold_db=10.0, agc_target_db=-20.0, notch_frequencies=None, pll_state=None):
    """dsp/sam.py:132-270 -> (audio, carrier offset Hz, pll)."""
    if iq.size == 0:
        return np.empty(0, dtype=F32), 0.0, pll_state
    pll = pll_state or CarrierRecoveryPLLOracle(sample_rate, pll_bandwidth, pll_damping)
    ci, cq, f_off = pll.process(iq)
    sb = sideband.lower()
    audio = ci + cq if sb == "usb" else ci - cq if sb == "lsb" else ci     # :214-221, float32 arithmetic
    if enable_noise_blanker:
        audio = noise_blanker(audio, noise_blanker_threshold_db, 3)
    if enable_highpass and highpass_hz > 0:
        audio = highpass_filter(audio, sample_rate, highpass_hz)
    if enable_lowpass and lowpass_hz > 0:
        audio = lowpass_filter(audio, sample_rate, lowpass_hz)
    for f in notch_frequencies or []:
        if 0 < f < sample_rate / 2:
            audio = notch_filter(audio, sample_rate, f)
    if enable_agc:
        audio = apply_agc(audio, sample_rate, target_db=agc_target_db, attack_ms=5.0, release_ms=50.0)
    audio = resample_poly(audio, sample_rate, audio_rate)
    return (audio if enable_agc else soft_clip_agc(audio)), f_off, pll


# ---- capture._process_channel_dsp_stateless ------------------------------------------------------

@dataclass
class OracleChannelConfig:
    """The fields of capture.ChannelConfig (capture.py:442-501) this path reads."""
    mode: str
    offset_hz: float = 0.0
    audio_rate: int = 48_000
    squelch_db: float | None = None
    enable_deemphasis: bool = True
    deemphasis_tau_us: float = 75.0
    enable_mpx_filter: bool = True
    mpx_cutoff_hz: float = 15_000
    enable_fm_highpass: bool = False
    fm_highpass_hz: float = 100
    enable_fm_lowpass: bool = False
    fm_lowpass_hz: float = 3_000
    enable_am_highpass: bool = True
    am_highpass_hz: float = 100
    enable_am_lowpass: bool = True
    am_lowpass_hz: float = 5_000
    enable_ssb_bandpass: bool = True
    ssb_bandpass_low_hz: float = 300
    ssb_bandpass_high_hz: float = 3_000
    ssb_mode: str = "usb"
    ssb_bfo_offset_hz: float = 1500.0
    sam_sideband: str = "dsb"
    sam_pll_bandwidth_hz: float = 50.0
    enable_agc: bool = False
    agc_target_db: float = -20.0
    notch_frequencies: list = field(default_factory=list)
    id: str = "ch"


def process_channel_dsp_stateless(samples, sample_rate, cfg):
    """capture.py:298-439: shift -> RSSI -> mode demod -> validity gate (validation.py:41-52) ->
    audio power."""
    metrics = {}
    if samples.size == 0:
        return None, metrics
    if not np.isfinite(samples).all():
        return None, metrics
    base = samples if cfg.offset_hz == 0.0 else freq_shift(samples, cfg.offset_hz, sample_rate)
    metrics["rssi_db"] = rssi_db(base)
    notch = cfg.notch_frequencies if cfg.notch_frequencies else None
    audio = None
    if cfg.mode == "wbfm":
        audio = wbfm_demod(base, sample_rate, cfg.audio_rate, cfg.enable_deemphasis, cfg.deemphasis_tau_us * 1e-6,
                           cfg.enable_mpx_filter, cfg.mpx_cutoff_hz, cfg.enable_fm_highpass, cfg.fm_highpass_hz,
                           notch)
    elif cfg.mode == "nbfm":
        audio = nbfm_demod(base, sample_rate, cfg.audio_rate, cfg.enable_deemphasis, cfg.deemphasis_tau_us * 1e-6,
                           cfg.enable_fm_highpass, cfg.fm_highpass_hz, cfg.enable_fm_lowpass, cfg.fm_lowpass_hz,
                           notch)
    elif cfg.mode == "am":
        audio = am_demod(base, sample_rate, cfg.audio_rate, cfg.enable_agc, cfg.enable_am_highpass,
                         cfg.am_highpass_hz, cfg.enable_am_lowpass, cfg.am_lowpass_hz, cfg.agc_target_db, notch)
    elif cfg.mode == "ssb":
        audio = ssb_demod(base, sample_rate, cfg.audio_rate, cfg.ssb_mode, cfg.enable_agc, cfg.enable_ssb_bandpass,
                          cfg.ssb_bandpass_low_hz, cfg.ssb_bandpass_high_hz, cfg.agc_target_db, notch,
                          cfg.ssb_bfo_offset_hz)
    elif cfg.mode == "sam":                                                # capture.py:385-398 (sam_demod_simple: fresh PLL)
        audio = sam_demod(base, sample_rate, cfg.audio_rate, cfg.sam_sideband, cfg.sam_pll_bandwidth_hz,
                          enable_agc=cfg.enable_agc, enable_highpass=cfg.enable_am_highpass, highpass_hz=cfg.am_highpass_hz,
                          enable_lowpass=cfg.enable_am_lowpass, lowpass_hz=cfg.am_lowpass_hz,
                          agc_target_db=cfg.agc_target_db)[0]
    elif cfg.mode == "raw":
        audio = np.empty(base.size * 2, dtype=F32)
        audio[0::2] = base.real
        audio[1::2] = base.imag
    elif cfg.mode in ("p25", "dmr", "nxdn", "dstar", "ysf"):
        metrics["signal_power_db"] = float(10.0 * np.log10(np.mean(np.abs(base) ** 2) + 1e-10))
        return None, metrics
    if audio is not None and audio.size > 0:
        if not (np.isfinite(audio).all() and float(np.max(np.abs(audio))) <= 1.2):
            return None, metrics
        metrics["signal_power_db"] = float(10.0 * np.log10(np.mean(audio ** 2) + 1e-10))
    return audio, metrics


def squelch(audio, metrics, squelch_db):
    """capture.py:2918-2921."""
    if audio is not None and squelch_db is not None and metrics.get("rssi_db", 0.0) < squelch_db:
        return np.zeros_like(audio)
    return audio


# ---- synthetic inputs ---------------------------------------------------------------

def synth_c1(seed=1, n=120_000, fs=2_400_000, offset=200_000.0, dev=75_000.0, amp=0.3, sigma=0.01, t0=0):
    """One WBFM carrier (1 kHz + 5 kHz tones) at +offset plus AWGN, cf32."""
    rng = np.random.default_rng(seed)
    t = (np.arange(n) + t0) / fs
    msg_phase = (dev / 1000.0) * 0.6 * np.sin(2 * np.pi * 1000.0 * t) + (dev / 5000.0) * 0.3 * np.sin(2 * np.pi * 5000.0 * t)
    x = amp * np.exp(1j * (2 * np.pi * offset * t + msg_phase))
    x = x + sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x.astype(np.complex64)


def synth_c2(seed=2, n=500_000, fs=10_000_000, n_ch=16, keyed_off=(), sigma=0.003):
    """16 NBFM carriers (5 kHz deviation, distinct tones) on a 500 kHz grid, int16 interleaved I,Q.
    Returns (int16 [n,2], offsets list)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    offs = [-3_750_000 + 500_000 * k for k in range(n_ch)]
    x = sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    for k, off in enumerate(offs):
        if k in keyed_off:
            continue
        tone = 300.0 + 170.0 * k
        amp = 0.02 + 0.002 * k
        x = x + amp * np.exp(1j * (2 * np.pi * off * t + (5000.0 / tone) * np.sin(2 * np.pi * tone * t)))
    q = np.empty((n, 2), dtype=np.int16)
    q[:, 0] = np.clip(np.round(x.real * 32768.0), -32768, 32767).astype(np.int16)
    q[:, 1] = np.clip(np.round(x.imag * 32768.0), -32768, 32767).astype(np.int16)
    return q, offs


def cs16_to_cf32(q):
    """cli.py:449-453: int16 pairs / 32768 -> complex64."""
    f = q.astype(F32) / F32(32768.0)
    return (f[..., 0] + 1j * f[..., 1]).astype(np.complex64)


# ---- optional clean-up stages (dsp/filters.py:267-459), restated with the same numpy/scipy calls ----

def noise_blanker(x, threshold_db=10.0, blanking_width=3):
    """dsp/filters.py:267-343."""
    from scipy.ndimage import binary_dilation

    if x.size == 0:
        return x.astype(np.float32, copy=False)
    mag = np.abs(x)
    med = np.median(mag)
    if med < 1e-10:
        return x.astype(np.float32, copy=False)
    mask = mag > med * (10 ** (threshold_db / 20.0))
    if not np.any(mask):
        return x.astype(np.float32, copy=False)
    if blanking_width > 0:
        mask = binary_dilation(mask, structure=np.ones(2 * blanking_width + 1, dtype=bool))
    y = x.copy()
    y[mask] = 0
    return y.astype(np.float32)


def spectral_noise_reduction(x, sample_rate, reduction_db=12.0, fft_size=1024, overlap=0.5):
    """dsp/filters.py:346-459."""
    from scipy import signal as sg

    if x.size == 0 or x.size < fft_size:
        return x.astype(np.float32, copy=False)
    hop = int(fft_size * (1 - overlap))
    window = sg.windows.hann(fft_size, sym=False).astype(np.float32)
    n_frames = (len(x) - fft_size) // hop + 1
    padded = (n_frames - 1) * hop + fft_size
    stft = np.zeros((n_frames, fft_size // 2 + 1), dtype=np.complex64)
    for i in range(n_frames):
        stft[i] = np.fft.rfft(x[i * hop:i * hop + fft_size] * window)
    mag, phase = np.abs(stft), np.angle(stft)
    noise = np.percentile(mag, 10, axis=0) * (10 ** (reduction_db / 20.0))
    gain = np.maximum(np.maximum(0.0, 1.0 - (noise / np.maximum(mag, 1e-10)) ** 2), 0.1)
    clean = mag * gain * np.exp(1j * phase)
    out = np.zeros(padded, dtype=np.float32)
    wsum = np.zeros(padded, dtype=np.float32)
    for i in range(n_frames):
        fr = np.fft.irfft(clean[i], n=fft_size).astype(np.float32)
        out[i * hop:i * hop + fft_size] += fr * window
        wsum[i * hop:i * hop + fft_size] += window ** 2
    out /= np.maximum(wsum, 1e-10)
    return out[:len(x)].astype(np.float32)
