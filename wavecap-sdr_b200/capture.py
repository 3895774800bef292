"""GPU versions of the DSP entry points of `wavecapsdr.capture` (capture.py).

Mirrors: `ChannelConfig` (:442-501), `freq_shift` (:180-193), `decimate_iq_for_p25` (:203-295),
`_process_channel_dsp_stateless` (:298-439), `_validate_audio_output` (:147-162) and the squelch rule
of `_apply_stateful_processing` (:2918-2921). The threads / watchdogs / subscribers of the reference's
Capture class stay in the reference (out of scope).

GPU-only addition: `process_channels_batch` runs B chunks x C channels in a handful of launches,
reading each IQ chunk from HBM once for all channels (the reference submits one Python task per
channel to a 3-thread pool, capture.py:2489-2597).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from . import _native as N
from .dsp import _stages as S
from .dsp import am as AM
from .dsp import fm as FM
from .dsp import sam as SAM

AUDIO_MAX_ABS = 1.2  # validation.py:9


@dataclass
class ChannelConfig:
    """Per-channel parameter block; same fields and defaults as capture.py:442-501."""
    id: str
    capture_id: str
    mode: str
    offset_hz: float = 0.0
    audio_rate: int = 48_000
    squelch_db: float | None = None
    name: str | None = None
    auto_name: str | None = None
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
    agc_attack_ms: float = 5.0
    agc_release_ms: float = 50.0
    enable_noise_blanker: bool = False
    noise_blanker_threshold_db: float = 10.0
    notch_frequencies: list[float] = field(default_factory=list)
    enable_noise_reduction: bool = False
    noise_reduction_db: float = 12.0
    enable_rds: bool = True
    enable_pocsag: bool = False
    pocsag_baud: int = 1200


def apply_mode_defaults(mode: str, cfg: ChannelConfig) -> ChannelConfig:
    """Mode defaults of CaptureManager._apply_mode_defaults (capture.py:3425-3470)."""
    if mode == "wbfm":
        cfg.enable_deemphasis, cfg.deemphasis_tau_us = True, 75.0
        cfg.enable_mpx_filter, cfg.mpx_cutoff_hz = True, 15_000
        cfg.enable_fm_highpass = cfg.enable_fm_lowpass = cfg.enable_agc = False
    elif mode == "nbfm":
        cfg.enable_deemphasis = cfg.enable_mpx_filter = False
        cfg.enable_fm_highpass, cfg.fm_highpass_hz = False, 300
        cfg.enable_fm_lowpass, cfg.fm_lowpass_hz = False, 3_000
        cfg.enable_agc = False
    elif mode == "am":
        cfg.enable_am_highpass, cfg.am_highpass_hz = True, 100
        cfg.enable_am_lowpass, cfg.am_lowpass_hz = True, 5_000
        cfg.enable_agc = True
    elif mode == "ssb":
        cfg.enable_ssb_bandpass, cfg.ssb_bandpass_low_hz, cfg.ssb_bandpass_high_hz = True, 300, 3_000
        cfg.enable_agc = True
    return cfg


def _n(x) -> int:
    return int(x.numel()) if hasattr(x, "numel") else int(np.asarray(x).size)


def freq_shift(iq, offset_hz: float, sample_rate: int):
    """Mix with exp(-j 2 pi round(offset)/fs n), float32 phase restarted at n=0 (capture.py:166-193).
    Returns the input itself when offset is 0 or the input is empty."""
    if offset_hz == 0.0 or _n(iq) == 0:
        return iq
    x = S.to_device(iq, np.complex64).reshape(-1)
    _, base, _, _ = S.front(x, S.FMT_CF32, x.numel(), 1, [S.MODE_NONE], [float(offset_hz)], None, int(sample_rate),
                            want_out=False, want_base=True)
    return S.like_input(base.reshape(-1), iq)


def decimate_iq_for_p25(iq, sample_rate: int):
    """To ~48 kHz for P25 (capture.py:203-295): plain subsampling when down > 50, else polyphase
    resampling of I and Q."""
    target = 48000
    if sample_rate <= target or _n(iq) == 0:
        return iq, sample_rate
    g = math.gcd(sample_rate, target)
    up, down = target // g, sample_rate // g
    if down > 50:
        factor = sample_rate // target
        x = S.to_device(iq, np.complex64).reshape(-1)
        return S.like_input(x[::factor].contiguous(), iq), sample_rate // factor
    import torch

    x = S.to_device(iq, np.complex64).reshape(-1)
    ri = torch.view_as_real(x)
    rows = torch.stack([ri[:, 0], ri[:, 1]]).contiguous()
    y = S.resample(rows, up, down)
    out = torch.complex(y[0], y[1]).to(torch.complex64)
    return S.like_input(out, iq), (sample_rate * up) // down


_MODE_CODE = {"wbfm": S.MODE_WBFM, "nbfm": S.MODE_NBFM, "am": S.MODE_AM, "ssb": S.MODE_SSB, "raw": S.MODE_RAW}
_DIGITAL = ("p25", "dmr", "nxdn", "dstar", "ysf")


def _chain_signature(cfg: ChannelConfig, sample_rate: int):
    """(kind, iir stages, agc?, target, audio_rate) — channels with equal signatures share launches."""
    notch = tuple(cfg.notch_frequencies) if cfg.notch_frequencies else ()
    nb = None  # the reference's capture path never forwards the blanker flags to wbfm/nbfm_demod (capture.py:340-369)
    nr = float(cfg.noise_reduction_db) if cfg.enable_noise_reduction else None
    if cfg.mode == "wbfm":
        st = FM.fm_post_chain(sample_rate, wide=True, enable_deemphasis=cfg.enable_deemphasis,
                              deemphasis_tau=cfg.deemphasis_tau_us * 1e-6, enable_mpx_filter=cfg.enable_mpx_filter,
                              mpx_cutoff_hz=cfg.mpx_cutoff_hz, enable_highpass=cfg.enable_fm_highpass,
                              highpass_hz=cfg.fm_highpass_hz, notch_frequencies=notch)
        return ("fm", tuple(st), False, 0.0, int(cfg.audio_rate), nb, nr)
    if cfg.mode == "nbfm":
        st = FM.fm_post_chain(sample_rate, wide=False, enable_deemphasis=cfg.enable_deemphasis,
                              deemphasis_tau=cfg.deemphasis_tau_us * 1e-6, enable_highpass=cfg.enable_fm_highpass,
                              highpass_hz=cfg.fm_highpass_hz, enable_lowpass=cfg.enable_fm_lowpass,
                              lowpass_hz=cfg.fm_lowpass_hz, notch_frequencies=notch)
        return ("fm", tuple(st), False, 0.0, int(cfg.audio_rate), nb, nr)
    if cfg.mode == "am":
        st = AM.am_post_chain(sample_rate, cfg.enable_am_highpass, cfg.am_highpass_hz, cfg.enable_am_lowpass,
                              cfg.am_lowpass_hz, notch)
        return ("am", tuple(st), bool(cfg.enable_agc), float(cfg.agc_target_db), int(cfg.audio_rate))
    if cfg.mode == "ssb":
        st = AM.ssb_post_chain(sample_rate, cfg.enable_ssb_bandpass, cfg.ssb_bandpass_low_hz,
                               cfg.ssb_bandpass_high_hz, notch)
        return ("am", tuple(st), bool(cfg.enable_agc), float(cfg.agc_target_db), int(cfg.audio_rate))
    if cfg.mode == "raw":
        return ("raw",)
    if cfg.mode in _DIGITAL:
        return ("digital",)
    if cfg.mode == "sam":
        # sam_demod_simple (capture.py:385-398): the AM tail behind a carrier-recovery PLL; no notch list is forwarded
        st = AM.am_post_chain(sample_rate, cfg.enable_am_highpass, cfg.am_highpass_hz, cfg.enable_am_lowpass, cfg.am_lowpass_hz, ())
        return ("sam", tuple(st), bool(cfg.enable_agc), float(cfg.agc_target_db), int(cfg.audio_rate),
                SAM.SIDEBAND.get(str(cfg.sam_sideband).lower(), 0), float(cfg.sam_pll_bandwidth_hz))
    return ("unknown",)


def _plan_eligible(sig) -> bool:
    """chains the one-call plan (csrc/analog.cu wc_analog_run) covers: FM without blanker / spectral NR, AM / SSB, and the
    metrics-only modes. RAW (IQ pass-through), SAM (sequential carrier-recovery loop) and the optional clean-up stages take the
    stage-by-stage path."""
    if sig[0] == "fm":
        return sig[5] is None and sig[6] is None
    return sig[0] in ("am", "digital", "unknown")


def _run_plan(x, sample_rate, cfgs, sigs, modes, bfo, n, n_chunks, fmt, apply_squelch, return_device, results):
    """process_channels_batch through ONE wc_analog_run call (+ one device->host transfer of the metrics)."""
    from .analog_plan import get_plan

    chains = [s[:5] if s[0] in ("fm", "am") else (s[0],) for s in sigs]
    squelch = [(c.squelch_db if apply_squelch else None) for c in cfgs]
    plan = get_plan(sample_rate, n, fmt, modes, [float(c.offset_hz) for c in cfgs], bfo, squelch, chains)
    audio, metrics = plan.run(x.reshape(n_chunks, n, 2) if fmt == S.FMT_CS16 else x.reshape(n_chunks, n), n_chunks)
    m = metrics.cpu().numpy().astype(np.float64)          # [3][C][B]: rssi_db | signal_power_db | valid (one transfer)
    # the audio buffer is the plan's (overwritten by the next call): hand out views of ONE copy of it
    a_all = audio.clone() if return_device else audio.cpu().numpy()
    rssi_l, sig_l, valid_l = m[0].tolist(), m[1].tolist(), m[2].tolist()
    for ci, cfg in enumerate(cfgs):
        kind = sigs[ci][0]
        n_a, off = plan.audio_len[ci], n_chunks * plan.audio_off[ci]
        rows = a_all[off:off + n_chunks * n_a].reshape(n_chunks, n_a) if n_a > 0 else None
        for b in range(n_chunks):
            valid = valid_l[ci][b]
            if valid == 0.0:
                continue                                   # non-finite IQ: chunk dropped, empty metrics (capture.py:323-325)
            rssi = rssi_l[ci][b]
            if kind == "digital":
                results[b][ci] = (None, {"rssi_db": rssi, "signal_power_db": rssi})
            elif kind == "unknown" or valid < 1.0:
                results[b][ci] = (None, {"rssi_db": rssi})   # no audio path / _validate_audio_output failed (:433-435)
            else:
                results[b][ci] = (rows[b], {"rssi_db": rssi, "signal_power_db": sig_l[ci][b]})
    return results


def process_channels_batch(samples, sample_rate: int, cfgs: list[ChannelConfig], *, n_chunks: int = 1,
                           in_fmt: str = "cf32", apply_squelch: bool = False, return_device: bool = False,
                           want_fm_baseband: bool = False, use_plan: bool = True):
    """`_process_channel_dsp_stateless` for every (chunk, channel) pair of a batch.

    samples: complex64 [n_chunks*N] / [n_chunks, N] (in_fmt="cf32") or interleaved int16 I,Q
    [n_chunks, N, 2] (in_fmt="cs16", scaled by 1/32768 like cli.py:449-453); numpy or CUDA tensor.
    Returns results[chunk][channel] = (audio float32 | None, metrics dict) with the reference's keys
    (`rssi_db`, `signal_power_db`). With apply_squelch, audio of channels whose rssi_db is below
    cfg.squelch_db is zeroed (capture.py:2918-2921). With want_fm_baseband, WBFM channels with enable_rds at a capture
    rate >= 114 kHz also get `fm_baseband` in their metrics: quadrature_demod(freq_shift(iq)) before the MPX filter, the
    input the reference hands its RDS decoder (capture.py:2869-2884) — it is the front end's own output, no extra pass.
    """
    import torch

    N.ensure_init()
    n_ch = len(cfgs)
    fmt = S.FMT_CS16 if in_fmt == "cs16" else S.FMT_CF32
    total = int(samples.numel()) if N.is_torch_cuda(samples) else int(np.asarray(samples).size)
    n = total // (2 * n_chunks) if fmt == S.FMT_CS16 else total // n_chunks
    results = [[(None, {}) for _ in range(n_ch)] for _ in range(n_chunks)]
    if n == 0 or n_ch == 0:
        return results
    if use_plan and not want_fm_baseband:
        sigs = [_chain_signature(c, sample_rate) for c in cfgs]
        if all(_plan_eligible(s) for s in sigs):
            modes = [_MODE_CODE.get(c.mode, S.MODE_NONE) for c in cfgs]
            bfo = [(c.ssb_bfo_offset_hz if c.ssb_mode.lower() == "usb" else -c.ssb_bfo_offset_hz) if c.mode == "ssb" else 0.0
                   for c in cfgs]
            if N.is_torch_cuda(samples):
                src = S.to_device(samples, np.int16 if fmt == S.FMT_CS16 else np.complex64)
            else:   # staged into the plan's own device buffer (fixed address: repeated calls replay the captured graph)
                src = np.ascontiguousarray(samples, dtype=np.int16 if fmt == S.FMT_CS16 else np.complex64)
            return _run_plan(src, sample_rate, cfgs, sigs, modes, bfo, n, n_chunks, fmt, apply_squelch, return_device, results)
    if in_fmt == "cs16":
        x = S.to_device(samples, np.int16).reshape(n_chunks, -1, 2)
    else:
        x = S.to_device(samples, np.complex64).reshape(n_chunks, -1)

    sigs = [_chain_signature(c, sample_rate) for c in cfgs]
    modes = [_MODE_CODE.get(c.mode, S.MODE_NONE) for c in cfgs]
    bfo = [(c.ssb_bfo_offset_hz if c.ssb_mode.lower() == "usb" else -c.ssb_bfo_offset_hz) if c.mode == "ssb" else 0.0
           for c in cfgs]
    want_base = any(s[0] in ("raw", "sam") for s in sigs)
    # channels whose chain has nothing between discriminator and rms_normalize take sum(out**2) from the front end
    want_ss = any(s[0] == "fm" and not s[1] and s[5] is None and s[6] is None for s in sigs)
    out, base, power, nonfinite, *rest = S.front(x, fmt, n, n_chunks, modes, [float(c.offset_hz) for c in cfgs], bfo,
                                                 int(sample_rate), want_out=True, want_base=want_base, want_sumsq=want_ss)
    out_ss = rest[0] if rest else None

    audio = [None] * n_ch          # per channel: CUDA [n_chunks, n_audio]
    stat_parts = []                # per run of channels: [k, 2, n_chunks] float64 (audio power, invalid flag)
    have = []                      # channel index of every row of the concatenated statistics
    c = 0
    while c < n_ch:                 # runs of adjacent channels with identical chains share launches
        e = c + 1
        while e < n_ch and sigs[e] == sigs[c]:
            e += 1
        sig = sigs[c]
        if sig[0] in ("fm", "am"):
            rows = out[c:e].reshape((e - c) * n_chunks, n)
            if sig[0] == "fm":
                ss_rows = out_ss[c:e].reshape(-1) if out_ss is not None else None
                a, p, inv = FM.fm_tail(rows, int(sample_rate), sig[4], sig[1], want_stats=True, blanker_db=sig[5], nr_db=sig[6],
                                       sumsq_rows=ss_rows)
            else:
                a, p, inv = AM.am_tail(rows, int(sample_rate), sig[4], sig[1], sig[2], sig[3], want_stats=True)
            a = a.reshape(e - c, n_chunks, -1)
            # per-run statistics stay one tensor [channels of the run][power | invalid][chunk]: a handful of launches per
            # run instead of three per channel
            stat_parts.append(torch.stack([p.reshape(e - c, n_chunks).double(), inv.reshape(e - c, n_chunks).double()], dim=1))
            have.extend(range(c, e))
            for i in range(c, e):
                audio[i] = a[i - c]
        elif sig[0] == "sam":
            # fresh PLL per (channel, chunk) like the stateless reference call; every sequence of the run advances in one launch
            alpha, beta = SAM.pll_coefficients(float(sample_rate), sig[6], 0.707)
            rows, _, _ = SAM.pll_rows(base[c:e].reshape((e - c) * n_chunks, n), alpha, beta, sig[5])
            a, p, inv = AM.am_tail(rows, int(sample_rate), sig[4], sig[1], sig[2], sig[3], want_stats=True)
            a = a.reshape(e - c, n_chunks, -1)
            stat_parts.append(torch.stack([p.reshape(e - c, n_chunks).double(), inv.reshape(e - c, n_chunks).double()], dim=1))
            have.extend(range(c, e))
            for i in range(c, e):
                audio[i] = a[i - c]
        elif sig[0] == "raw":
            for i in range(c, e):
                audio[i] = torch.view_as_real(base[i]).reshape(n_chunks, 2 * n)   # interleaved I,Q (capture.py:415-420)
                pw_i = (audio[i].double() ** 2).sum(dim=1)
                inv_i = ((~torch.isfinite(audio[i]).all(dim=1)) | (audio[i].abs().amax(dim=1) > AUDIO_MAX_ABS))
                stat_parts.append(torch.stack([pw_i, inv_i.double()]).unsqueeze(0))
                have.append(i)
        c = e

    # one device->host transfer for all per-(channel, chunk) statistics instead of three per channel
    stats_h = None
    if have:
        stats = stat_parts[0] if len(stat_parts) == 1 else torch.cat(stat_parts, dim=0)  # [k][2][n_chunks], rows in `have` order
        stats_h = stats.cpu().numpy()
    power_h = power.cpu().numpy()
    nonfinite_h = nonfinite.cpu().numpy()
    slot = {ci: j for j, ci in enumerate(have)}
    # dB values for the whole [channel][chunk] grid at once, in the reference's float32 arithmetic (capture.py:331-334,436-437)
    rssi_all = (np.float32(10.0) * np.log10((power_h / n).astype(np.float32) + np.float32(1e-10))).astype(np.float64)
    for ci, cfg in enumerate(cfgs):
        a_h = inv_h = sp_db = None
        if audio[ci] is not None:
            inv_h = stats_h[slot[ci], 1] != 0
            a_h = audio[ci] if return_device else audio[ci].cpu().numpy()
            n_a = int(a_h.shape[-1])
            sp_db = 10.0 * np.log10((stats_h[slot[ci], 0] / max(n_a, 1)).astype(np.float32) + np.float32(1e-10))
        digital = sigs[ci][0] == "digital"
        squelch = cfg.squelch_db if (apply_squelch and cfg.squelch_db is not None) else None
        for b in range(n_chunks):
            if nonfinite_h[b]:
                continue            # non-finite IQ: chunk dropped, empty metrics (capture.py:323-325)
            rssi = float(rssi_all[ci, b])
            if digital:
                # same power of the shifted IQ (capture.py:426-428)
                results[b][ci] = (None, {"rssi_db": rssi, "signal_power_db": rssi})
                continue
            if a_h is None or inv_h[b]:
                results[b][ci] = (None, {"rssi_db": rssi})   # no audio path / _validate_audio_output failed (:433-435)
                continue
            au = a_h[b]
            if squelch is not None and rssi < squelch:
                au = torch.zeros_like(au) if return_device else np.zeros_like(au)
            results[b][ci] = (au, {"rssi_db": rssi, "signal_power_db": float(sp_db[b])})
            if want_fm_baseband and cfg.mode == "wbfm" and cfg.enable_rds and sample_rate >= 114000:
                results[b][ci][1]["fm_baseband"] = out[ci, b] if return_device else out[ci, b].cpu().numpy()
    return results


def _process_channel_dsp_stateless(samples, sample_rate: int, cfg: ChannelConfig):
    """Stateless per-channel DSP (capture.py:298-439): (audio | None, {rssi_db, signal_power_db})."""
    if _n(samples) == 0:
        return None, {}
    return process_channels_batch(samples, sample_rate, [cfg], n_chunks=1)[0][0]


def _validate_audio_output(audio, context: str = "") -> bool:
    """finite and max|x| <= 1.2 (capture.py:147-162, validation.py:41-52)."""
    a = audio.cpu().numpy() if hasattr(audio, "is_cuda") else np.asarray(audio)
    if a.size == 0:
        return True
    return bool(np.isfinite(a).all() and float(np.max(np.abs(a))) <= AUDIO_MAX_ABS)


# ---- output stage: wire packers and audio level metering (capture.py:102-144, 633-661) ----------------------------

def _pack(samples, fmt: int, np_dtype):
    import torch

    if _n(samples) == 0:
        return b""
    if N.is_torch_cuda(samples):
        x = samples.contiguous()
        x = torch.view_as_real(x.to(torch.complex64)).reshape(-1) if x.is_complex() else x.to(torch.float32).reshape(-1)
    else:
        a = np.asarray(samples)
        a = a.astype(np.complex64, copy=False).view(np.float32) if np.iscomplexobj(a) else np.ascontiguousarray(a, dtype=np.float32)
        x = torch.from_numpy(np.ascontiguousarray(a).reshape(-1)).cuda()
    N.ensure_init()
    out = torch.empty((x.numel(),), dtype=torch.int16 if fmt == 0 else torch.float32, device=x.device)
    N.check(N.lib().wc_pack(S.ptr(x), S.ptr(out), int(x.numel()), fmt, S.stream()))
    return out.cpu().numpy().astype(np_dtype, copy=False).tobytes()


def pack_iq16(samples) -> bytes:
    """Interleaved int16 I,Q: clip to [-1, 1], x 32767, truncate (capture.py:102-115). Unlike the reference, the
    caller's array is not clipped in place."""
    return _pack(samples, 0, np.int16)


def pack_pcm16(samples) -> bytes:
    """16-bit PCM (capture.py:118-130)."""
    return _pack(samples, 0, np.int16)


def pack_f32(samples) -> bytes:
    """Clipped float32 (capture.py:133-144)."""
    return _pack(samples, 1, np.float32)


def audio_levels(audio_rows):
    """Channel._update_audio_metrics (capture.py:633-661) for a batch: audio_rows CUDA/numpy float32 [n_seq, n] ->
    (rms_db [n_seq], peak_db [n_seq], clipping_count [n_seq]) with the reference's -100 dB floor."""
    import torch

    N.ensure_init()
    x = S.to_device(audio_rows, np.float32)
    x = x.reshape(1, -1) if x.dim() == 1 else x.contiguous()
    n_seq, n = int(x.shape[0]), int(x.shape[1])
    ss = torch.empty((n_seq,), dtype=torch.float64, device=x.device)
    pk = torch.empty((n_seq,), dtype=torch.float32, device=x.device)
    cc = torch.empty((n_seq,), dtype=torch.int32, device=x.device)
    N.check(N.lib().wc_audio_levels(S.ptr(x), n, n, n_seq, S.ptr(ss), S.ptr(pk), S.ptr(cc), S.stream()))
    rms = np.sqrt((ss.cpu().numpy() / n).astype(np.float32)).astype(np.float64)
    peak = pk.cpu().numpy().astype(np.float64)
    with np.errstate(divide="ignore"):
        rms_db = np.where(rms > 1e-10, 20.0 * np.log10(np.maximum(rms, 1e-300)), -100.0)
        peak_db = np.where(peak > 1e-10, 20.0 * np.log10(np.maximum(peak, 1e-300)), -100.0)
    return rms_db, peak_db, cc.cpu().numpy()


def signal_metrics(iq, sample_rate: int, offsets_hz, want_snr: bool = True, in_fmt: str = "cf32"):
    """`Channel.update_signal_metrics` (capture.py:749-798) for every channel of a chunk in one call: freq_shift by each
    channel's offset, RSSI from the mean squared magnitude, and the SNR estimate from the magnitudes np.partition would
    put at ranks n//10 and n - n//10 - 1 (exact radix select on the GPU). Returns (rssi_db float list, snr_db list with
    None where the reference leaves None)."""
    import torch

    N.ensure_init()
    fmt = {"cf32": 0, "cs16": 1}[in_fmt]
    x = S.to_device(iq, np.complex64 if fmt == 0 else np.int16)
    n = int(x.numel()) if fmt == 0 else int(x.numel()) // 2
    offs = np.ascontiguousarray(np.atleast_1d(np.asarray(offsets_hz, dtype=np.float64)))
    k = int(offs.size)
    if n == 0:
        return [None] * k, [None] * k
    mag = torch.empty((k, n), dtype=torch.float32, device=x.device)
    power = torch.empty((k,), dtype=torch.float64, device=x.device)
    pct = torch.empty((k, 2), dtype=torch.float32, device=x.device)
    scratch = torch.empty((int(N.lib().wc_front_chan_scratch_bytes(k)),), dtype=torch.uint8, device=x.device)
    N.check(N.lib().wc_signal_metrics(S.ptr(x), fmt, n, int(sample_rate), N.np_ptr(offs), k, 1 if want_snr else 0,
                                      S.ptr(mag), S.ptr(power), S.ptr(pct), S.ptr(scratch), S.stream()))
    p32 = (power.cpu().numpy() / n).astype(np.float32)
    rssi = [float(np.float32(10.0) * np.log10(p + np.float32(1e-10))) for p in p32]
    snr: list = [None] * k
    k_noise, k_signal = n // 10, n - n // 10 - 1
    if want_snr and k_noise > 0 and k_signal > k_noise:
        q = pct.cpu().numpy()
        for c in range(k):
            noise_power, signal_power = q[c, 0] ** 2, q[c, 1] ** 2
            if noise_power > 1e-10:
                snr[c] = float(np.float32(10.0) * np.log10(signal_power / noise_power))
    return rssi, snr


class SignalMeter:
    """The per-channel state `update_signal_metrics` keeps (rssi_db, snr_db, the every-10th-call SNR throttle,
    capture.py:773-777) for a bank of channels metered together."""

    def __init__(self, offsets_hz):
        self.offsets_hz = list(np.atleast_1d(offsets_hz))
        self.rssi_db = [None] * len(self.offsets_hz)
        self.snr_db = [None] * len(self.offsets_hz)
        self._snr_counter = 0

    def update(self, iq, sample_rate: int, in_fmt: str = "cf32") -> None:
        if _n(iq) == 0:
            return
        self._snr_counter += 1
        want = self._snr_counter % 10 == 0
        rssi, snr = signal_metrics(iq, sample_rate, self.offsets_hz, want_snr=want, in_fmt=in_fmt)
        self.rssi_db = rssi
        if want:
            self.snr_db = snr
