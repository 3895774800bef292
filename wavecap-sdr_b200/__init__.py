"""wavecap_sdr_b200 — B200-native replacement for WaveCap-SDR's per-capture DSP hot path.

Module layout mirrors the reference (`wavecapsdr.*`) for the functions on that path only:

    wavecap_sdr_b200.dsp.channelizer   PolyphaseChannelizer, ChannelCalculator, channelize_samples
    wavecap_sdr_b200.dsp.fm / am / agc / filters
    wavecap_sdr_b200.dsp.fft           FFTBackend registry with the "cuda" backend
    wavecap_sdr_b200.dsp.p25.c4fm      C4FMDemodulator (+ benchmark_dsp.py helper classes)
    wavecap_sdr_b200.decoders.p25      CQPSKDemodulator
    wavecap_sdr_b200.capture           freq_shift, decimate_iq_for_p25, _process_channel_dsp_stateless
    wavecap_sdr_b200.install           monkey-patch the above over an importable `wavecapsdr`

All compute goes through libwcsdr_b200.so (hand-written sm_100a kernels behind a C ABI, see
include/wcsdr_b200.h). No CPU fallback exists.
"""
__version__ = "0.1.0"
