/* wcsdr_b200 — C ABI of the B200-native DSP hot path for WaveCap-SDR.
 *
 * The reference (TobiasWooldridge/WaveCap-SDR) is pure Python: its hot path has no FFI today, the
 * boundary is "Python function/class names + numpy array contracts" (SURVEY.md §8b). This header is
 * the C boundary a maintainer binds with ctypes: plain pointers, sizes, opaque
 * handles, `int` status (0 = ok, <0 = error, message via wc_last_error()). No torch types.
 *
 * Pointer naming: `*_dev` = device pointer (any allocator: cudaMalloc, torch, cupy), `*_host` = host
 * pointer (pinned memory recommended). `stream` is a cudaStream_t passed as void*, taken literally
 * (NULL = the CUDA default stream); `*_host` entry points use the handle's own stream and synchronise. All functions are thread-compatible per handle; stateless entry points are
 * re-entrant.
 *
 * Reference paths are relative to /root/reference/backend/.
 */
#ifndef WCSDR_B200_H
#define WCSDR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ---------------------------------------------------------------------------------- */
int wc_init(int device);              /* cudaSetDevice + capability check (needs sm_100)          */
const char* wc_last_error(void);      /* thread-local message of the last failing call            */
int wc_device_info(int* sm_count, int* cc_major, int* cc_minor, long long* total_mem_bytes);
const char* wc_version(void);

/* ---- polyphase channelizer: wavecapsdr/dsp/channelizer.py:28-158 (PolyphaseChannelizer) -------- */
typedef struct wc_chan wc_chan;
#define WC_CHAN_OUT_COMPLEX 0 /* process(): complex64 frames [F][M]                channelizer.py:91  */
#define WC_CHAN_OUT_FM 1      /* fused quadrature_demod() of every bin: float32 [F][M] dsp/fm.py:65   */
#define WC_CHAN_OUT_AUDIO 2   /* nbfm_demod() of every bin: float32 [ceil(F/D)][M]   dsp/fm.py:317-406 */
#define WC_CHAN_IN_CF32 0     /* complex64 IQ                                                          */
#define WC_CHAN_IN_CS16 1     /* interleaved int16 I,Q, scaled by 1/32768 like cli.py:449-453          */

/* __init__ (channelizer.py:35-67): M = int(fs/bw) made even; prototype = firwin(M*T-1, 0.9*bw/(fs/2),
 * kaiser 8.0) designed on the host inside the library (no scipy at run time). */
int wc_chan_create(double sample_rate, int channel_bandwidth, int taps_per_channel, wc_chan** out);
void wc_chan_destroy(wc_chan* h);
int wc_chan_info(const wc_chan* h, int* channel_count, double* channel_sample_rate, int* taps_per_channel);
int wc_chan_get_arms(const wc_chan* h, double* arms /* [M][T] float64, .arms attribute */);
int wc_chan_get_history(wc_chan* h, void* arm_history_host /* complex64 [M][T], .arm_history */);
long long wc_chan_frames_for(const wc_chan* h, long long n_samples); /* floor((N-M)/(M/2))+1 or 0     */
int wc_chan_reset(wc_chan* h);                                       /* reset(), channelizer.py:139   */
/* process(): n_chunks consecutive process() calls of n_samples each in ONE launch (chunk c starts
 * at iq + c*chunk_stride samples; frames never straddle chunks, exactly like consecutive calls).
 * Output: [n_chunks*F][M] complex64 (mode 0) or float32 (mode 1, scaled by fm_scale). */
int wc_chan_process(wc_chan* h, const void* iq_dev, long long n_samples, int n_chunks, long long chunk_stride,
                    int mode, float fm_scale, void* out_dev, void* stream);
/* advance the carried history (arm_history) as if process() had been called on iq_dev[0..n_samples) without
 * computing outputs — for time-sharded multi-GPU runs where another rank emitted the tail of the call. */
int wc_chan_carry_from(wc_chan* h, const void* iq_dev, long long n_samples, void* stream);
/* the same from just the tail of the previous call: tail_dev = its samples [(F - T) * M/2, (F + 1) * M/2) (T + 1 hop rows,
 * complex64), e.g. fetched from the GPU that owns the end of the previous block of a striped capture. */
int wc_chan_carry_tail(wc_chan* h, const void* tail_dev, void* stream);
/* frames each CTA of the M=256 kernels walks (rounded up to 8, clamped to [16, 256]); 0 = sized so the grid is ~6 waves.
 * Every run re-reads T-1 halo rows, which local L2 absorbs but NVLink does not: callers whose iq_dev is peer memory
 * (wc_peer_open) raise it so the halo traffic stays at (T-1)/run_frames of the slab. */
int wc_chan_set_run_frames(wc_chan* h, int run_frames);
/* same, host buffers: H2D copy + kernels + D2H copy + sync (the reference-facing call). */
int wc_chan_process_host(wc_chan* h, const void* iq_host, long long n_samples, int n_chunks, int mode,
                         float fm_scale, void* out_host);

/* Input formats and the audio mode (SURVEY §8d "C5 + audio /20"). in_fmt WC_CHAN_IN_CS16 reads the capture as
 * interleaved int16 I,Q (4 bytes per sample over PCIe and HBM instead of 8; converted to the exact float32 the
 * reference's /32768.0 produces as each row enters the FIR; M = 256 / 9-tap grid, 16-byte aligned chunks).
 * Mode WC_CHAN_OUT_AUDIO is nbfm_demod(extract_channel(process(x), k), demod_rate, audio_rate) with the reference's
 * defaults (dsp/fm.py:317-406: quadrature_demod -> rms_normalize over the chunk -> resample_poly -> soft_clip) for
 * every channel k: float32 [n_chunks][wc_chan_audio_len][M], built for integer decimation (audio_rate divides
 * demod_rate: resample_poly's up = 1), e.g. 976560 -> 48828 = /20 for the 125 MS/s grid. fm_scale is ignored in this
 * mode (the discriminator scale is float32(demod_rate/(2 pi 75000)), dsp/fm.py:94). */
int wc_chan_audio_config(wc_chan* h, int demod_rate, int audio_rate);
long long wc_chan_audio_len(const wc_chan* h, long long n_samples);   /* ceil(frames / D) per chunk and channel */
int wc_chan_process_ex(wc_chan* h, const void* iq_dev, int in_fmt, long long n_samples, int n_chunks, long long chunk_stride,
                       int mode, float fm_scale, void* out_dev, void* stream);
int wc_chan_process_host_ex(wc_chan* h, const void* iq_host, int in_fmt, long long n_samples, int n_chunks, int mode,
                            float fm_scale, void* out_host);

/* ---- analog demod chain: wavecapsdr/capture.py:298-439, dsp/fm.py, dsp/am.py, dsp/agc.py, dsp/filters.py ----
 * Stage-level operators on device buffers; the Python host (wavecap_sdr_b200/capture.py, dsp/*.py) chains
 * them exactly as the reference chains numpy/scipy calls. Sequences are channel-major:
 * seq = channel*n_chunks + chunk, each `seq_stride` floats apart. */
#define WC_MODE_NONE 0 /* p25/dmr/...: only RSSI power (capture.py:422-428) */
#define WC_MODE_WBFM 1
#define WC_MODE_NBFM 2
#define WC_MODE_AM 3
#define WC_MODE_SSB 4
#define WC_MODE_RAW 5
#define WC_FMT_CF32 0
#define WC_FMT_CS16 1 /* interleaved int16 I,Q scaled by 1/32768 (cli.py:449-453) */

/* capture.freq_shift (capture.py:166-193; float32 phase, restarted per chunk, offset rounded to int Hz,
 * untouched when offset == 0) + RSSI power sum (capture.py:331-334) + demod front end
 * (FM: dsp/fm.py:65-97; AM: |x| dsp/am.py:99; SSB: Re(x*exp(+j2pi*bfo*t)) dsp/am.py:23-42,223) for every
 * channel of every chunk in ONE launch; the IQ tile is staged in shared memory once for all channels.
 * modes/offsets_hz/bfo_hz are host arrays [n_ch]; out_dev float32 [n_ch][n_chunks][n] (float2 for RAW);
 * base_out_dev optional complex64 [n_ch][n_chunks][n]; power_dev float64 [n_ch][n_chunks] (sum |x|^2);
 * nonfinite_dev int32 [n_chunks] (validation.py:37-38); chan_scratch_dev >= wc_front_chan_scratch_bytes. */
int wc_front_chan_scratch_bytes(int n_ch);
int wc_front_run(const void* iq_dev, int fmt, int n, int n_chunks, long long chunk_stride, int n_ch,
                 const int* modes, const double* offsets_hz, const double* bfo_hz, int sample_rate,
                 float* out_dev, void* base_out_dev, double* power_dev, int* nonfinite_dev, void* chan_scratch_dev,
                 void* stream);
/* same, and out_sumsq_dev float64 [n_ch][n_chunks] (optional) = sum(out**2) per sequence: the input of
 * rms_normalize (dsp/fm.py:42-62) when no filter sits between discriminator and normaliser (nbfm defaults,
 * capture.py:3444-3452) — saves the separate wc_sumsq pass over the full-rate signal. */
int wc_front_run_ex(const void* iq_dev, int fmt, int n, int n_chunks, long long chunk_stride, int n_ch,
                    const int* modes, const double* offsets_hz, const double* bfo_hz, int sample_rate,
                    float* out_dev, void* base_out_dev, double* power_dev, double* out_sumsq_dev, int* nonfinite_dev,
                    void* chan_scratch_dev, void* stream);

/* scipy.signal.lfilter(b, a, x), zero initial state, float64 DF2T, order <= 10
 * (dsp/fm.py:123,178; dsp/filters.py:124,170,217,260; dsp/agc.py:93,100). Well-conditioned filters run as a block
 * scan; tf-form filters with large transient growth (the order-10 band-pass of dsp/filters.py:177-217) replay
 * lfilter's recursion operation for operation (bit-equal to scipy) — chosen at create time, wc_iir_is_sequential.
 * Handles may be shared between threads: every call owns its scratch (stream-ordered allocation). */
typedef struct wc_iir wc_iir;
int wc_iir_create(const double* b, int nb, const double* a, int na, wc_iir** out);
void wc_iir_destroy(wc_iir* h);
int wc_iir_is_sequential(const wc_iir* h);
int wc_iir_kind(const wc_iir* h);   /* 0 scan chained in double-double, 1 scan chained in float64, 2 sequential replay */
int wc_iir_lfilter(wc_iir* h, const float* x_dev, float* y_dev, int n, long long seq_stride, int n_seq,
                   int abs_input, void* stream);

/* sum(x**2) per sequence -> float64 (dsp/fm.py:58 rms_normalize, capture.py:436 signal power) */
int wc_sumsq(const float* x_dev, int n, long long seq_stride, int n_seq, double* out_dev, void* stream);
/* op 0: dsp.fm.soft_clip (fm.py:26-39); 1: dsp.agc.soft_clip (agc.py:58-70); 2: y = x * p0 */
int wc_elementwise(const float* x_dev, float* y_dev, long long total, int op, float p0, void* stream);
/* apply_agc tail (dsp/agc.py:228-242) given the two envelope-filter outputs */
int wc_agc_apply(const float* x_dev, const float* env_attack_dev, const float* env_release_dev, float* y_dev,
                 long long total, float target_linear, float max_gain_linear, void* stream);

/* scipy.signal.resample_poly(x.astype(f64), up, down) -> float32 (dsp/fm.py:184-221): zero-extended upfirdn,
 * only kept outputs computed, float64 accumulation. `taps` = firwin(2*10*max(up,down)+1, 1/max(up,down),
 * ("kaiser", 5.0)) * up, designed by the host like the reference does. */
typedef struct wc_resampler wc_resampler;
#define WC_EPI_NONE 0
#define WC_EPI_RMS_CLIP 1 /* rms_normalize (fm.py:42-62) then fm soft_clip: the wbfm/nbfm tail */
#define WC_EPI_CLIP 2     /* fm soft_clip only */
#define WC_EPI_RMS 3      /* rms_normalize only */
#define WC_EPI_CLIP_AGC 4 /* dsp.agc.soft_clip (am/ssb without AGC, am.py:139-141) */
int wc_resampler_create(int up, int down, const double* taps, int ntaps, wc_resampler** out);
void wc_resampler_destroy(wc_resampler* h);
long long wc_resampler_out_len(const wc_resampler* h, long long n_in);
int wc_resampler_run(wc_resampler* h, const float* x_dev, int n_in, long long seq_stride, int n_seq, float* out_dev,
                     int epilogue, const double* sumsq_dev, float target_rms, float min_rms, double* power_dev,
                     int* invalid_dev, float max_abs, void* stream);
/* RSSI dB per sequence + squelch select (capture.py:331-334, 2918-2921) */
int wc_finalize(float* audio_dev, int n_out, int n_chunks, int n_seq, const double* power_iq_dev, int n_in,
                const float* squelch_db_dev, const int* has_squelch_dev, float* rssi_db_dev,
                unsigned char* squelched_dev, void* stream);

/* Synchronous AM carrier recovery: CarrierRecoveryPLL.process (dsp/sam.py:73-123) for n_seq independent sequences —
 * lo = exp(-1j phase), mixed = iq * lo, phase error arctan2(Q, |I| + 1e-10), second-order loop filter, all float64,
 * sequential per sequence. state_dev float64 [n_seq][3] = (phase, frequency, integrator), read and updated (zeros = the
 * fresh PLL sam_demod builds, :195-198). Optional float32 [n_seq][n] outputs: audio_dev = sam_demod's sideband selection
 * (:214-221; 0 dsb, 1 usb, 2 lsb), coh_i_dev / coh_q_dev = the coherent components. exact != 0 replays the reference's
 * float64 arithmetic (coherent I/Q bit-equal on the goldens); exact == 0 keeps the loop state in float64 and evaluates the
 * oscillator, mixer and phase detector in float32 (faster; ~1e-6 relative RMS on a single carrier, up to 2e-4 with strong
 * neighbours in the band, see tests/test_sam_gpu.py). */
int wc_sam_pll(const void* iq_dev, long long seq_stride, int n, int n_seq, double alpha, double beta, int sideband, int exact,
               double* state_dev, float* audio_dev, float* coh_i_dev, float* coh_q_dev, void* stream);

/* ---- analog plan: the whole analog chain of a capture as one call ------------------------------------
 * Replaces one _process_channel_dsp_stateless call per (chunk, channel) from the capture's worker pool
 * (capture.py:298-439, caller :2489-2597) by ONE call per batch of chunks: front end (freq_shift, RSSI power, demod)
 * -> [IIR stages] -> [apply_agc] -> rms_normalize -> resample_poly with fused scale / soft clip -> dB metrics, validity
 * gate (validation.py:41-52), squelch (capture.py:2918-2921). The plan owns handles, scratch and statistics; a call
 * allocates nothing and never synchronises; repeated (iq_dev, audio_dev, metrics_dev, n_chunks) keys replay a captured
 * CUDA graph. Filter design stays with the caller (scipy, as in the reference): stages arrive as (b, a) pairs.
 * Build: create -> add_run (adjacent channels sharing one chain; kind 0 metrics only, 1 FM, 2 AM/SSB) -> add_iir* ->
 * [set_agc] -> finish. */
typedef struct wc_analog_plan wc_analog_plan;
int wc_analog_plan_create(int sample_rate, int chunk_len, int in_fmt /* 0 cf32, 1 cs16 */, int n_channels, const int* modes,
                          const double* offsets_hz, const double* bfo_hz, const float* squelch_db /* NaN = none */,
                          wc_analog_plan** out);
int wc_analog_plan_add_run(wc_analog_plan* p, int first, int count, int kind, int up, int down, const double* taps, int n_taps);
int wc_analog_plan_add_iir(wc_analog_plan* p, int run, const double* b, int nb, const double* a, int na);
int wc_analog_plan_set_agc(wc_analog_plan* p, int run, const double* attack_b, const double* attack_a, const double* release_b,
                           const double* release_a, float target_linear, float max_gain_linear);
int wc_analog_plan_finish(wc_analog_plan* p);
void wc_analog_plan_destroy(wc_analog_plan* p);
long long wc_analog_plan_audio_floats(const wc_analog_plan* p);            /* audio floats per chunk, all channels   */
int wc_analog_plan_audio_len(const wc_analog_plan* p, int channel);        /* audio samples per chunk of a channel   */
long long wc_analog_plan_audio_offset(const wc_analog_plan* p, int channel);
int wc_analog_plan_use_graph(wc_analog_plan* p, int on);                   /* default on                             */
/* audio_dev: channel c, chunk b at audio_dev + n_chunks * audio_offset(c) + b * audio_len(c); metrics_dev: float32
 * [3][n_channels][n_chunks] = rssi_db | signal_power_db | valid (1 audio valid, 0.5 RSSI only, 0 chunk dropped). */
int wc_analog_run(wc_analog_plan* p, const void* iq_dev, int n_chunks, float* audio_dev, float* metrics_dev, void* stream);

/* ---- spectrum / waterfall: wavecapsdr/dsp/fft/base.py:31-77 (FFTBackend), scipy_backend.py:38-79 ----
 * power_db = float32(20*log10(|fftshift(fft(iq[:N] * float32(hanning(N))))| + 1e-10)); consecutive groups of
 * `avg` frames are averaged in dB (the frontend's spectrum averaging, SpectrumAnalyzer.react.tsx:309-327).
 * fft_size must be a power of two. This is the kernel the "cuda" slot of the FFT registry
 * (dsp/fft/registry.py:167-174, today CuPy -> cuFFT) binds to. */
typedef struct wc_spectrum wc_spectrum;
int wc_spectrum_create(int fft_size, wc_spectrum** out);
void wc_spectrum_destroy(wc_spectrum* h);
int wc_spectrum_window(const wc_spectrum* h, float* window_host /* [fft_size], FFTBackend.window */);
int wc_spectrum_execute(wc_spectrum* h, const void* iq_dev, long long frame_stride, int n_frames, int avg,
                        float* power_db_dev /* [ceil(n_frames/avg)][fft_size] */, void* stream);
int wc_spectrum_execute_host(wc_spectrum* h, const void* iq_host, long long frame_stride, int n_frames, int avg,
                             float* power_db_host);

/* ---- P25 Phase-1 C4FM symbol recovery: wavecapsdr/dsp/p25/c4fm.py:2379-2807 (C4FMDemodulator) ----
 * One handle = n_channels independent stateful demodulators advanced by the same call (the reference runs one
 * Python object per control/voice channel). Filters: pass the float32 designs of design_baseband_lpf
 * (c4fm.py:95-132) and design_rrc_filter (:135-183), or NULL/0 to have them designed inside the library.
 * demod(): iq complex64 [n_channels][chan_stride] (first n_samples used) -> dibits uint8 / soft float32
 * [n_channels][max_sym], n_sym int32 [n_channels]; max_sym >= wc_c4fm_max_symbols(h, n_samples). Output depends on
 * the call (chunk) sequence exactly like the reference. */
typedef struct wc_c4fm wc_c4fm;
int wc_c4fm_create(int n_channels, int sample_rate, int symbol_rate, int wide_pulse, const float* lpf_taps, int n_lpf,
                   const float* rrc_taps, int n_rrc, wc_c4fm** out);
void wc_c4fm_destroy(wc_c4fm* h);
int wc_c4fm_info(const wc_c4fm* h, int* n_channels, double* samples_per_symbol, int* n_lpf, int* n_rrc);
int wc_c4fm_get_taps(const wc_c4fm* h, float* lpf, float* rrc);
int wc_c4fm_max_symbols(const wc_c4fm* h, int n_samples);
int wc_c4fm_reset(wc_c4fm* h, int channel /* -1 = all; C4FMDemodulator.reset(), c4fm.py:2505 */);
int wc_c4fm_demod(wc_c4fm* h, const void* iq_dev, long long chan_stride, int n_samples, unsigned char* dibits_dev,
                  float* soft_dev, int* n_sym_dev, int max_sym, void* stream);
int wc_c4fm_demod_host(wc_c4fm* h, const void* iq_host /* [n_channels][n_samples] */, int n_samples,
                       unsigned char* dibits_host, float* soft_host, int* n_sym_host, int max_sym);
/* C4FMDemodulator.demodulate_discriminator (c4fm.py:2817-2992): discriminator audio (radians/sample, float32)
 * [C][chan_stride] -> dibits / soft / n_sym as above. first_host float64 [C] = audio[0] per channel in the caller's own
 * precision (scales lfilter_zi on a channel's first call; may be NULL afterwards). Same object state as wc_c4fm_demod. */
int wc_c4fm_demod_disc(wc_c4fm* h, const float* audio_dev, long long chan_stride, int n_samples, const double* first_host,
                       unsigned char* dibits_dev, float* soft_dev, int* n_sym_dev, int max_sym, void* stream);
int wc_c4fm_demod_disc_host(wc_c4fm* h, const float* audio_host, int n_samples, const double* first_host,
                            unsigned char* dibits_host, float* soft_host, int* n_sym_host, int max_sym);
/* stand-alone stages behind the helper classes backend/benchmark_dsp.py:17-114 times:
 * _FMDemodulator.demodulate (c4fm.py:324-395), _Interpolator.filter (:2204-2253), _SoftSyncDetector.process (:2306-2329) */
int wc_c4fm_diffdemod(wc_c4fm* h, const void* iq_pairs_dev /* float2 [C][n] */, int n_samples, float* phases_dev, void* stream);
int wc_c4fm_interp(const float* samples_dev, int n, const int* offsets_dev, const double* mus_dev, int count, double* out_dev,
                   void* stream);
int wc_c4fm_sync_scores(const float* soft_dev, int n, const float* hist24_dev, double* scores_dev, float* new_hist24_dev,
                        void* stream);
/* state8 = {pll (get_timing_offset, c4fm.py:2809), gain, sample_point, buffer_pointer, fine_sync, symbols_since_sync,
 * sync_count, sync events accepted during the last call} */
int wc_c4fm_get_state(wc_c4fm* h, int channel, double* state8);

/* ---- P25 CQPSK / LSM symbol recovery: wavecapsdr/decoders/p25.py:190-669 (CQPSKDemodulator) ----
 * n_channels independent stateful demodulators advanced by one call. lpf_taps63 = _design_baseband_filter
 * (:375-388, float32[63]) and mmse_taps_129x8 = _generate_mmse_taps (:289-323, float32[129*8]); NULL = built inside.
 * demod(): iq complex64 [n_channels][chan_stride] -> dibits uint8 [n_channels][max_sym], n_sym int32 [n_channels]. */
typedef struct wc_cqpsk wc_cqpsk;
int wc_cqpsk_create(int n_channels, int sample_rate, int symbol_rate, const float* lpf_taps63, const float* mmse_taps_129x8,
                    wc_cqpsk** out);
void wc_cqpsk_destroy(wc_cqpsk* h);
int wc_cqpsk_max_symbols(const wc_cqpsk* h, int n_samples);
int wc_cqpsk_reset(wc_cqpsk* h, int channel /* -1 = all */);
int wc_cqpsk_demod(wc_cqpsk* h, const void* iq_dev, long long chan_stride, int n_samples, unsigned char* dibits_dev,
                   int* n_sym_dev, int max_sym, void* stream);
int wc_cqpsk_demod_host(wc_cqpsk* h, const void* iq_host, int n_samples, unsigned char* dibits_host, int* n_sym_host,
                        int max_sym);
/* state6 = {_freq_offset, _phase_acc, _symbol_clock, _symbol_time, _agc_gain, _omega} */
int wc_cqpsk_get_state(wc_cqpsk* h, int channel, double* state6);

/* ---- streaming complex FIR: wavecapsdr/dsp/filters.py:558-668 (fir_filter_complex, fir_decimate) ----
 * y = (taps (*) [zi | x])[::decim] as complex64 (only kept outputs are computed), zi_out = last n_taps-1 entries of
 * [zi | x] as complex128. zi_host NULL = zeros. x_dev/y_dev are device pointers, the rest host; synchronises. */
int wc_fir_complex(const void* x_dev, int n, const double* taps_host, int n_taps, int decim, const void* zi_host,
                   void* y_dev, void* zi_out_host, void* stream);

/* ---- trunking fan-out: phase-continuous NCO + two-stage FIR decimation for K channels of one wideband chunk ----
 * TrunkingSystem.on_raw_iq_callback (trunking/system.py:1434-1466 NCO, :1392-1406 filters, :1753-1779 stages) with
 * init_mode 1 / keep_f64 0, and VoiceRecorder.process_iq (:561-656) with init_mode 2 / keep_f64 1.
 * taps NULL = firwin(157, 0.8/decim1, kaiser 7.857) / firwin(73, 0.8/decim2, kaiser 7.857) designed inside; decim2 = 1
 * disables stage 2. init_mode: first-call filter state 0 = zeros, 1 = lfilter_zi(taps)*x[0] taken as previous INPUTS
 * (what fir_decimate does with the zi it is handed), 2 = scipy lfilter steady state (previous inputs = x[0]).
 * Output per call: [K][out_stride] complex64 (keep_f64 0) or complex128 (1), wc_ddc_out_len(h, n) valid per row;
 * decimation restarts at sample 0 of every call, exactly like `filtered[::D]` per chunk. */
typedef struct wc_ddc wc_ddc;
int wc_ddc_create(int n_channels, int sample_rate, const double* taps1, int n_taps1, int decim1, const double* taps2,
                  int n_taps2, int decim2, int init_mode, int keep_f64, wc_ddc** out);
void wc_ddc_destroy(wc_ddc* h);
int wc_ddc_get_taps(const wc_ddc* h, double* taps1, double* taps2, int* n1, int* n2);
int wc_ddc_set_offsets(wc_ddc* h, const double* offsets_hz /* [K] */);
int wc_ddc_reset(wc_ddc* h, int channel /* -1 = all */);
int wc_ddc_out_len(const wc_ddc* h, int n_samples);
int wc_ddc_process(wc_ddc* h, const void* iq_dev, int n_samples, void* out_dev, long long out_stride, void* stream);
int wc_ddc_process_host(wc_ddc* h, const void* iq_host, int n_samples, void* out_host /* [K][out_len] */);

/* ---- optional audio clean-up of the FM chains (off by default in the reference) ----
 * dsp/filters.py:267-343 noise_blanker: y = x with every sample within +-blanking_width of a sample whose |x| exceeds
 * median(|x|) * 10^(threshold_db/20) set to 0 (untouched when the median is < 1e-10); rows are seq_stride apart.
 * dsp/filters.py:346-459 spectral_noise_reduction with its default geometry (fft 1024, hop 512, periodic Hann):
 * out rows hold wc_spectral_nr_out_len(n) samples (what whole STFT frames cover; n itself when n < 1024). */
int wc_noise_blanker(const float* x_dev, float* y_dev, int n, long long seq_stride, int n_seq, float threshold_db,
                     int blanking_width, void* stream);
int wc_spectral_nr_out_len(int n);
int wc_spectral_nr(const float* x_dev, int n, long long seq_stride, int n_seq, float reduction_db, float* y_dev,
                   long long y_stride, void* stream);

/* ---- output stage: capture.pack_iq16 / pack_pcm16 / pack_f32 (capture.py:102-144) and
 * Channel._update_audio_metrics (capture.py:633-661: sum of squares, peak |x|, count of |x| > 0.95 per sequence) ---- */
int wc_pack(const float* x_dev, void* y_dev, long long total_floats, int fmt /* 0 int16, 1 clipped float32 */, void* stream);
/* Channel.update_signal_metrics (capture.py:749-798), all channels of a chunk in one call: power[c] = sum |freq_shift(iq,
 * offset_c)|^2 (RSSI = 10 log10(power/n + 1e-10)), pct[c] = the two order statistics np.partition picks (ranks n//10 and
 * n - n//10 - 1) — exact radix select; zeros when want_snr is 0 or n is too short (:783). */
int wc_signal_metrics(const void* iq_dev, int fmt, int n, int sample_rate, const double* offsets_hz, int n_ch, int want_snr,
                      float* mag_scratch_dev, double* power_dev, float* pct_dev, void* chan_scratch_dev, void* stream);
int wc_audio_levels(const float* x_dev, int n, long long seq_stride, int n_seq, double* sumsq_dev, float* peak_dev,
                    int* clip_count_dev, void* stream);

/* ---- P25 Phase 1 framing --------------------------------------------------------------------
 * wc_bch_decode: dsp/fec/bch.py:644-658 bch_decode / BCH_63_16_23.decode (:533-641) for `count` codewords at once.
 * bits63 uint8 [count][63] (codeword[0] first, 16 data bits then 47 parity), tracked_nac int32 [count] or NULL
 * (<= 0 = none) -> data int32 [count] (16-bit NAC|DUID, 0 when uncorrectable), errors int32 [count] (-1 = uncorrectable). */
int wc_bch_decode(const unsigned char* bits63_dev, const int* tracked_nac_dev, int count, int* data_dev, int* errors_dev,
                  void* stream);
int wc_bch_decode_host(const unsigned char* bits63_host, const int* tracked_nac_host, int count, int* data_host,
                       int* errors_host);
/* wc_p25framer_*: decoders/p25_framer.py:363-849 P25P1MessageFramer, one stateful framer per channel advanced by one
 * call (soft sync correlation :193-231 with threshold 60, status-symbol stripping, NID BCH decode with the NAC tracker
 * :320-349, message assembly :234-318 and dispatch :690-827). mode 0 = process_batch (:471-509), 1 =
 * process_with_soft_sync per symbol (:438-457), 2 = process (:459-469). dispatch_enabled = started and a listener is
 * set. Inputs: soft float32 / dibits uint8 [C][chan_stride], n_sym int32 [C] or NULL (= n_symbols for every channel).
 * Outputs (device, per channel c): msg_hdr int32 [C][max_msgs][6] = {duid, nac, nbits, corrected_bit_count, offset
 * into the channel's bit pool, 0}; msg_sym int64 [C][max_msgs] = symbols processed when the message was dispatched
 * (timestamp = ref + 1000*that/4800 on the host); msg_bits uint8 [C][pool_bytes] one byte per bit; summary int32
 * [C][8] = {n_msgs, valid NIDs, error code, symbol index of the error (-1), err_a, err_b, err_duid, pool bytes used}.
 * Error codes mirror the AssertionErrors the reference raises out of the batch (the channel stops at that symbol with
 * the reference's partial state): 1 placeholder dispatch, 2 below minimum length (a = length, b = minimum), 3 not
 * aligned to 196-bit blocks, 4 length mismatch, 5 invalid dibit (a = dibit), 6 output buffers full (never with the
 * sizes below). scores (optional, float32 [C][n_symbols]) receives the soft sync scores. */
/* 1/2-rate trellis (Viterbi) decoding: dsp/fec/trellis.py:214-272 TrellisDecoder.decode / trellis_decode for `count`
 * blocks at once (hard decisions, or soft float64 values per dibit), and the TSBK block decode of decoders/p25.py:2037-2109
 * (196 message bits -> deinterleave :2552-2660 -> decode -> 96 bits + last_block / protected / opcode / mfid / payload). */
int wc_trellis12_decode(const unsigned char* dibits_dev, long long stride, const int* n_dibits_dev, int n_fixed,
                        const double* soft_dev, int count, unsigned char* out_dev, long long out_stride, int* n_out_dev,
                        int* metric_dev, void* stream);
int wc_tsbk_decode(const unsigned char* bits196_dev, int count, unsigned char* bits96_dev, int* metric_dev, int* fields_dev,
                   unsigned char* data8_dev, void* stream);
int wc_tsbk_decode_host(const unsigned char* bits196_host, int count, unsigned char* bits96_host, int* metric_host,
                        int* fields_host, unsigned char* data8_host);
typedef struct wc_p25framer wc_p25framer;
int wc_p25framer_create(int n_channels, wc_p25framer** out);
void wc_p25framer_destroy(wc_p25framer* h);
int wc_p25framer_reset(wc_p25framer* h, int channel /* -1 = all */, int full /* 1 = also NAC tracker + symbol clock */);
int wc_p25framer_max_msgs(int n_symbols);
int wc_p25framer_pool_bytes(int n_symbols);
int wc_p25framer_process(wc_p25framer* h, const float* soft_dev, const unsigned char* dibits_dev, long long chan_stride,
                         const int* n_sym_dev, int n_symbols, int mode, int dispatch_enabled, float* scores_dev,
                         int* msg_hdr_dev, long long* msg_sym_dev, unsigned char* msg_bits_dev, int* summary_dev,
                         void* stream);
int wc_p25framer_process_host(wc_p25framer* h, const float* soft_host, const unsigned char* dibits_host, int n_symbols,
                              const int* n_sym_host, int mode, int dispatch_enabled, float* scores_host, int* msg_hdr_host,
                              long long* msg_sym_host, unsigned char* msg_bits_host, int* summary_host);
int wc_p25framer_get_state(wc_p25framer* h, int channel, int* state12);

/* ---- voice-channel discriminator path -----------------------------------------------------------
 * wc_fm_discriminator: trunking/system.py:708-717 (VoiceRecorder.process_iq) np.diff(np.unwrap([last | np.angle(iq)])):
 * iq complex64 (is_f64 0) / complex128 (1) [C][chan_stride] -> out float64 [C][n_samples]; last_phase float64 [C] in/out.
 * wc_discdemod_*: decoders/p25.py:1105-1345 DiscriminatorDemodulator, one stateful demodulator per channel advanced by
 * one call: audio float32 [C][chan_stride] -> dibits uint8 [C][max_sym] (+ the slicer input `output` as soft float32
 * [C][max_sym], optional), n_sym int32 [C]. mmse_taps_129x8 / lpf_taps65: the reference's float32 tables
 * (_generate_mmse_taps :1165-1186, _design_baseband_filter :1188-1195); NULL mmse taps = built inside. reset()
 * (:1335-1345) keeps the input gain. */
int wc_fm_discriminator(const void* iq_dev, int is_f64, long long chan_stride, int n_samples, int n_channels,
                        double* last_phase_dev, double* out_dev, void* stream);
typedef struct wc_discdemod wc_discdemod;
int wc_discdemod_create(int n_channels, int sample_rate, int symbol_rate, const float* mmse_taps_129x8,
                        const float* lpf_taps65, wc_discdemod** out);
void wc_discdemod_destroy(wc_discdemod* h);
int wc_discdemod_reset(wc_discdemod* h, int channel /* -1 = all */);
int wc_discdemod_max_symbols(const wc_discdemod* h, int n_samples);
int wc_discdemod_get_taps(const wc_discdemod* h, float* mmse_taps_129x8);
int wc_discdemod_demod(wc_discdemod* h, const float* audio_dev, long long chan_stride, int n_samples,
                       unsigned char* dibits_dev, float* soft_dev, int* n_sym_dev, int max_sym, void* stream);
int wc_discdemod_demod_host(wc_discdemod* h, const float* audio_host, int n_samples, unsigned char* dibits_host,
                            float* soft_host, int* n_sym_host, int max_sym);
int wc_discdemod_get_state(wc_discdemod* h, int channel, double* state8);

/* ---- control-channel scanner: trunking/cc_scanner.py:166-351 ControlChannelScanner._measure_channel /
 * _detect_sync_pattern for n_ch frequency offsets (candidates and the two band-edge noise probes) of one wideband block.
 * iq complex64 [n]; taps65 = firwin(65, 0.8/D, kaiser 6.0), D = max(1, fs // 48000) (host); y complex128 [n_ch][m] with
 * m = wc_ccscan_out_len; power_sum / power_max float64 [n_ch] = sum and max of |y|^2; corr float64 [n_ch] = best
 * normalised sync correlation (0 when the block is shorter than 250 kept samples). scratch_dev: 8 * n_ch bytes. */
int wc_ccscan_out_len(int n_samples, int sample_rate);
int wc_ccscan_measure(const void* iq_dev, int n_samples, int sample_rate, const double* offsets_hz, int n_ch,
                      const double* taps65, void* y_dev, double* power_sum_dev, double* power_max_dev, double* corr_dev,
                      void* scratch_dev, void* stream);

/* ---- peer memory: ONE capture fanned out over the GPUs of a box (SURVEY §8e modes ii/iii; the reference's analogue is
 * Capture._run_thread handing the SAME `samples` array to every channel worker, capture.py:2541-2546). The ingest
 * rank allocates the IQ region with wc_peer_alloc and ships the 64-byte handle to the other processes (any transport);
 * they map it with wc_peer_open and pass pointers into it to wc_chan_process / wc_front_run / ... like any device
 * pointer: the kernels then pull their slab over NVLink themselves (no broadcast, no staging copy). wc_flag_set /
 * wc_flag_wait order producer and consumers on the stream: sequence numbers in uint32 words of the same region,
 * system-scope release/acquire; a wait gives up after timeout_ms and sets *timed_out_dev (int32, optional) to 1. */
#define WC_PEER_HANDLE_BYTES 64
int wc_peer_alloc(long long bytes, void** dev_out, void* handle_out /* WC_PEER_HANDLE_BYTES */);
int wc_peer_free(void* dev);
int wc_peer_open(const void* handle, void** dev_out);
int wc_peer_close(void* dev);
/* staging alternative / cross-check: plain async copy between any two device addresses (local or mapped) */
int wc_peer_copy(void* dst_dev, const void* src_dev, long long bytes, void* stream);
int wc_flag_set(unsigned* flag_dev, unsigned value, void* stream);
/* waits until flags_dev[i * stride_words] >= value (wrap-safe) for every i < n_flags (<= 1024) */
int wc_flag_wait(const unsigned* flags_dev, int n_flags, long long stride_words, unsigned value, int timeout_ms,
                 int* timed_out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WCSDR_B200_H */
