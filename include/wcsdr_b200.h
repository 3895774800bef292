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
/* same, host buffers: H2D copy + kernels + D2H copy + sync (the reference-facing call). */
int wc_chan_process_host(wc_chan* h, const void* iq_host, long long n_samples, int n_chunks, int mode,
                         float fm_scale, void* out_host);

#ifdef __cplusplus
}
#endif
#endif /* WCSDR_B200_H */
