// Streaming complex FIR / FIR-decimate and the trunking fan-out (NCO + two-stage decimation).
//
//  * wc_fir_complex  — wavecapsdr/dsp/filters.py:558-668 `fir_filter_complex` / `fir_decimate`:
//        y[i] = sum_j taps[j] * state[n_zi + i - j],  state = [zi | x]  (complex128 x float64),
//        y -> complex64, decimation = y[::D] restarted at index 0 of every call, new zi = last n_zi
//        entries of state. Only the kept outputs are computed.
//  * wc_ddc_*        — the per-channel work of TrunkingSystem.on_raw_iq_callback
//        (trunking/system.py:1434-1466 phase-continuous NCO, :1392-1406 Kaiser(7.857) 157/73-tap filter
//        design, :1753-1779 two fir_decimate stages) and of VoiceRecorder.process_iq (:561-656, same
//        chain with scipy lfilter and complex128 between the stages), for K channels of ONE wideband
//        chunk in one pass: every CTA stages a raw IQ tile in shared memory once and loops over the
//        channels (mix with exp(-j 2 pi off n / fs), 157-tap stage-1 outputs), so HBM sees each IQ
//        byte once per chunk no matter how many channels are extracted.
//
// NCO: phase = -2 pi * off * n / fs with n = sample_idx + i kept in float64 like the reference; the
// turn count off*n is reduced modulo fs exactly (fmod) before sincospi, so the phasor is accurate to
// float64 rounding, then rounded to complex64 (the reference casts the exponential to complex64) and
// multiplied in float32.
#include <math.h>
#include <string.h>
#include <vector>

#include "../../include/wcsdr_b200.h"
#include "common.cuh"

namespace wc {

// ---------------------------------------------------------------------------------------------
// generic batched FIR-decimate: x [K][n] (float2 or double2), hist [K][T-1] double2 (oldest first)
// ---------------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(128) fir_dec_kernel(const TIn* __restrict__ x, long long x_stride, int n,
                                                      const double2* __restrict__ hist, const double* __restrict__ taps, int T,
                                                      int D, TOut* __restrict__ y, long long y_stride, int n_out) {
    extern __shared__ double s_taps[];
    for (int i = threadIdx.x; i < T; i += blockDim.x) s_taps[i] = taps[i];
    __syncthreads();
    const int k = blockIdx.y;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_out) return;
    const TIn* xc = x + (long long)k * x_stride;
    const double2* hc = hist + (long long)k * (T - 1);
    const int i = m * D;
    double ar = 0.0, ai = 0.0;
    for (int j = 0; j < T; ++j) {
        const int g = i - j;
        double vx, vy;
        if (g >= 0) {
            const TIn v = xc[g];
            vx = (double)v.x;
            vy = (double)v.y;
        } else {
            const double2 v = hc[(T - 1) + g];
            vx = v.x;
            vy = v.y;
        }
        ar = fma(s_taps[j], vx, ar);
        ai = fma(s_taps[j], vy, ai);
    }
    TOut o;
    o.x = ar;
    o.y = ai;
    y[(long long)k * y_stride + m] = o;
}

// new_hist = last (T-1) entries of [old_hist | x]
template <typename TIn>
__global__ void fir_hist_kernel(const TIn* __restrict__ x, long long x_stride, int n, const double2* __restrict__ old_h,
                                double2* __restrict__ new_h, int hl) {
    const int k = blockIdx.x;
    for (int i = threadIdx.x; i < hl; i += blockDim.x) {
        const int g = n - hl + i;
        double2 v;
        if (g >= 0) {
            const TIn s = x[(long long)k * x_stride + g];
            v = make_double2((double)s.x, (double)s.y);
        } else {
            v = old_h[(long long)k * hl + (hl + g)];
        }
        new_h[(long long)k * hl + i] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// DDC stage 1: NCO + FIR-decimate for all channels from one staged tile
// ---------------------------------------------------------------------------------------------
constexpr int DDC_THREADS = 256;
constexpr int DDC_OUT_TILE = 64;      // max stage-1 outputs per CTA per channel (4 threads each)
constexpr int DDC_MAX_T1 = 256;
constexpr int DDC_MAX_SPAN = 2560;    // raw samples staged per CTA (>= (out_tile - 1) * D1 + T1); 2 x 20 KB of shared memory

struct DdcChan {
    double offset_hz;
    long long sample_idx;   // index of the first sample of this call
    int shift;              // 0: offset == 0, samples pass through untouched
    int pad;
};

struct DdcArgs {
    const float2* x;        // [n] wideband chunk
    int n;
    int K;
    double fs;
    const DdcChan* ch;      // [K]
    const double2* hist;    // [K][T1-1] previous mixed samples (complex128)
    const double* taps;     // [T1]
    int T1, D1;
    void* y;                // [K][y_stride] float2 or double2
    long long y_stride;
    int n1;                 // stage-1 outputs per channel
    int out_f64;
    int ch_per_cta;
    int out_tile;           // stage-1 outputs per CTA: min(64, (DDC_MAX_SPAN - T1) / D1 + 1)
    int span_max;           // staged samples per CTA, (out_tile - 1) * D1 + T1 rounded up to even
};

__device__ __forceinline__ float2 ddc_mix(float2 s, const DdcChan& c, int g, double fs) {
    if (!c.shift) return s;
    // phase = -2 pi * off * (sample_idx + g) / fs, wanted as np.exp(1j*phase).astype(complex64). The turn count is
    // reduced in float64 (four FP64 operations; |error| < 1e-9 turn), the sine/cosine are evaluated in float32 on the
    // leading part and corrected to first order for the float32 remainder of the reduced phase — a float64 fmod +
    // sincospi per (sample, channel) made this kernel FP64-pipe bound (96 channels: 0.85 ms per 50 ms chunk).
    const double nn = (double)(c.sample_idx + (long long)g);
    double t = (c.offset_hz * nn) * (1.0 / fs);     // the reciprocal is loop invariant where this is inlined
    t -= rint(t);                                   // turns in [-0.5, 0.5]
    const float th = (float)t;
    const float tl = (float)(t - (double)th);
    float sn, cs;
    sincospif(-2.0f * th, &sn, &cs);
    const float d = -6.283185307179586f * tl;       // remaining angle, |d| < 2e-7
    const float er = fmaf(-sn, d, cs), ei = fmaf(cs, d, sn);
    return make_float2(__fsub_rn(__fmul_rn(s.x, er), __fmul_rn(s.y, ei)), __fadd_rn(__fmul_rn(s.x, ei), __fmul_rn(s.y, er)));
}

// Dynamic shared memory: mixed samples as complex128 [span_max] | raw samples complex64 [span_max] | taps float64 [T1].
// The mixed samples are widened once when they are written (each feeds T1/D1 ~ 5 outputs), not once per tap: the FIR
// loop is two DFMAs per tap with no conversion — same values, 40 % fewer FP64-pipe operations.
__global__ void __launch_bounds__(DDC_THREADS) ddc_stage1_kernel(const DdcArgs a) {
    extern __shared__ __align__(16) unsigned char ddc_smem[];
    double2* mix = reinterpret_cast<double2*>(ddc_smem);
    float2* raw = reinterpret_cast<float2*>(mix + a.span_max);
    double* s_taps = reinterpret_cast<double*>(raw + a.span_max);
    const int m0 = blockIdx.x * a.out_tile;          // first stage-1 output of this tile
    const int hl = a.T1 - 1;
    const int g0 = m0 * a.D1 - hl;                   // first raw sample needed
    const int n_out = min(a.out_tile, a.n1 - m0);
    const int span = (n_out - 1) * a.D1 + a.T1;      // raw samples covering the tile's outputs
    for (int i = threadIdx.x; i < span; i += DDC_THREADS) {
        const int g = g0 + i;
        raw[i] = (g >= 0 && g < a.n) ? a.x[g] : make_float2(0.f, 0.f);
    }
    for (int i = threadIdx.x; i < a.T1; i += DDC_THREADS) s_taps[i] = a.taps[i];
    __syncthreads();
    const int k_lo = blockIdx.y * a.ch_per_cta, k_hi = min(a.K, k_lo + a.ch_per_cta);
    // 4 threads per output, each sums a quarter of the taps
    const int o = threadIdx.x >> 2, part = threadIdx.x & 3;
    for (int k = k_lo; k < k_hi; ++k) {
        const DdcChan c = a.ch[k];
        const double2* hc = a.hist + (long long)k * hl;
        for (int i = threadIdx.x; i < span; i += DDC_THREADS) {
            const int g = g0 + i;
            const float2 m = (g >= 0) ? ddc_mix(raw[i], c, g, a.fs) : make_float2(0.f, 0.f);
            mix[i] = make_double2((double)m.x, (double)m.y);
        }
        __syncthreads();
        double ar = 0.0, ai = 0.0;
        if (o < n_out) {
            // output m = m0 + o sits at raw index i0 = o*D1 + hl (tile-relative); tap j reads i0 - j
            const int i0 = o * a.D1 + hl;
            for (int j = part; j < a.T1; j += 4) {
                const int i = i0 - j;
                const int g = g0 + i;
                const double2 v = (g >= 0) ? mix[i] : hc[hl + g];
                ar = fma(s_taps[j], v.x, ar);
                ai = fma(s_taps[j], v.y, ai);
            }
        }
        ar += __shfl_xor_sync(0xffffffffu, ar, 1);
        ai += __shfl_xor_sync(0xffffffffu, ai, 1);
        ar += __shfl_xor_sync(0xffffffffu, ar, 2);
        ai += __shfl_xor_sync(0xffffffffu, ai, 2);
        if (o < n_out && part == 0) {
            const long long idx = (long long)k * a.y_stride + m0 + o;
            if (a.out_f64) reinterpret_cast<double2*>(a.y)[idx] = make_double2(ar, ai);
            else reinterpret_cast<float2*>(a.y)[idx] = make_float2((float)ar, (float)ai);
        }
        __syncthreads();
    }
}

// stage-1 history: last T1-1 mixed samples of [old_hist | mixed x]; first call: zi initialisation
//   init_mode 1: hist[j] = template[j] * mixed[0]   (fir_decimate fed lfilter_zi*x[0] as "previous inputs",
//                                                   trunking/system.py:1756-1761 with dsp/filters.py:540-551)
//   init_mode 2: hist[j] = mixed[0]                 (scipy lfilter steady state, system.py:633-637)
__global__ void ddc_hist1_kernel(const float2* __restrict__ x, int n, double fs, const DdcChan* __restrict__ ch,
                                 const double2* __restrict__ old_h, double2* __restrict__ new_h, int hl) {
    const int k = blockIdx.x;
    const DdcChan c = ch[k];
    for (int i = threadIdx.x; i < hl; i += blockDim.x) {
        const int g = n - hl + i;
        double2 v;
        if (g >= 0) {
            const float2 s = ddc_mix(x[g], c, g, fs);
            v = make_double2((double)s.x, (double)s.y);
        } else {
            v = old_h[(long long)k * hl + (hl + g)];
        }
        new_h[(long long)k * hl + i] = v;
    }
}

__global__ void ddc_init1_kernel(const float2* __restrict__ x, double fs, const DdcChan* __restrict__ ch,
                                 const unsigned char* __restrict__ need, const double* __restrict__ tmpl, int init_mode,
                                 double2* __restrict__ h, int hl) {
    const int k = blockIdx.x;
    if (!need[k]) return;
    const float2 s = ddc_mix(x[0], ch[k], 0, fs);
    for (int i = threadIdx.x; i < hl; i += blockDim.x) {
        const double t = (init_mode == 1) ? tmpl[i] : 1.0;
        h[(long long)k * hl + i] = (init_mode == 0) ? make_double2(0.0, 0.0) : make_double2(t * (double)s.x, t * (double)s.y);
    }
}

template <typename TIn>
__global__ void ddc_init2_kernel(const TIn* __restrict__ y1, long long stride, const unsigned char* __restrict__ need,
                                 const double* __restrict__ tmpl, int init_mode, double2* __restrict__ h, int hl) {
    const int k = blockIdx.x;
    if (!need[k]) return;
    const TIn s = y1[(long long)k * stride];
    for (int i = threadIdx.x; i < hl; i += blockDim.x) {
        const double t = (init_mode == 1) ? tmpl[i] : 1.0;
        h[(long long)k * hl + i] = (init_mode == 0) ? make_double2(0.0, 0.0) : make_double2(t * (double)s.x, t * (double)s.y);
    }
}

__global__ void zero_hist_kernel(double2* h0, double2* h1, int hl, int lo) {
    const int k = lo + blockIdx.x;
    for (int i = threadIdx.x; i < hl; i += blockDim.x) {
        h0[(long long)k * hl + i] = make_double2(0.0, 0.0);
        h1[(long long)k * hl + i] = make_double2(0.0, 0.0);
    }
}

// scipy.signal.lfilter_zi(b, 1.0) for an FIR: zi[k] = sum_{j > k} b[j]
static void fir_lfilter_zi(const std::vector<double>& b, std::vector<double>& zi) {
    const int T = (int)b.size();
    zi.assign(T > 1 ? T - 1 : 0, 0.0);
    double s = 0.0;
    for (int k = T - 2; k >= 0; --k) {
        s += b[k + 1];
        zi[k] = s;
    }
}

void firwin_kaiser_lowpass(int numtaps, double cutoff, double beta, std::vector<double>& h);  // channelizer.cu

}  // namespace wc

using namespace wc;

struct wc_ddc {
    int K = 0, fs = 0, T1 = 0, D1 = 1, T2 = 0, D2 = 1, init_mode = 1, keep_f64 = 0;
    std::vector<double> taps1, taps2, tmpl1, tmpl2;
    double *d_taps1 = nullptr, *d_taps2 = nullptr, *d_tmpl1 = nullptr, *d_tmpl2 = nullptr;
    double2* d_hist1[2] = {nullptr, nullptr};
    double2* d_hist2[2] = {nullptr, nullptr};
    int cur = 0;
    std::vector<DdcChan> chan;          // host copy of the NCO state
    std::vector<double> last_offset;
    std::vector<unsigned char> need1, need2;
    DdcChan* d_chan = nullptr;
    unsigned char *d_need1 = nullptr, *d_need2 = nullptr;
    void* d_y1 = nullptr; size_t y1_bytes = 0;
    void* d_in = nullptr; size_t in_bytes = 0;
    void* d_out = nullptr; size_t out_bytes = 0;
    cudaStream_t stream = nullptr;
};

static int grow(void** p, size_t* cap, size_t need) {
    if (*cap >= need) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    WC_CUDA(cudaMalloc(p, need));
    *cap = need;
    return 0;
}

extern "C" {

/* ---- stateless functional form: fir_filter_complex / fir_decimate ---- */
int wc_fir_complex(const void* x_dev, int n, const double* taps_host, int n_taps, int decim, const void* zi_host,
                   void* y_dev, void* zi_out_host, void* stream_v) {
    WC_REQUIRE(x_dev && taps_host && y_dev, "wc_fir_complex: null argument");
    WC_REQUIRE(n >= 1 && n_taps >= 1 && n_taps <= 4096 && decim >= 1, "wc_fir_complex: bad sizes");
    cudaStream_t s = (cudaStream_t)stream_v;
    const int hl = n_taps - 1;
    // one stream-ordered scratch block: taps | history in | history out; released on every exit path
    const size_t hbytes = sizeof(double2) * (size_t)(hl > 0 ? hl : 1);
    const size_t tbytes = (sizeof(double) * (size_t)n_taps + 15) & ~size_t(15);
    struct Scratch {
        void* p = nullptr;
        cudaStream_t s;
        ~Scratch() {
            if (p) cudaFreeAsync(p, s);
        }
    } ws;
    ws.s = s;
    WC_CUDA(cudaMallocAsync(&ws.p, tbytes + 2 * hbytes, s));
    double* d_taps = reinterpret_cast<double*>(ws.p);
    double2* d_h = reinterpret_cast<double2*>(reinterpret_cast<char*>(ws.p) + tbytes);
    double2* d_h2 = reinterpret_cast<double2*>(reinterpret_cast<char*>(ws.p) + tbytes + hbytes);
    WC_CUDA(cudaMemcpyAsync(d_taps, taps_host, sizeof(double) * n_taps, cudaMemcpyHostToDevice, s));
    if (hl > 0) {
        if (zi_host) WC_CUDA(cudaMemcpyAsync(d_h, zi_host, sizeof(double2) * hl, cudaMemcpyHostToDevice, s));
        else WC_CUDA(cudaMemsetAsync(d_h, 0, sizeof(double2) * hl, s));
    }
    const int n_out = (n + decim - 1) / decim;
    dim3 g((n_out + 127) / 128, 1);
    fir_dec_kernel<float2, float2><<<g, 128, sizeof(double) * n_taps, s>>>(reinterpret_cast<const float2*>(x_dev), n, n, d_h,
                                                                         d_taps, n_taps, decim,
                                                                         reinterpret_cast<float2*>(y_dev), n_out, n_out);
    if (hl > 0 && zi_out_host) {
        fir_hist_kernel<float2><<<1, 256, 0, s>>>(reinterpret_cast<const float2*>(x_dev), n, n, d_h, d_h2, hl);
        WC_CUDA(cudaMemcpyAsync(zi_out_host, d_h2, sizeof(double2) * hl, cudaMemcpyDeviceToHost, s));
    }
    WC_CUDA(cudaGetLastError());
    // taps_host / zi_host are pageable caller memory and zi_out_host is read by the caller on return
    WC_CUDA(cudaStreamSynchronize(s));
    return 0;
}

/* ---- trunking fan-out ---- */
int wc_ddc_create(int n_channels, int sample_rate, const double* taps1, int n_taps1, int decim1, const double* taps2,
                  int n_taps2, int decim2, int init_mode, int keep_f64, wc_ddc** out) {
    WC_REQUIRE(out != nullptr, "wc_ddc_create: out is null");
    WC_REQUIRE(n_channels >= 1 && sample_rate > 0 && decim1 >= 1 && decim2 >= 1, "wc_ddc_create: bad parameters");
    WC_REQUIRE(init_mode >= 0 && init_mode <= 2, "wc_ddc_create: init_mode must be 0 (zeros), 1 (fir_decimate) or 2 (lfilter)");
    wc_ddc* h = new wc_ddc();
    h->K = n_channels;
    h->fs = sample_rate;
    h->D1 = decim1;
    h->D2 = decim2;
    h->init_mode = init_mode;
    h->keep_f64 = keep_f64 ? 1 : 0;
    // trunking/system.py:1392-1406: firwin(157, 0.8/D1, kaiser 7.857), firwin(73, 0.8/D2, kaiser 7.857)
    if (taps1 && n_taps1 > 0) h->taps1.assign(taps1, taps1 + n_taps1);
    else firwin_kaiser_lowpass(157, 0.8 / decim1, 7.857, h->taps1);
    if (decim2 > 1) {
        if (taps2 && n_taps2 > 0) h->taps2.assign(taps2, taps2 + n_taps2);
        else firwin_kaiser_lowpass(73, 0.8 / decim2, 7.857, h->taps2);
    }
    h->T1 = (int)h->taps1.size();
    h->T2 = (int)h->taps2.size();
    if (h->T1 > DDC_MAX_T1 || h->T1 > DDC_MAX_SPAN || h->T1 < 2 || (decim2 > 1 && h->T2 < 2)) {
        set_error("wc_ddc_create: stage-1 filter (%d taps, /%d) exceeds the tile (%d taps, span %d)", h->T1, decim1, DDC_MAX_T1,
                  DDC_MAX_SPAN);
        delete h;
        return -1;
    }
    fir_lfilter_zi(h->taps1, h->tmpl1);
    fir_lfilter_zi(h->taps2, h->tmpl2);
    h->chan.assign(n_channels, DdcChan{0.0, 0, 0, 0});
    h->last_offset.assign(n_channels, 0.0);
    h->need1.assign(n_channels, 1);
    h->need2.assign(n_channels, 1);
    const size_t K = (size_t)n_channels;
    const int hl1 = h->T1 - 1, hl2 = h->T2 > 0 ? h->T2 - 1 : 1;
    bool ok = cudaMalloc((void**)&h->d_taps1, sizeof(double) * h->T1) == cudaSuccess &&
              cudaMalloc((void**)&h->d_tmpl1, sizeof(double) * hl1) == cudaSuccess &&
              cudaMalloc((void**)&h->d_taps2, sizeof(double) * (h->T2 > 0 ? h->T2 : 1)) == cudaSuccess &&
              cudaMalloc((void**)&h->d_tmpl2, sizeof(double) * hl2) == cudaSuccess &&
              cudaMalloc((void**)&h->d_hist1[0], sizeof(double2) * hl1 * K) == cudaSuccess &&
              cudaMalloc((void**)&h->d_hist1[1], sizeof(double2) * hl1 * K) == cudaSuccess &&
              cudaMalloc((void**)&h->d_hist2[0], sizeof(double2) * hl2 * K) == cudaSuccess &&
              cudaMalloc((void**)&h->d_hist2[1], sizeof(double2) * hl2 * K) == cudaSuccess &&
              cudaMalloc((void**)&h->d_chan, sizeof(DdcChan) * K) == cudaSuccess &&
              cudaMalloc((void**)&h->d_need1, K) == cudaSuccess && cudaMalloc((void**)&h->d_need2, K) == cudaSuccess &&
              cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
    if (ok) {
        cudaMemcpy(h->d_taps1, h->taps1.data(), sizeof(double) * h->T1, cudaMemcpyHostToDevice);
        cudaMemcpy(h->d_tmpl1, h->tmpl1.data(), sizeof(double) * hl1, cudaMemcpyHostToDevice);
        if (h->T2 > 0) {
            cudaMemcpy(h->d_taps2, h->taps2.data(), sizeof(double) * h->T2, cudaMemcpyHostToDevice);
            cudaMemcpy(h->d_tmpl2, h->tmpl2.data(), sizeof(double) * (h->T2 - 1), cudaMemcpyHostToDevice);
        }
        for (int b = 0; b < 2; ++b) {
            cudaMemset(h->d_hist1[b], 0, sizeof(double2) * hl1 * K);
            cudaMemset(h->d_hist2[b], 0, sizeof(double2) * hl2 * K);
        }
        ok = cudaGetLastError() == cudaSuccess;
    }
    if (!ok) {
        set_error("wc_ddc_create: CUDA setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete h;
        return -2;
    }
    *out = h;
    return 0;
}

void wc_ddc_destroy(wc_ddc* h) {
    if (!h) return;
    cudaFree(h->d_taps1);
    cudaFree(h->d_taps2);
    cudaFree(h->d_tmpl1);
    cudaFree(h->d_tmpl2);
    for (int b = 0; b < 2; ++b) {
        cudaFree(h->d_hist1[b]);
        cudaFree(h->d_hist2[b]);
    }
    cudaFree(h->d_chan);
    cudaFree(h->d_need1);
    cudaFree(h->d_need2);
    if (h->d_y1) cudaFree(h->d_y1);
    if (h->d_in) cudaFree(h->d_in);
    if (h->d_out) cudaFree(h->d_out);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int wc_ddc_get_taps(const wc_ddc* h, double* taps1, double* taps2, int* n1, int* n2) {
    WC_REQUIRE(h != nullptr, "wc_ddc_get_taps: null handle");
    if (n1) *n1 = h->T1;
    if (n2) *n2 = h->T2;
    if (taps1) memcpy(taps1, h->taps1.data(), sizeof(double) * h->T1);
    if (taps2 && h->T2) memcpy(taps2, h->taps2.data(), sizeof(double) * h->T2);
    return 0;
}

/* offsets as passed to phase_continuous_freq_shift: a changed offset restarts that channel's phase at sample 0
 * (system.py:1447-1449); filter states are kept (the reference keeps them across hunting, too). */
int wc_ddc_set_offsets(wc_ddc* h, const double* offsets_hz) {
    WC_REQUIRE(h && offsets_hz, "wc_ddc_set_offsets: null argument");
    for (int k = 0; k < h->K; ++k) h->chan[k].offset_hz = offsets_hz[k];
    return 0;
}

int wc_ddc_reset(wc_ddc* h, int channel) {
    WC_REQUIRE(h != nullptr, "wc_ddc_reset: null handle");
    WC_REQUIRE(channel >= -1 && channel < h->K, "wc_ddc_reset: channel %d out of range", channel);
    const int lo = channel < 0 ? 0 : channel, hi = channel < 0 ? h->K : channel + 1;
    for (int k = lo; k < hi; ++k) {
        h->chan[k].sample_idx = 0;
        h->last_offset[k] = 0.0;
        h->need1[k] = 1;
        h->need2[k] = 1;
    }
    zero_hist_kernel<<<hi - lo, 128, 0, h->stream>>>(h->d_hist1[0], h->d_hist1[1], h->T1 - 1, lo);
    if (h->T2 > 1) zero_hist_kernel<<<hi - lo, 128, 0, h->stream>>>(h->d_hist2[0], h->d_hist2[1], h->T2 - 1, lo);
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

/* output lengths for a call of n samples: n1 = ceil(n / D1) stage-1 samples, ceil(n1 / D2) outputs per channel */
int wc_ddc_out_len(const wc_ddc* h, int n_samples) {
    if (!h || n_samples <= 0) return 0;
    const int n1 = (n_samples + h->D1 - 1) / h->D1;
    return h->D2 > 1 ? (n1 + h->D2 - 1) / h->D2 : n1;
}

int wc_ddc_process(wc_ddc* h, const void* iq_dev, int n_samples, void* out_dev, long long out_stride, void* stream_v) {
    WC_REQUIRE(h && iq_dev && out_dev, "wc_ddc_process: null argument");
    if (n_samples <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream_v;
    const int K = h->K;
    const int n1 = (n_samples + h->D1 - 1) / h->D1;
    const int n2 = wc_ddc_out_len(h, n_samples);
    WC_REQUIRE(out_stride >= n2, "wc_ddc_process: out_stride %lld < %d", out_stride, n2);
    // NCO bookkeeping (system.py:1443-1466)
    for (int k = 0; k < K; ++k) {
        DdcChan& c = h->chan[k];
        c.shift = (c.offset_hz != 0.0) ? 1 : 0;
        if (c.shift && c.offset_hz != h->last_offset[k]) {
            c.sample_idx = 0;
            h->last_offset[k] = c.offset_hz;
        }
    }
    WC_CUDA(cudaMemcpyAsync(h->d_chan, h->chan.data(), sizeof(DdcChan) * K, cudaMemcpyHostToDevice, s));
    WC_CUDA(cudaMemcpyAsync(h->d_need1, h->need1.data(), K, cudaMemcpyHostToDevice, s));
    WC_CUDA(cudaMemcpyAsync(h->d_need2, h->need2.data(), K, cudaMemcpyHostToDevice, s));
    const float2* x = reinterpret_cast<const float2*>(iq_dev);
    const int hl1 = h->T1 - 1;
    ddc_init1_kernel<<<K, 128, 0, s>>>(x, (double)h->fs, h->d_chan, h->d_need1, h->d_tmpl1, h->init_mode, h->d_hist1[h->cur], hl1);
    const bool two = h->D2 > 1;
    const bool y1_f64 = h->keep_f64 != 0;
    const size_t esz1 = y1_f64 ? sizeof(double2) : sizeof(float2);
    void* y1 = out_dev;
    long long y1_stride = out_stride;
    if (two) {
        if (grow(&h->d_y1, &h->y1_bytes, esz1 * (size_t)K * n1)) return -2;
        y1 = h->d_y1;
        y1_stride = n1;
    }
    DdcArgs a;
    a.x = x;
    a.n = n_samples;
    a.K = K;
    a.fs = (double)h->fs;
    a.ch = h->d_chan;
    a.hist = h->d_hist1[h->cur];
    a.taps = h->d_taps1;
    a.T1 = h->T1;
    a.D1 = h->D1;
    a.y = y1;
    a.y_stride = y1_stride;
    a.n1 = n1;
    a.out_f64 = y1_f64 ? 1 : 0;
    a.out_tile = (DDC_MAX_SPAN - h->T1) / h->D1 + 1;
    if (a.out_tile > DDC_OUT_TILE) a.out_tile = DDC_OUT_TILE;
    const int tiles = (n1 + a.out_tile - 1) / a.out_tile;
    // enough CTAs to fill the machine: split the channel loop when there are few tiles
    int groups = 1;
    // ncu (96 channels): with 2 CTAs per SM the kernel ran at 26 % of the warp slots, half a wave of its 4 resident CTAs per
    // SM; the split now aims at four full waves (two: 0.323 ms, four: 0.301 ms at 96 channels) (the raw tile is re-staged per group, 1/12 of a group's work at 96 channels)
    while (tiles * groups < 16 * sm_count() && groups < K) groups *= 2;
    a.ch_per_cta = (K + groups - 1) / groups;
    groups = (K + a.ch_per_cta - 1) / a.ch_per_cta;
    a.span_max = (((a.out_tile - 1) * h->D1 + h->T1) + 1) & ~1;
    const size_t smem1 = (sizeof(double2) + sizeof(float2)) * (size_t)a.span_max + sizeof(double) * (size_t)h->T1;
    static std::atomic<unsigned long long> attr_done{0};
    WC_CUDA(smem_optin(ddc_stage1_kernel, (int)((sizeof(double2) + sizeof(float2)) * DDC_MAX_SPAN + sizeof(double) * DDC_MAX_T1), attr_done));
    ddc_stage1_kernel<<<dim3(tiles, groups), DDC_THREADS, smem1, s>>>(a);
    ddc_hist1_kernel<<<K, 128, 0, s>>>(x, n_samples, (double)h->fs, h->d_chan, h->d_hist1[h->cur], h->d_hist1[h->cur ^ 1], hl1);
    if (two) {
        const int hl2 = h->T2 - 1;
        dim3 g2((n2 + 127) / 128, K);
        const size_t sm = sizeof(double) * h->T2;
        if (y1_f64) {
            ddc_init2_kernel<double2><<<K, 128, 0, s>>>(reinterpret_cast<const double2*>(y1), n1, h->d_need2, h->d_tmpl2,
                                                       h->init_mode, h->d_hist2[h->cur], hl2);
            fir_dec_kernel<double2, double2><<<g2, 128, sm, s>>>(reinterpret_cast<const double2*>(y1), n1, n1, h->d_hist2[h->cur],
                                                                 h->d_taps2, h->T2, h->D2, reinterpret_cast<double2*>(out_dev),
                                                                 out_stride, n2);
            fir_hist_kernel<double2><<<K, 128, 0, s>>>(reinterpret_cast<const double2*>(y1), n1, n1, h->d_hist2[h->cur],
                                                      h->d_hist2[h->cur ^ 1], hl2);
        } else {
            ddc_init2_kernel<float2><<<K, 128, 0, s>>>(reinterpret_cast<const float2*>(y1), n1, h->d_need2, h->d_tmpl2,
                                                      h->init_mode, h->d_hist2[h->cur], hl2);
            fir_dec_kernel<float2, float2><<<g2, 128, sm, s>>>(reinterpret_cast<const float2*>(y1), n1, n1, h->d_hist2[h->cur],
                                                               h->d_taps2, h->T2, h->D2, reinterpret_cast<float2*>(out_dev),
                                                               out_stride, n2);
            fir_hist_kernel<float2><<<K, 128, 0, s>>>(reinterpret_cast<const float2*>(y1), n1, n1, h->d_hist2[h->cur],
                                                     h->d_hist2[h->cur ^ 1], hl2);
        }
    }
    WC_CUDA(cudaGetLastError());
    h->cur ^= 1;
    // advance the NCO sample index (system.py:1459-1466), wrap at one second of samples
    for (int k = 0; k < K; ++k) {
        h->need1[k] = 0;
        h->need2[k] = 0;
        DdcChan& c = h->chan[k];
        if (c.shift) {
            c.sample_idx += n_samples;
            if (c.sample_idx >= h->fs) c.sample_idx %= h->fs;
        }
    }
    return 0;
}

int wc_ddc_process_host(wc_ddc* h, const void* iq_host, int n_samples, void* out_host) {
    WC_REQUIRE(h && iq_host && out_host, "wc_ddc_process_host: null argument");
    if (n_samples <= 0) return 0;
    const int n2 = wc_ddc_out_len(h, n_samples);
    const size_t esz = h->keep_f64 ? sizeof(double2) : sizeof(float2);
    if (grow(&h->d_in, &h->in_bytes, sizeof(float2) * (size_t)n_samples)) return -2;
    if (grow(&h->d_out, &h->out_bytes, esz * (size_t)h->K * n2)) return -2;
    WC_CUDA(cudaMemcpyAsync(h->d_in, iq_host, sizeof(float2) * (size_t)n_samples, cudaMemcpyHostToDevice, h->stream));
    int rc = wc_ddc_process(h, h->d_in, n_samples, h->d_out, n2, h->stream);
    if (rc) return rc;
    WC_CUDA(cudaMemcpyAsync(out_host, h->d_out, esz * (size_t)h->K * n2, cudaMemcpyDeviceToHost, h->stream));
    WC_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

}  // extern "C"
