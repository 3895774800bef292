// Optional audio clean-up stages of the FM chains (off by default in the reference):
//   wc_noise_blanker   wavecapsdr/dsp/filters.py:267-343  median-referenced impulse blanking with dilation
//   wc_spectral_nr     wavecapsdr/dsp/filters.py:346-459  STFT (1024, hop 512, periodic Hann) Wiener-style gain against
//                      the per-bin 10th-percentile noise floor, overlap-add resynthesis
// Both work on batches of float32 sequences (seq_stride apart), one launch sequence per call.
#include <math.h>
#include <vector>

#include "../../include/wcsdr_b200.h"
#include "common.cuh"

namespace wc {

// ---------------------------------------------------------------------------------------------
// noise blanker
// ---------------------------------------------------------------------------------------------
// exact order statistics of |x| by 4-pass radix select on the IEEE bit pattern (non-negative floats
// order like their unsigned bit patterns); np.median of an even-length array is the float32 mean of
// the two middle values.
__global__ void __launch_bounds__(1024) nb_median_kernel(const float* __restrict__ x, int n, long long stride,
                                                         float* __restrict__ med) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_rank;
    const float* xs = x + (long long)blockIdx.x * stride;
    float vals[2];
    const int ranks[2] = {(n - 1) / 2, n / 2};
    for (int q = 0; q < 2; ++q) {
        if (q == 1 && ranks[1] == ranks[0]) {
            vals[1] = vals[0];
            break;
        }
        unsigned prefix = 0u, mask = 0u;
        unsigned rank = (unsigned)ranks[q];
        for (int pass = 3; pass >= 0; --pass) {
            for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const unsigned u = __float_as_uint(fabsf(xs[i]));
                if ((u & mask) == prefix) atomicAdd(&hist[(u >> (8 * pass)) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned cum = 0u;
                int b = 0;
                for (; b < 256; ++b) {
                    if (cum + hist[b] > rank) break;
                    cum += hist[b];
                }
                s_prefix = prefix | ((unsigned)b << (8 * pass));
                s_rank = rank - cum;
            }
            __syncthreads();
            prefix = s_prefix;
            rank = s_rank;
            mask |= 0xffu << (8 * pass);
            __syncthreads();
        }
        vals[q] = __uint_as_float(prefix);
    }
    if (threadIdx.x == 0) med[blockIdx.x] = __fmul_rn(__fadd_rn(vals[0], vals[1]), 0.5f);
}

__global__ void nb_apply_kernel(const float* __restrict__ x, float* __restrict__ y, int n, long long stride,
                                const float* __restrict__ med, float thr_lin, int width) {
    const int seq = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* xs = x + (long long)seq * stride;
    const float m = med[seq];
    const float v = xs[i];
    float out = v;
    if (!(m < 1e-10f)) {
        const float thr = __fmul_rn(m, thr_lin);
        const int lo = max(0, i - width), hi = min(n - 1, i + width);
        bool hit = false;
        for (int j = lo; j <= hi; ++j) hit |= fabsf(xs[j]) > thr;
        if (hit) out = 0.0f;
    }
    y[(long long)seq * stride + i] = out;
}

// ---------------------------------------------------------------------------------------------
// spectral noise reduction
// ---------------------------------------------------------------------------------------------
constexpr int NR_N = 1024, NR_HOP = 512, NR_BINS = NR_N / 2 + 1;

// in-place radix-2 DIT FFT of 1024 complex points in shared memory by 256 threads (input already bit-reversed)
__device__ void nr_fft1024(float2* buf, const float2* tw) {
    for (int len = 2; len <= NR_N; len <<= 1) {
        const int half = len >> 1, tstep = NR_N / len;
        for (int b = threadIdx.x; b < NR_N / 2; b += blockDim.x) {
            const int grp = b / half, k = b - grp * half;
            const int i0 = grp * len + k, i1 = i0 + half;
            const float2 w = tw[k * tstep];
            const float2 a = buf[i0], c = cmul(buf[i1], w);
            buf[i0] = cadd(a, c);
            buf[i1] = csub(a, c);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ int nr_bitrev10(int i) { return (int)(__brev((unsigned)i) >> 22); }

// analysis: frame f of sequence s -> spectrum bins 0..512 (complex) and magnitudes
__global__ void __launch_bounds__(256) nr_analysis_kernel(const float* __restrict__ x, long long stride, const float* __restrict__ win,
                                                          int n_frames, float2* __restrict__ X, float* __restrict__ mag) {
    __shared__ float2 buf[NR_N];
    __shared__ float2 tw[NR_N / 2];
    const int f = blockIdx.x, s = blockIdx.y;
    const float* xs = x + (long long)s * stride + (long long)f * NR_HOP;
    for (int i = threadIdx.x; i < NR_N / 2; i += blockDim.x) {
        float sn, cs;
        sincospif(-(float)i * (2.0f / NR_N), &sn, &cs);
        tw[i] = make_float2(cs, sn);
    }
    for (int i = threadIdx.x; i < NR_N; i += blockDim.x) buf[nr_bitrev10(i)] = make_float2(__fmul_rn(xs[i], win[i]), 0.f);
    __syncthreads();
    nr_fft1024(buf, tw);
    const long long o = ((long long)s * n_frames + f) * NR_BINS;
    for (int b = threadIdx.x; b < NR_BINS; b += blockDim.x) {
        const float2 v = buf[b];
        X[o + b] = v;
        mag[o + b] = (float)sqrt((double)v.x * v.x + (double)v.y * v.y);
    }
}

// noise floor: 10th percentile (numpy 'linear') of each bin's magnitudes across frames, by rank counting
__global__ void __launch_bounds__(256) nr_floor_kernel(const float* __restrict__ mag, int n_frames, float* __restrict__ floor_out) {
    extern __shared__ float v[];
    const int b = blockIdx.x, s = blockIdx.y;
    const float* m = mag + (long long)s * n_frames * NR_BINS + b;
    for (int i = threadIdx.x; i < n_frames; i += blockDim.x) v[i] = m[(long long)i * NR_BINS];
    __syncthreads();
    const double pos = 0.1 * (double)(n_frames - 1);
    const int lo = (int)floor(pos);
    const int hi = min(lo + 1, n_frames - 1);
    __shared__ float s_lo, s_hi;
    for (int i = threadIdx.x; i < n_frames; i += blockDim.x) {
        const float vi = v[i];
        int rank = 0;
        for (int j = 0; j < n_frames; ++j) rank += (v[j] < vi) || (v[j] == vi && j < i);
        if (rank == lo) s_lo = vi;
        if (rank == hi) s_hi = vi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double t = pos - (double)lo;
        const double a = (double)s_lo, c = (double)s_hi;
        floor_out[(long long)s * NR_BINS + b] = (float)(a + (c - a) * t);
    }
}

// synthesis: gain, Hermitian extension, inverse FFT, window -> frames[s][f][1024]
__global__ void __launch_bounds__(256) nr_synthesis_kernel(const float2* __restrict__ X, const float* __restrict__ mag,
                                                           const float* __restrict__ nfloor, const float* __restrict__ win,
                                                           int n_frames, float red_lin, float* __restrict__ frames) {
    __shared__ float2 buf[NR_N];
    __shared__ float2 tw[NR_N / 2];
    const int f = blockIdx.x, s = blockIdx.y;
    for (int i = threadIdx.x; i < NR_N / 2; i += blockDim.x) {
        float sn, cs;
        sincospif(-(float)i * (2.0f / NR_N), &sn, &cs);
        tw[i] = make_float2(cs, sn);
    }
    const long long o = ((long long)s * n_frames + f) * NR_BINS;
    for (int b = threadIdx.x; b < NR_BINS; b += blockDim.x) {
        const float m = mag[o + b];
        const float ns = __fmul_rn(nfloor[(long long)s * NR_BINS + b], red_lin);
        const float r = __fdiv_rn(ns, fmaxf(m, 1e-10f));
        float g = fmaxf(0.0f, __fsub_rn(1.0f, __fmul_rn(r, r)));
        g = fmaxf(g, 0.1f);
        float2 v = X[o + b];
        v.x *= g;
        v.y *= g;
        if (b == 0 || b == NR_N / 2) v.y = 0.f;            // irfft ignores the imaginary part of DC and Nyquist
        // inverse via conj(fft(conj(X))): load conj(X) bit-reversed, with the Hermitian mirror
        buf[nr_bitrev10(b)] = make_float2(v.x, -v.y);
        if (b > 0 && b < NR_N / 2) buf[nr_bitrev10(NR_N - b)] = make_float2(v.x, v.y);   // conj(conj(X[b])) = X[b]
    }
    __syncthreads();
    nr_fft1024(buf, tw);
    float* fr = frames + ((long long)s * n_frames + f) * NR_N;
    for (int i = threadIdx.x; i < NR_N; i += blockDim.x) fr[i] = __fmul_rn(buf[i].x * (1.0f / NR_N), win[i]);
}

__global__ void nr_overlap_add_kernel(const float* __restrict__ frames, const float* __restrict__ win, int n_frames, int out_len,
                                      float* __restrict__ y, long long y_stride) {
    const int s = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= out_len) return;
    const float* fr = frames + (long long)s * n_frames * NR_N;
    float acc = 0.f, ws = 0.f;
    const int f_hi = min(j / NR_HOP, n_frames - 1);
    const int f_lo = (j >= NR_N) ? (j - NR_N) / NR_HOP + 1 : 0;
    for (int f = f_lo; f <= f_hi; ++f) {               // ascending frame order, like the reference's loop
        const int i = j - f * NR_HOP;
        if (i < 0 || i >= NR_N) continue;
        acc = __fadd_rn(acc, fr[(long long)f * NR_N + i]);
        ws = __fadd_rn(ws, __fmul_rn(win[i], win[i]));
    }
    y[(long long)s * y_stride + j] = __fdiv_rn(acc, fmaxf(ws, 1e-10f));
}

}  // namespace wc

using namespace wc;

extern "C" {

int wc_noise_blanker(const float* x_dev, float* y_dev, int n, long long seq_stride, int n_seq, float threshold_db,
                     int blanking_width, void* stream_v) {
    WC_REQUIRE(x_dev && y_dev, "wc_noise_blanker: null argument");
    WC_REQUIRE(n >= 0 && n_seq >= 1 && blanking_width >= 0 && seq_stride >= n, "wc_noise_blanker: bad sizes");
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream_v;
    float* d_med = nullptr;
    WC_CUDA(cudaMallocAsync((void**)&d_med, sizeof(float) * n_seq, s));
    nb_median_kernel<<<n_seq, 1024, 0, s>>>(x_dev, n, seq_stride, d_med);
    const float thr_lin = (float)pow(10.0, (double)threshold_db / 20.0);
    nb_apply_kernel<<<dim3((n + 255) / 256, n_seq), 256, 0, s>>>(x_dev, y_dev, n, seq_stride, d_med, thr_lin, blanking_width);
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaFreeAsync(d_med, s));
    return 0;
}

/* output length of spectral_noise_reduction for an n-sample input: the reference returns the (n_frames-1)*hop + fft_size
 * samples its frames cover (<= n), or the input itself when n < fft_size */
int wc_spectral_nr_out_len(int n) {
    if (n < NR_N) return n;
    return ((n - NR_N) / NR_HOP) * NR_HOP + NR_N;
}

int wc_spectral_nr(const float* x_dev, int n, long long seq_stride, int n_seq, float reduction_db, float* y_dev,
                   long long y_stride, void* stream_v) {
    WC_REQUIRE(x_dev && y_dev, "wc_spectral_nr: null argument");
    WC_REQUIRE(n >= 0 && n_seq >= 1 && seq_stride >= n, "wc_spectral_nr: bad sizes");
    cudaStream_t s = (cudaStream_t)stream_v;
    if (n == 0) return 0;
    if (n < NR_N) {
        WC_CUDA(cudaMemcpy2DAsync(y_dev, sizeof(float) * y_stride, x_dev, sizeof(float) * seq_stride, sizeof(float) * n, n_seq,
                                  cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    const int n_frames = (n - NR_N) / NR_HOP + 1;
    WC_REQUIRE(n_frames <= 8192, "wc_spectral_nr: %d frames exceed the 8192-frame percentile buffer", n_frames);
    const int out_len = wc_spectral_nr_out_len(n);
    WC_REQUIRE(y_stride >= out_len, "wc_spectral_nr: y_stride %lld < %d", y_stride, out_len);
    // scipy.signal.windows.hann(1024, sym=False) -> float32
    std::vector<float> hw(NR_N);
    for (int i = 0; i < NR_N; ++i) hw[i] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)i / (double)NR_N));
    const size_t per = (size_t)n_seq * n_frames;
    float *d_win = nullptr, *d_mag = nullptr, *d_floor = nullptr, *d_frames = nullptr;
    float2* d_X = nullptr;
    WC_CUDA(cudaMallocAsync((void**)&d_win, sizeof(float) * NR_N, s));
    WC_CUDA(cudaMallocAsync((void**)&d_mag, sizeof(float) * per * NR_BINS, s));
    WC_CUDA(cudaMallocAsync((void**)&d_X, sizeof(float2) * per * NR_BINS, s));
    WC_CUDA(cudaMallocAsync((void**)&d_floor, sizeof(float) * (size_t)n_seq * NR_BINS, s));
    WC_CUDA(cudaMallocAsync((void**)&d_frames, sizeof(float) * per * NR_N, s));
    WC_CUDA(cudaMemcpyAsync(d_win, hw.data(), sizeof(float) * NR_N, cudaMemcpyHostToDevice, s));
    WC_CUDA(cudaStreamSynchronize(s));  // hw is a stack-lifetime host buffer
    nr_analysis_kernel<<<dim3(n_frames, n_seq), 256, 0, s>>>(x_dev, seq_stride, d_win, n_frames, d_X, d_mag);
    nr_floor_kernel<<<dim3(NR_BINS, n_seq), 256, sizeof(float) * n_frames, s>>>(d_mag, n_frames, d_floor);
    const float red_lin = (float)pow(10.0, (double)reduction_db / 20.0);
    nr_synthesis_kernel<<<dim3(n_frames, n_seq), 256, 0, s>>>(d_X, d_mag, d_floor, d_win, n_frames, red_lin, d_frames);
    nr_overlap_add_kernel<<<dim3((out_len + 255) / 256, n_seq), 256, 0, s>>>(d_frames, d_win, n_frames, out_len, y_dev, y_stride);
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaFreeAsync(d_win, s));
    WC_CUDA(cudaFreeAsync(d_mag, s));
    WC_CUDA(cudaFreeAsync(d_X, s));
    WC_CUDA(cudaFreeAsync(d_floor, s));
    WC_CUDA(cudaFreeAsync(d_frames, s));
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// wire packers + audio level metering: the output stage of the analog chain
//   capture.pack_iq16 / pack_pcm16 / pack_f32 (capture.py:102-144): clip to [-1, 1], scale by 32767, truncate to
//   int16 (numpy astype truncates toward zero); Channel._update_audio_metrics (capture.py:633-661): RMS, peak and
//   the count of |x| > 0.95 per sequence — fused so that audio leaves the GPU already as PCM16 with its levels.
// ---------------------------------------------------------------------------------------------
namespace wc {

__global__ void pack16_kernel(const float* __restrict__ x, short* __restrict__ y, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float v = fminf(fmaxf(x[i], -1.0f), 1.0f);
    y[i] = (short)(int)__fmul_rn(v, 32767.0f);   // float -> int conversion truncates toward zero like astype(int16)
}

__global__ void clip_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    y[i] = fminf(fmaxf(x[i], -1.0f), 1.0f);
}

// per sequence: sum x^2 (float64), max |x|, count(|x| > 0.95)
__global__ void __launch_bounds__(256) audio_levels_kernel(const float* __restrict__ x, int n, long long stride,
                                                           double* __restrict__ sumsq, float* __restrict__ peak,
                                                           int* __restrict__ clip_count) {
    __shared__ double rs[8];
    __shared__ float rp[8];
    __shared__ int rc[8];
    const float* xs = x + (long long)blockIdx.x * stride;
    double s = 0.0;
    float p = 0.f;
    int c = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = xs[i], a = fabsf(v);
        s += (double)v * (double)v;
        p = fmaxf(p, a);
        c += a > 0.95f;
    }
    s = warp_sum(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        p = fmaxf(p, __shfl_xor_sync(0xffffffffu, p, o));
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) {
        rs[threadIdx.x >> 5] = s;
        rp[threadIdx.x >> 5] = p;
        rc[threadIdx.x >> 5] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0;
        float tp = 0.f;
        int tc = 0;
        for (int w = 0; w < 8; ++w) {
            ts += rs[w];
            tp = fmaxf(tp, rp[w]);
            tc += rc[w];
        }
        sumsq[blockIdx.x] = ts;
        peak[blockIdx.x] = tp;
        clip_count[blockIdx.x] = tc;
    }
}

}  // namespace wc

extern "C" {

/* fmt 0: int16 PCM / interleaved IQ16 (pack_pcm16, pack_iq16 — complex64 input is just 2*total floats);
 * fmt 1: clipped float32 (pack_f32) */
int wc_pack(const float* x_dev, void* y_dev, long long total_floats, int fmt, void* stream_v) {
    WC_REQUIRE(x_dev && y_dev, "wc_pack: null argument");
    WC_REQUIRE(fmt == 0 || fmt == 1, "wc_pack: fmt must be 0 (int16) or 1 (float32)");
    if (total_floats <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream_v;
    const unsigned blocks = (unsigned)((total_floats + 255) / 256);
    if (fmt == 0) wc::pack16_kernel<<<blocks, 256, 0, s>>>(x_dev, reinterpret_cast<short*>(y_dev), total_floats);
    else wc::clip_f32_kernel<<<blocks, 256, 0, s>>>(x_dev, reinterpret_cast<float*>(y_dev), total_floats);
    WC_CUDA(cudaGetLastError());
    return 0;
}

int wc_audio_levels(const float* x_dev, int n, long long seq_stride, int n_seq, double* sumsq_dev, float* peak_dev,
                    int* clip_count_dev, void* stream_v) {
    WC_REQUIRE(x_dev && sumsq_dev && peak_dev && clip_count_dev, "wc_audio_levels: null argument");
    WC_REQUIRE(n >= 1 && n_seq >= 1, "wc_audio_levels: bad sizes");
    wc::audio_levels_kernel<<<n_seq, 256, 0, (cudaStream_t)stream_v>>>(x_dev, n, seq_stride, sumsq_dev, peak_dev, clip_count_dev);
    WC_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
