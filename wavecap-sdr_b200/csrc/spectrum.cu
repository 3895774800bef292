// Spectrum / waterfall producer (config C3): FFTBackend.execute of wavecapsdr/dsp/fft/
// (base.py:31-77, scipy_backend.py:38-79):
//     w = float32(np.hanning(N));  X = fft(iq[:N] * w);  P = 20*log10(|fftshift(X)| + 1e-10) -> float32
// plus the frontend's K-frame mean of dB frames (SpectrumAnalyzer.react.tsx:309-327) fused in.
//
// N = 65536 (tuned, four-step 256 x 256):
//   pass A  one CTA = 16 columns n2 of one frame: coalesced load (128 B runs) * Hann window into
//           shared memory, 256-point FFT down each column (16 threads, two in-register radix-16
//           passes), twiddle W_N^(n2*k1) from two 256-entry tables, coalesced store to a scratch
//           that is sized to stay L2-resident.
//   pass B  one CTA = 16 rows k1: 256-point FFT along each row, |X|, 20*log10, accumulate the K
//           frames of an averaging group in registers, fftshift folded into the store index.
//   HBM traffic per input sample: 8 B in + 4/K B out (the 16 B of scratch traffic stays in L2).
// Any other power of two: windowed radix-2 Stockham passes in global memory (correct, untuned).
#include <math.h>
#include <vector>

#include "../../include/wcsdr_b200.h"
#include "common.cuh"
#include "fft16.cuh"

namespace wc {

constexpr int SP_N1 = 256, SP_N2 = 256, SP_N = SP_N1 * SP_N2;
constexpr int SP_COLS = 16;        // columns (pass A) / rows (pass B) per CTA
constexpr int SP_STRIDE = 273;     // complex words per shared row: >= 272 and == 1 (mod 16) -> bank skew
constexpr int SP_THREADS = 256;

struct SpSmem {
    u64 tile[SP_COLS * SP_STRIDE];  // 34.9 KB
    float2 tw[256];                 // tw[k1*16 + t] = exp(-2 pi i k1 t / 256)
    float2 wa[256];                 // exp(-2 pi i m / 65536), m = 0..255
    float2 wb[256];                 // exp(-2 pi i m / 256),   m = 0..255
};

__device__ __forceinline__ void sp_tables(SpSmem& sm, int tid) {
    float s, c;
    sincospif(-(float)((tid >> 4) * (tid & 15)) * (1.0f / 128.0f), &s, &c);
    sm.tw[tid] = make_float2(c, s);
    sincospif(-(float)tid * (1.0f / 32768.0f), &s, &c);
    sm.wa[tid] = make_float2(c, s);
    sincospif(-(float)tid * (1.0f / 128.0f), &s, &c);
    sm.wb[tid] = make_float2(c, s);
}

// 256-point FFT of the shared row `reg` by the 16 threads of one half-warp (t = lane & 15).
// On return thread t holds X[t + 16*k2] in v[rev4(k2)].
__device__ __forceinline__ void fft256_row(u64* reg, const float2* tw, int t, u64 (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = reg[t + 16 * i];
    __syncwarp();
    fft16(v);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        u64 w = v[rev4(k1)];
        if (k1 > 0) {
            const float2 q = tw[k1 * 16 + t];
            w = twid(w, q.x, -q.y);
        }
        reg[t * 17 + k1] = w;
    }
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = reg[n2 * 17 + t];
    __syncwarp();
    fft16(v);
}

// pass A: grid (N2/16, n_frames)
__global__ void __launch_bounds__(SP_THREADS, 2) spectrum_pass_a(const float2* __restrict__ iq, long long frame_stride,
                                                                 const float* __restrict__ window,
                                                                 u64* __restrict__ scratch) {
    __shared__ SpSmem sm;
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * SP_COLS;
    const long long frame = blockIdx.y;
    sp_tables(sm, tid);
    const float2* x = iq + frame * frame_stride;
#pragma unroll 4
    for (int j = 0; j < 16; ++j) {
        const int id = tid + SP_THREADS * j;
        const int n1 = id >> 4, col = id & 15;
        const int n = n1 * SP_N2 + c0 + col;
        const float2 s = x[n];
        const float w = window[n];
        sm.tile[col * SP_STRIDE + n1] = pk2(s.x * w, s.y * w);
    }
    __syncthreads();
    {
        const int g = tid >> 4, t = tid & 15;
        u64* reg = sm.tile + g * SP_STRIDE;
        u64 v[16];
        fft256_row(reg, sm.tw, t, v);
        __syncwarp();
        const int n2 = c0 + g;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int k1 = t + 16 * q;
            const int m = n2 * k1;  // < 65536
            const float2 a = sm.wa[m & 255], b = sm.wb[m >> 8];
            const float2 wv = cmul(a, b);  // exp(-2 pi i m / 65536)
            reg[k1] = twid(v[rev4(q)], wv.x, -wv.y);
        }
    }
    __syncthreads();
    u64* T = scratch + frame * SP_N;
#pragma unroll 4
    for (int j = 0; j < 16; ++j) {
        const int id = tid + SP_THREADS * j;
        const int k1 = id >> 4, col = id & 15;
        T[k1 * SP_N2 + c0 + col] = sm.tile[col * SP_STRIDE + k1];
    }
}

// pass B: grid (N1/16, n_groups); each CTA accumulates `avg` frames.
__global__ void __launch_bounds__(SP_THREADS, 2) spectrum_pass_b(const u64* __restrict__ scratch, int avg, int n_frames,
                                                                 float* __restrict__ out) {
    __shared__ SpSmem sm;
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * SP_COLS;
    const int grp = blockIdx.y;
    sp_tables(sm, tid);
    const int g = tid >> 4, t = tid & 15;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    const int f0 = grp * avg;
    const int cnt = min(avg, n_frames - f0);
    for (int f = 0; f < cnt; ++f) {
        const u64* T = scratch + (long long)(f0 + f) * SP_N;
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < 16; ++j) sm.tile[j * SP_STRIDE + tid] = T[(r0 + j) * SP_N2 + tid];
        __syncthreads();
        u64 v[16];
        fft256_row(sm.tile + g * SP_STRIDE, sm.tw, t, v);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float re = lo2(v[rev4(q)]), im = hi2(v[rev4(q)]);
            const float mag = sqrtf(fmaf(re, re, im * im));
            acc[q] += 6.02059991327962f * __log2f(mag + 1e-10f);  // 20*log10(x) = 20*log10(2)*log2(x)
        }
    }
    // stage as [k2][16 rows] so that the 16 adjacent k1 of one k2 leave as one 64-byte run
    __syncthreads();
    float* so = reinterpret_cast<float*>(sm.tile);
    const float inv = 1.0f / (float)cnt;
#pragma unroll
    for (int q = 0; q < 16; ++q) so[(t + 16 * q) * 17 + g] = acc[q] * inv;
    __syncthreads();
    float* o = out + (long long)grp * SP_N;
#pragma unroll 4
    for (int j = 0; j < 16; ++j) {
        const int id = tid + SP_THREADS * j;
        const int k2 = id >> 4, row = id & 15;
        const int k = (r0 + row) + SP_N1 * k2;
        o[k ^ (SP_N / 2)] = so[k2 * 17 + row];  // fftshift
    }
}

// ---- generic power-of-two path: Stockham autosort radix-2 in global memory --------------------
__global__ void sp_window_kernel(const float2* __restrict__ iq, long long frame_stride, const float* __restrict__ window,
                                 float2* __restrict__ dst, int n) {
    const long long frame = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float2 s = iq[frame * frame_stride + i];
        const float w = window[i];
        dst[frame * n + i] = make_float2(s.x * w, s.y * w);
    }
}

// one Stockham stage: Ns = current sub-transform size (1, 2, 4, ...)
__global__ void sp_stockham_kernel(const float2* __restrict__ src, float2* __restrict__ dst, int n, int ns) {
    const long long frame = blockIdx.y;
    const float2* s = src + frame * n;
    float2* d = dst + frame * n;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n / 2; j += gridDim.x * blockDim.x) {
        const int k = j & (ns - 1);
        float sn, cs;
        sincospif(-(float)k / (float)ns, &sn, &cs);
        const float2 a = s[j];
        const float2 b = cmul(s[j + n / 2], make_float2(cs, sn));
        const int j0 = ((j - k) << 1) + k;
        d[j0] = make_float2(a.x + b.x, a.y + b.y);
        d[j0 + ns] = make_float2(a.x - b.x, a.y - b.y);
    }
}

__global__ void sp_db_kernel(const float2* __restrict__ X, int n, int avg, int n_frames, float* __restrict__ out) {
    const int grp = blockIdx.y;
    const int f0 = grp * avg;
    const int cnt = min(avg, n_frames - f0);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int f = 0; f < cnt; ++f) {
            const float2 v = X[(long long)(f0 + f) * n + k];
            acc += 6.02059991327962f * __log2f(sqrtf(fmaf(v.x, v.x, v.y * v.y)) + 1e-10f);
        }
        out[(long long)grp * n + (k ^ (n / 2))] = acc / (float)cnt;
    }
}

}  // namespace wc

using namespace wc;

struct wc_spectrum {
    int n = 0;
    float* d_window = nullptr;
    std::vector<float> h_window;
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    void* d_in = nullptr;
    size_t in_bytes = 0;
    void* d_out = nullptr;
    size_t out_bytes = 0;
    cudaStream_t stream = nullptr;
};

static int sp_ensure(void** p, size_t* cap, size_t need) {
    if (*cap >= need) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    WC_CUDA(cudaMalloc(p, need));
    *cap = need;
    return 0;
}

extern "C" {

int wc_spectrum_create(int fft_size, wc_spectrum** out) {
    WC_REQUIRE(out != nullptr, "wc_spectrum_create: out is null");
    WC_REQUIRE(fft_size >= 2 && fft_size <= (1 << 24) && (fft_size & (fft_size - 1)) == 0,
               "wc_spectrum_create: fft_size %d must be a power of two in [2, 2^24] (no CPU fallback exists)", fft_size);
    wc_spectrum* h = new wc_spectrum();
    h->n = fft_size;
    // np.hanning(M): 0.5 + 0.5*cos(pi*n/(M-1)), n = 1-M, 3-M, ..., M-1  -> float32 (fft/base.py:54-59)
    h->h_window.resize(fft_size);
    for (int i = 0; i < fft_size; ++i) {
        const double nn = (double)(1 - fft_size + 2 * i);
        h->h_window[i] = (float)(0.5 + 0.5 * cos(M_PI * nn / (double)(fft_size - 1)));
    }
    if (cudaMalloc(&h->d_window, sizeof(float) * fft_size) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("wc_spectrum_create: CUDA allocation failed");
        delete h;
        return -2;
    }
    cudaMemcpy(h->d_window, h->h_window.data(), sizeof(float) * fft_size, cudaMemcpyHostToDevice);
    *out = h;
    return 0;
}

void wc_spectrum_destroy(wc_spectrum* h) {
    if (!h) return;
    cudaFree(h->d_window);
    if (h->d_scratch) cudaFree(h->d_scratch);
    if (h->d_in) cudaFree(h->d_in);
    if (h->d_out) cudaFree(h->d_out);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int wc_spectrum_window(const wc_spectrum* h, float* window_host) {
    WC_REQUIRE(h && window_host, "wc_spectrum_window: null argument");
    for (int i = 0; i < h->n; ++i) window_host[i] = h->h_window[i];
    return 0;
}

// n_frames frames of fft_size samples, frame f at iq + f*frame_stride; consecutive groups of `avg`
// frames are averaged in dB. out: float32 [ceil(n_frames/avg)][fft_size], fftshifted.
int wc_spectrum_execute(wc_spectrum* h, const void* iq_dev, long long frame_stride, int n_frames, int avg,
                        float* power_db_dev, void* stream_v) {
    WC_REQUIRE(h && iq_dev && power_db_dev, "wc_spectrum_execute: null argument");
    WC_REQUIRE(n_frames >= 1 && avg >= 1, "wc_spectrum_execute: n_frames and avg must be >= 1");
    cudaStream_t st = (cudaStream_t)stream_v;
    const int n = h->n;
    const float2* iq = reinterpret_cast<const float2*>(iq_dev);
    // process in slabs whose scratch stays L2-resident (<= 64 MB)
    int slab = (int)((64ull << 20) / (sizeof(float2) * (size_t)n));
    if (slab < avg) slab = avg;
    slab -= slab % avg;
    if (slab > n_frames) slab = ((n_frames + avg - 1) / avg) * avg;
    const size_t need = sizeof(float2) * (size_t)n * slab * (n == SP_N ? 1 : 2);
    if (sp_ensure(&h->d_scratch, &h->scratch_bytes, need)) return -2;
    for (int f0 = 0; f0 < n_frames; f0 += slab) {
        const int cnt = (n_frames - f0 < slab) ? n_frames - f0 : slab;
        const int groups = (cnt + avg - 1) / avg;
        float* o = power_db_dev + (long long)(f0 / avg) * n;
        const float2* x = iq + (long long)f0 * frame_stride;
        if (n == SP_N) {
            u64* T = reinterpret_cast<u64*>(h->d_scratch);
            spectrum_pass_a<<<dim3(SP_N2 / SP_COLS, cnt), SP_THREADS, 0, st>>>(x, frame_stride, h->d_window, T);
            spectrum_pass_b<<<dim3(SP_N1 / SP_COLS, groups), SP_THREADS, 0, st>>>(T, avg, cnt, o);
        } else {
            float2* A = reinterpret_cast<float2*>(h->d_scratch);
            float2* B = A + (size_t)n * slab;
            int bx = (n / 2 + 255) / 256;
            if (bx > 1024) bx = 1024;
            if (bx < 1) bx = 1;
            sp_window_kernel<<<dim3(bx, cnt), 256, 0, st>>>(x, frame_stride, h->d_window, A, n);
            for (int ns = 1; ns < n; ns <<= 1) {
                sp_stockham_kernel<<<dim3(bx, cnt), 256, 0, st>>>(A, B, n, ns);
                float2* t = A;
                A = B;
                B = t;
            }
            sp_db_kernel<<<dim3(bx, groups), 256, 0, st>>>(A, n, avg, cnt, o);
        }
        WC_CUDA(cudaGetLastError());
    }
    return 0;
}

int wc_spectrum_execute_host(wc_spectrum* h, const void* iq_host, long long frame_stride, int n_frames, int avg,
                             float* power_db_host) {
    WC_REQUIRE(h && iq_host && power_db_host, "wc_spectrum_execute_host: null argument");
    WC_REQUIRE(n_frames >= 1 && avg >= 1, "wc_spectrum_execute_host: n_frames and avg must be >= 1");
    const int n = h->n;
    const size_t in_need = sizeof(float2) * (size_t)n * n_frames;
    const int groups = (n_frames + avg - 1) / avg;
    const size_t out_need = sizeof(float) * (size_t)n * groups;
    if (sp_ensure(&h->d_in, &h->in_bytes, in_need)) return -2;
    if (sp_ensure(&h->d_out, &h->out_bytes, out_need)) return -2;
    // only the first fft_size samples of each frame are ever read: copy just those
    WC_CUDA(cudaMemcpy2DAsync(h->d_in, sizeof(float2) * n, iq_host, sizeof(float2) * frame_stride, sizeof(float2) * n,
                              n_frames, cudaMemcpyHostToDevice, h->stream));
    int rc = wc_spectrum_execute(h, h->d_in, n, n_frames, avg, reinterpret_cast<float*>(h->d_out), h->stream);
    if (rc) return rc;
    WC_CUDA(cudaMemcpyAsync(power_db_host, h->d_out, out_need, cudaMemcpyDeviceToHost, h->stream));
    WC_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

}  // extern "C"
