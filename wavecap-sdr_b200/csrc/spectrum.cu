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
#include <stdlib.h>
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
// fft256_core: same, with the row already in registers (v[i] = x[t + 16*i]); `reg` is only the exchange area.
__device__ __forceinline__ void fft256_core(u64* reg, const float2* tw, int t, u64 (&v)[16]) {
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
__device__ __forceinline__ void fft256_row(u64* reg, const float2* tw, int t, u64 (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = reg[t + 16 * i];
    fft256_core(reg, tw, t, v);
}

// ---- Ampere-style async copies (LDGSTS): global -> shared without staging in registers ----
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
// cp_async_commit / cp_async_wait<N> live in common.cuh

#ifdef WC_DEV   // superseded variants, kept for A/B timing in dev builds only (-DWC_DEV)
// pass A: grid (N2/16, n_frames)
__global__ void __launch_bounds__(SP_THREADS, 4) spectrum_pass_a(const float2* __restrict__ iq, long long frame_stride,
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
__global__ void __launch_bounds__(SP_THREADS, 4) spectrum_pass_b(const u64* __restrict__ scratch, int avg, int n_frames,
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
#endif  // WC_DEV

#ifdef WC_DEV
// ---- pipelined, persistent versions (the ones the 65536-point path launches) ------------------------
// Both passes are streaming kernels whose only problem is latency, so each CTA keeps TWO tiles in shared
// memory: while the 256-point FFTs of tile k run, the LDGSTS copies of tile k+1 are in flight.
constexpr int SPB_STRIDE = 274;   // pass-B row stride in complex words: even (16-byte copies), >= 272 exchange area

struct SpSmemA {
    u64 tile[2][SP_COLS * SP_STRIDE];
    float2 tw[256], wa[256], wb[256];
};
struct SpSmemB {
    u64 tile[2][SP_COLS * SPB_STRIDE];
    float2 tw[256];
};

// pass A: CTA b owns column tile (b & 15) for frames (b >> 4), (b >> 4) + G, ...; the Hann window of its 16
// columns lives in registers for the whole kernel.
__global__ void __launch_bounds__(SP_THREADS, 2) spectrum_pass_a2(const float2* __restrict__ iq, long long frame_stride,
                                                                  const float* __restrict__ window, u64* __restrict__ scratch,
                                                                  int n_frames) {
    extern __shared__ __align__(16) unsigned char sp_raw[];
    SpSmemA& sm = *reinterpret_cast<SpSmemA*>(sp_raw);
    const int tid = threadIdx.x;
    const int c0 = (blockIdx.x & 15) * SP_COLS;
    const int G = gridDim.x >> 4;
    {
        float sn, cs;
        sincospif(-(float)((tid >> 4) * (tid & 15)) * (1.0f / 128.0f), &sn, &cs);
        sm.tw[tid] = make_float2(cs, sn);
        sincospif(-(float)tid * (1.0f / 32768.0f), &sn, &cs);
        sm.wa[tid] = make_float2(cs, sn);
        sincospif(-(float)tid * (1.0f / 128.0f), &sn, &cs);
        sm.wb[tid] = make_float2(cs, sn);
    }
    const int g = tid >> 4, t = tid & 15;
    float wv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) wv[i] = __ldg(window + (t + 16 * i) * SP_N2 + c0 + g);

    auto issue = [&](int f, int buf) {
        const float2* x = iq + (long long)f * frame_stride;
#pragma unroll 4
        for (int j = 0; j < 16; ++j) {
            const int id = tid + SP_THREADS * j;
            const int n1 = id >> 4, col = id & 15;
            cp_async8(&sm.tile[buf][col * SP_STRIDE + n1], x + n1 * SP_N2 + c0 + col);
        }
        cp_async_commit();
    };
    int f = blockIdx.x >> 4;
    if (f < n_frames) issue(f, 0);
    int buf = 0;
    for (; f < n_frames; f += G, buf ^= 1) {
        const bool more = f + G < n_frames;
        if (more) issue(f + G, buf ^ 1);
        if (more) cp_async_wait<1>();
        else cp_async_wait<0>();
        __syncthreads();
        u64* reg = sm.tile[buf] + g * SP_STRIDE;
        u64 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = mul2(reg[t + 16 * i], bc2(wv[i]));
        fft256_core(reg, sm.tw, t, v);
        __syncwarp();
        const int n2 = c0 + g;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int k1 = t + 16 * q;
            const int m = n2 * k1;  // < 65536
            const float2 a = sm.wa[m & 255], b = sm.wb[m >> 8];
            const float2 wq = cmul(a, b);  // exp(-2 pi i m / 65536)
            reg[k1] = twid(v[rev4(q)], wq.x, -wq.y);
        }
        __syncthreads();
        u64* T = scratch + (long long)f * SP_N;
#pragma unroll 4
        for (int j = 0; j < 16; ++j) {
            const int id = tid + SP_THREADS * j;
            const int k1 = id >> 4, col = id & 15;
            T[k1 * SP_N2 + c0 + col] = sm.tile[buf][col * SP_STRIDE + k1];
        }
        __syncthreads();  // tile[buf] is the target of the copies issued at the top of the next iteration but one
    }
}

// pass B: work item = (16 rows, averaging group); a CTA walks items blockIdx.x, +gridDim.x, ... and inside an
// item the frames of the group, always with the next tile's copies in flight.
__global__ void __launch_bounds__(SP_THREADS, 3) spectrum_pass_b2(const u64* __restrict__ scratch, int avg, int n_frames,
                                                                  int n_groups, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char sp_raw[];
    SpSmemB& sm = *reinterpret_cast<SpSmemB*>(sp_raw);
    const int tid = threadIdx.x;
    {
        float sn, cs;
        sincospif(-(float)((tid >> 4) * (tid & 15)) * (1.0f / 128.0f), &sn, &cs);
        sm.tw[tid] = make_float2(cs, sn);
    }
    const int g = tid >> 4, t = tid & 15;
    const int n_items = n_groups * 16;
    auto frames_in = [&](int item) { return min(avg, n_frames - (item >> 4) * avg); };
    auto issue = [&](int item, int fr, int buf) {
        const u64* T = scratch + (long long)((item >> 4) * avg + fr) * SP_N + (long long)(item & 15) * SP_COLS * SP_N2;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int id = tid + SP_THREADS * j;   // 16-byte chunk: 128 per row
            const int row = id >> 7, cc = id & 127;
            cp_async16(&sm.tile[buf][row * SPB_STRIDE + 2 * cc], T + row * SP_N2 + 2 * cc);
        }
        cp_async_commit();
    };
    int item = blockIdx.x, fr = 0, buf = 0;
    if (item < n_items) issue(item, 0, 0);
    float acc[16];
    while (item < n_items) {
        const int cnt = frames_in(item);
        if (fr == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.f;
        }
        // next tile in this CTA's sequence
        int nitem = item, nfr = fr + 1;
        if (nfr >= cnt) {
            nitem = item + gridDim.x;
            nfr = 0;
        }
        const bool more = nitem < n_items;
        if (more) issue(nitem, nfr, buf ^ 1);
        if (more) cp_async_wait<1>();
        else cp_async_wait<0>();
        __syncthreads();
        u64 v[16];
        fft256_row(sm.tile[buf] + g * SPB_STRIDE, sm.tw, t, v);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float re = lo2(v[rev4(q)]), im = hi2(v[rev4(q)]);
            const float mag = sqrtf(fmaf(re, re, im * im));
            acc[q] += 6.02059991327962f * __log2f(mag + 1e-10f);  // 20*log10(x) = 20*log10(2)*log2(x)
        }
        __syncthreads();  // all exchange traffic in tile[buf] done
        if (fr + 1 >= cnt) {
            // stage as [k2][16 rows] so that the 16 adjacent k1 of one k2 leave as one 64-byte run
            float* so = reinterpret_cast<float*>(sm.tile[buf]);
            const float inv = 1.0f / (float)cnt;
#pragma unroll
            for (int q = 0; q < 16; ++q) so[(t + 16 * q) * 17 + g] = acc[q] * inv;
            __syncthreads();
            float* o = out + (long long)(item >> 4) * SP_N;
            const int r0 = (item & 15) * SP_COLS;
#pragma unroll 4
            for (int j = 0; j < 16; ++j) {
                const int id = tid + SP_THREADS * j;
                const int k2 = id >> 4, row = id & 15;
                const int k = (r0 + row) + SP_N1 * k2;
                o[k ^ (SP_N / 2)] = so[k2 * 17 + row];  // fftshift
            }
            __syncthreads();
        }
        item = nitem;
        fr = nfr;
        buf ^= 1;
    }
}
#endif  // WC_DEV

// ---- register-direct versions: no staging through shared memory -------------------------------------
// ncu on the staged kernels shows them bound by shared-memory wavefronts (stage in, FFT read, exchange, stage out),
// not by DRAM. Here every thread loads its 16 FFT inputs straight from global memory into registers with lanes
// running along the contiguous direction (128-byte segments), and shared memory carries only the one 16x16
// exchange in the middle of each 256-point FFT.
//
// pass A3: thread (t = tid >> 4, col = tid & 15) owns column c0+col, rows n1 = t + 16 i. The 16 threads of one
// column sit in 16 different half-warps, so the exchange is CTA-wide (__syncthreads).
struct SpSmemA3 {
    u64 ex[SP_COLS * SP_STRIDE];   // per column: 16 x 17 exchange area (stride 273: columns land on distinct banks)
    float2 tw[256];
};

#ifdef WC_DEV
__global__ void __launch_bounds__(SP_THREADS, 3) spectrum_pass_a3(const float2* __restrict__ iq, long long frame_stride,
                                                                  const float* __restrict__ window, u64* __restrict__ scratch,
                                                                  int n_frames) {
    __shared__ SpSmemA3 sm;
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * SP_COLS;
    {
        float sn, cs;
        sincospif(-(float)((tid >> 4) * (tid & 15)) * (1.0f / 128.0f), &sn, &cs);
        sm.tw[tid] = make_float2(cs, sn);
    }
    const int t = tid >> 4, col = tid & 15;
    const int n2 = c0 + col;
    // four-step twiddle W_65536^(n2 * k1), k1 = t + 16 kb: base * step^kb
    float2 base, step;
    {
        float sn, cs;
        sincospif(-(float)(n2 * t) * (1.0f / 32768.0f), &sn, &cs);
        base = make_float2(cs, sn);
        sincospif(-(float)(n2 * 16) * (1.0f / 32768.0f), &sn, &cs);
        step = make_float2(cs, sn);
    }
    u64* ex = sm.ex + col * SP_STRIDE;
    const u64* xw = reinterpret_cast<const u64*>(iq);
    float wv[16];   // this thread's 16 Hann coefficients: the column tile is fixed, so they are loaded once per CTA
#pragma unroll
    for (int i = 0; i < 16; ++i) wv[i] = __ldg(window + (t + 16 * i) * SP_N2 + n2);
    for (int f = blockIdx.y; f < n_frames; f += gridDim.y) {
        const u64* x = xw + (long long)f * frame_stride;
        u64 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __ldcs(x + (t + 16 * i) * SP_N2 + n2);   // streamed once: evict-first
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = mul2(v[i], bc2(wv[i]));
        fft16(v);
        __syncthreads();   // previous frame's exchange reads are done
#pragma unroll
        for (int ka = 0; ka < 16; ++ka) {
            u64 w = v[rev4(ka)];
            if (ka > 0) {
                const float2 q = sm.tw[ka * 16 + t];
                w = twid(w, q.x, -q.y);
            }
            ex[t * 17 + ka] = w;
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = ex[m * 17 + t];   // thread t now plays ka = t
        fft16(v);
        u64* T = scratch + (long long)f * SP_N;
        float2 w = base;
#pragma unroll
        for (int kb = 0; kb < 16; ++kb) {
            const int k1 = t + 16 * kb;
            T[k1 * SP_N2 + n2] = twid(v[rev4(kb)], w.x, -w.y);
            w = cmul(w, step);
        }
    }
}
#endif  // WC_DEV

// pass A5 = A3 with the next frame's 16 operands prefetched into registers while the current frame is transformed
// (A3 loads, waits, computes, stores: between the load bursts nothing is in flight). The Hann coefficients are re-read
// from L1 every frame instead of being held in 16 registers, which pays for the prefetch buffer.
template <int MINB>
__global__ void __launch_bounds__(SP_THREADS, MINB) spectrum_pass_a5(const float2* __restrict__ iq, long long frame_stride,
                                                                     const float* __restrict__ window, u64* __restrict__ scratch,
                                                                     int n_frames) {
    __shared__ SpSmemA3 sm;
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * SP_COLS;
    {
        float sn, cs;
        sincospif(-(float)((tid >> 4) * (tid & 15)) * (1.0f / 128.0f), &sn, &cs);
        sm.tw[tid] = make_float2(cs, sn);
    }
    const int t = tid >> 4, col = tid & 15;
    const int n2 = c0 + col;
    float2 base, step;
    {
        float sn, cs;
        sincospif(-(float)(n2 * t) * (1.0f / 32768.0f), &sn, &cs);
        base = make_float2(cs, sn);
        sincospif(-(float)(n2 * 16) * (1.0f / 32768.0f), &sn, &cs);
        step = make_float2(cs, sn);
    }
    u64* ex = sm.ex + col * SP_STRIDE;
    const u64* xw = reinterpret_cast<const u64*>(iq);
    const float* wp = window + t * SP_N2 + n2;
    u64 nx[16];
    int f = blockIdx.y;
    if (f < n_frames) {
        const u64* x = xw + (long long)f * frame_stride;
#pragma unroll
        for (int i = 0; i < 16; ++i) nx[i] = __ldcs(x + (t + 16 * i) * SP_N2 + n2);
    }
    for (; f < n_frames; f += gridDim.y) {
        u64 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = mul2(nx[i], bc2(__ldg(wp + 16 * i * SP_N2)));
        const int fn = f + gridDim.y;
        if (fn < n_frames) {
            const u64* x = xw + (long long)fn * frame_stride;
#pragma unroll
            for (int i = 0; i < 16; ++i) nx[i] = __ldcs(x + (t + 16 * i) * SP_N2 + n2);
        }
        fft16(v);
        __syncthreads();   // previous frame's exchange reads are done
#pragma unroll
        for (int ka = 0; ka < 16; ++ka) {
            u64 w = v[rev4(ka)];
            if (ka > 0) {
                const float2 q = sm.tw[ka * 16 + t];
                w = twid(w, q.x, -q.y);
            }
            ex[t * 17 + ka] = w;
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = ex[m * 17 + t];
        fft16(v);
        u64* T = scratch + (long long)f * SP_N;
        float2 w = base;
#pragma unroll
        for (int kb = 0; kb < 16; ++kb) {
            const int k1 = t + 16 * kb;
            T[k1 * SP_N2 + n2] = twid(v[rev4(kb)], w.x, -w.y);
            w = cmul(w, step);
        }
    }
}

// pass B3: thread (g = tid >> 4 row, t = tid & 15); the next frame's row is prefetched into registers while the
// current one is transformed; exchange is half-warp local.
struct SpSmemB3 {
    u64 ex[SP_COLS * SP_STRIDE];
    float2 tw[256];
};

#ifdef WC_DEV
__global__ void __launch_bounds__(SP_THREADS, 2) spectrum_pass_b3(const u64* __restrict__ scratch, int avg, int n_frames,
                                                                  float* __restrict__ out) {
    __shared__ SpSmemB3 sm;
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * SP_COLS;
    const int grp = blockIdx.y;
    {
        float sn, cs;
        sincospif(-(float)((tid >> 4) * (tid & 15)) * (1.0f / 128.0f), &sn, &cs);
        sm.tw[tid] = make_float2(cs, sn);
    }
    __syncthreads();
    const int g = tid >> 4, t = tid & 15;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    const int f0 = grp * avg;
    const int cnt = min(avg, n_frames - f0);
    u64* ex = sm.ex + g * SP_STRIDE;
    u64 nx[16];
    {
        const u64* T = scratch + (long long)f0 * SP_N + (long long)(r0 + g) * SP_N2;
#pragma unroll
        for (int i = 0; i < 16; ++i) nx[i] = __ldg(T + t + 16 * i);
    }
    for (int f = 0; f < cnt; ++f) {
        u64 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = nx[i];
        if (f + 1 < cnt) {
            const u64* T = scratch + (long long)(f0 + f + 1) * SP_N + (long long)(r0 + g) * SP_N2;
#pragma unroll
            for (int i = 0; i < 16; ++i) nx[i] = __ldg(T + t + 16 * i);
        }
        fft256_core(ex, sm.tw, t, v);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float re = lo2(v[rev4(q)]), im = hi2(v[rev4(q)]);
            const float mag = sqrtf(fmaf(re, re, im * im));
            acc[q] += 6.02059991327962f * __log2f(mag + 1e-10f);  // 20*log10(x) = 20*log10(2)*log2(x)
        }
        __syncwarp();
    }
    __syncthreads();
    float* so = reinterpret_cast<float*>(sm.ex);
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
#endif  // WC_DEV

// pass B5 = B3 made persistent over averaging groups: a CTA walks groups grp, grp + gridDim.y, ... and the row of the
// NEXT frame (also across a group boundary) is always in flight while the current one is transformed. B3 handles one group
// of `avg` (typically 4) frames per CTA, so a quarter of its loads and its table set-up were never overlapped.
__global__ void __launch_bounds__(SP_THREADS, 2) spectrum_pass_b5(const u64* __restrict__ scratch, int avg, int n_frames,
                                                                  int n_groups, float* __restrict__ out) {
    __shared__ SpSmemB3 sm;
    float* const so_buf = reinterpret_cast<float*>(sm.ex);   // output staging shares the exchange area
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * SP_COLS;
    {
        float sn, cs;
        sincospif(-(float)((tid >> 4) * (tid & 15)) * (1.0f / 128.0f), &sn, &cs);
        sm.tw[tid] = make_float2(cs, sn);
    }
    __syncthreads();
    const int g = tid >> 4, t = tid & 15;
    u64* ex = sm.ex + g * SP_STRIDE;
    const long long row_off = (long long)(r0 + g) * SP_N2 + t;
    u64 nx[16];
    int grp = blockIdx.y;
    if (grp < n_groups) {
        const u64* T = scratch + (long long)(grp * avg) * SP_N + row_off;
#pragma unroll
        for (int i = 0; i < 16; ++i) nx[i] = __ldg(T + 16 * i);
    }
    for (; grp < n_groups; grp += gridDim.y) {
        const int f0 = grp * avg;
        const int cnt = min(avg, n_frames - f0);
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
        for (int f = 0; f < cnt; ++f) {
            u64 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = nx[i];
            // next frame of this group, or the first frame of this CTA's next group
            long long nf = -1;
            if (f + 1 < cnt) nf = f0 + f + 1;
            else if (grp + (int)gridDim.y < n_groups) nf = (long long)(grp + gridDim.y) * avg;
            if (nf >= 0) {
                const u64* T = scratch + nf * SP_N + row_off;
#pragma unroll
                for (int i = 0; i < 16; ++i) nx[i] = __ldg(T + 16 * i);
            }
            fft256_core(ex, sm.tw, t, v);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float re = lo2(v[rev4(q)]), im = hi2(v[rev4(q)]);
                const float mag = sqrtf(fmaf(re, re, im * im));
                acc[q] += 6.02059991327962f * __log2f(mag + 1e-10f);  // 20*log10(x) = 20*log10(2)*log2(x)
            }
            __syncwarp();
        }
        __syncthreads();   // every half-warp is done with its exchange area
        const float inv = 1.0f / (float)cnt;
#pragma unroll
        for (int q = 0; q < 16; ++q) so_buf[(t + 16 * q) * 17 + g] = acc[q] * inv;
        __syncthreads();
        float* o = out + (long long)grp * SP_N;
#pragma unroll 4
        for (int j = 0; j < 16; ++j) {
            const int id = tid + SP_THREADS * j;
            const int k2 = id >> 4, row = id & 15;
            const int k = (r0 + row) + SP_N1 * k2;
            o[k ^ (SP_N / 2)] = so_buf[k2 * 17 + row];  // fftshift
        }
        __syncthreads();   // staged output consumed before the next group's exchanges overwrite it
    }
}

#ifdef WC_DEV   // measured and not adopted (profiles/r02_c3_notes.md): 8.9 B/sample of DRAM traffic, but 183 GS/s against 228
// ---- both passes in ONE kernel, the scratch never leaves L2: groups of 16 CTAs walk their frames together ----------------
// The two launches above move 8 (in) + 8 (scratch write) + 8 (scratch read) + 4/K (out) bytes per sample through HBM because a
// slab's scratch does not survive in L2 between them. Here the 16 CTAs that hold the 16 column tiles of a frame in pass A are
// the same 16 that hold its 16 row tiles in pass B; they take every frame of their share of the averaging groups through
// A and B together, exchanging it through a private 4-frame ring (2 MB per group, 36 MB for the 18 groups of a 148-SM grid)
// that is re-dirtied in place every few microseconds and therefore stays in L2. One monotonic arrival counter per ring slot
// orders the hand-over (a single counter per group would not do: a CTA runs up to one frame ahead of the others, so a sum of
// arrivals can reach 16 (j + 1) before every CTA has delivered frame j); the wait for frame j is taken only after pass A of
// frame j + 1 has been issued, so it has a whole A phase (~2 us) to complete in:
//
//     A(0) arrive | A(1) arrive  wait(0) B(0) | A(2) arrive  wait(1) B(1) | ...           (per CTA)
//
// A(j) writes ring slot j & 3. A CTA that is writing slot j has passed wait(j - 2), i.e. every CTA of the group has arrived
// for frame j - 2 and is at the earliest in B(j - 3), which reads slot (j - 3) & 3: four slots, no write-after-read hazard.
// The grid is launched cooperatively (co-residency guaranteed, never more CTAs than fit), ring reads bypass L1 (ld.global.cg:
// L1 is not coherent across SMs and slots are reused).
constexpr int SPG_SLOTS = 4;
struct SpGroupArgs {
    const u64* iq;
    long long frame_stride;
    const float* window;
    u64* ring;           // [CTA groups][SPG_SLOTS][65536]
    unsigned* cnt;       // [CTA groups][SPG_SLOTS] arrivals per ring slot, zero at launch
    int avg, n_frames, n_groups;
    float* out;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(SP_THREADS, 2) spectrum_group_kernel(const SpGroupArgs a) {
    __shared__ SpSmemA3 sm;
    float* const so_buf = reinterpret_cast<float*>(sm.ex);   // output staging shares the exchange area
    const int tid = threadIdx.x;
    const int c = blockIdx.x & 15;       // column tile in pass A, row tile in pass B
    const int cg = blockIdx.x >> 4;      // CTA group
    const int n_cg = gridDim.x >> 4;
    {
        float sn, cs;
        sincospif(-(float)((tid >> 4) * (tid & 15)) * (1.0f / 128.0f), &sn, &cs);
        sm.tw[tid] = make_float2(cs, sn);
    }
    __syncthreads();
    const int t = tid >> 4, col = tid & 15;   // pass A: (row of the 16 x 16 block, column); pass B: (row g = t, lane tb = col)
    const int n2 = c * SP_COLS + col;
    float2 base, step;
    {
        float sn, cs;
        sincospif(-(float)(n2 * t) * (1.0f / 32768.0f), &sn, &cs);
        base = make_float2(cs, sn);
        sincospif(-(float)(n2 * 16) * (1.0f / 32768.0f), &sn, &cs);
        step = make_float2(cs, sn);
    }
    u64* const exA = sm.ex + col * SP_STRIDE;
    u64* const exB = sm.ex + t * SP_STRIDE;
    const float* const wp = a.window + t * SP_N2 + n2;
    u64* const ring = a.ring + (size_t)cg * SPG_SLOTS * SP_N;
    unsigned* const cnt = a.cnt + cg * SPG_SLOTS;
    const long long row_off = (long long)(c * SP_COLS + t) * SP_N2 + col;

    // this group's averaging groups: cg, cg + n_cg, ...; J = its frames in processing order
    int J = 0;
    for (int ag = cg; ag < a.n_groups; ag += n_cg) J += min(a.avg, a.n_frames - ag * a.avg);
    if (J == 0) return;
    int agA = cg, fA = 0;                 // next frame pass A loads
    int agB = cg, fB = 0;                 // next frame pass B takes
    u64 nx[16];
    {
        const u64* x = a.iq + (long long)(agA * a.avg + fA) * a.frame_stride;
#pragma unroll
        for (int i = 0; i < 16; ++i) nx[i] = __ldcs(x + (t + 16 * i) * SP_N2 + n2);
    }
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;

    auto pass_b = [&](int jb) {
        if (tid == 0) {
            const unsigned target = 16u * (unsigned)(jb / SPG_SLOTS + 1);   // frames jb, jb - 4, ... all went through this slot
            const unsigned* cs = cnt + (jb & (SPG_SLOTS - 1));
            const long long t0 = clock64();
            while (ld_acquire_u32(cs) < target) {
                if (clock64() - t0 > (1ll << 31)) __trap();   // ~1 s: cannot happen under a cooperative launch
            }
        }
        __syncthreads();
        const u64* T = ring + (size_t)(jb & (SPG_SLOTS - 1)) * SP_N + row_off;
        u64 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __ldcg(T + 16 * i);
        fft256_core(exB, sm.tw, col, v);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float re = lo2(v[rev4(q)]), im = hi2(v[rev4(q)]);
            const float mag = sqrtf(fmaf(re, re, im * im));
            acc[q] += 6.02059991327962f * __log2f(mag + 1e-10f);  // 20*log10(x) = 20*log10(2)*log2(x)
        }
        const int cntB = min(a.avg, a.n_frames - agB * a.avg);
        if (++fB == cntB) {               // averaging group complete: mean, fftshift, store
            __syncthreads();              // every half-warp is done with its exchange area
            const float inv = 1.0f / (float)cntB;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                so_buf[(col + 16 * q) * 17 + t] = acc[q] * inv;
                acc[q] = 0.f;
            }
            __syncthreads();
            float* o = a.out + (long long)agB * SP_N;
#pragma unroll 4
            for (int j = 0; j < 16; ++j) {
                const int id = tid + SP_THREADS * j;
                const int k2 = id >> 4, row = id & 15;
                const int k = (c * SP_COLS + row) + SP_N1 * k2;
                o[k ^ (SP_N / 2)] = so_buf[k2 * 17 + row];  // fftshift
            }
            agB += n_cg;
            fB = 0;
        }
    };

    for (int j = 0; j < J; ++j) {
        // ---- pass A of frame j (column tile c): window, column FFT-256, four-step twiddle, into ring slot j & 3
        u64 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = mul2(nx[i], bc2(__ldg(wp + 16 * i * SP_N2)));
        if (++fA == min(a.avg, a.n_frames - agA * a.avg)) {
            agA += n_cg;
            fA = 0;
        }
        if (j + 1 < J) {
            const u64* x = a.iq + (long long)(agA * a.avg + fA) * a.frame_stride;
#pragma unroll
            for (int i = 0; i < 16; ++i) nx[i] = __ldcs(x + (t + 16 * i) * SP_N2 + n2);
        }
        fft16(v);
        __syncthreads();   // the previous phase's exchange reads and staged output are consumed
#pragma unroll
        for (int ka = 0; ka < 16; ++ka) {
            u64 w = v[rev4(ka)];
            if (ka > 0) {
                const float2 q = sm.tw[ka * 16 + t];
                w = twid(w, q.x, -q.y);
            }
            exA[t * 17 + ka] = w;
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = exA[m * 17 + t];
        fft16(v);
        {
            u64* T = ring + (size_t)(j & (SPG_SLOTS - 1)) * SP_N;
            float2 w = base;
#pragma unroll
            for (int kb = 0; kb < 16; ++kb) {
                const int k1 = t + 16 * kb;
                T[k1 * SP_N2 + n2] = twid(v[rev4(kb)], w.x, -w.y);
                w = cmul(w, step);
            }
        }
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(cnt + (j & (SPG_SLOTS - 1)), 1u);
        }
        // ---- pass B of frame j - 1 (row tile c)
        if (j >= 1) pass_b(j - 1);
    }
    pass_b(J - 1);
}
#endif  // WC_DEV

#ifdef WC_DEV
// ---- fused persistent version: both passes in ONE kernel, scratch is a small L2-resident ring --------
// The two-pass design moves 8 (in) + 8 (scratch write) + 8 (scratch read) + 4/K (out) bytes per sample through
// HBM because a slab's scratch does not survive in L2 between two launches. Here a single persistent grid pulls
// work items from an in-order ticket counter:
//     step s:  A items (column tile x frame) of averaging group s,  then  B items (row tile) of group s - LAG
// so pass B of a frame runs a few dozen microseconds after its pass A, while the frame's 512 KB are still in L2,
// and the scratch is a ring of R frames that is re-dirtied in place instead of being written back. Dependencies
// are per-frame / per-group completion counters in global memory: a B item spins until the 16 column tiles of its
// frames are done, an A item spins until the previous tenant of its ring slot has been consumed. Tickets are
// handed out in dependency order and a CTA only takes a ticket when it is running, so every awaited item is
// either finished or owned by a resident CTA: no co-residency requirement, no deadlock.
// Ring data is read with ld.global.cg (L2 only): L1 is not coherent across SMs and ring slots are reused.
struct SpFusedArgs {
    const u64* iq;
    long long frame_stride;
    const float* window;
    u64* ring;           // [R][65536]
    int R;               // ring frames
    int avg, n_frames, n_groups, lag;   // lag in groups
    float* out;
    int* ctrl;           // [0] ticket, [1 .. 1+n_frames) a_done per frame, then b_done per group
};

__device__ __forceinline__ void sp_spin_until(const int* ctr, int target) {
    const volatile int* v = ctr;
    while (*v < target) __nanosleep(64);
    __threadfence();
}

__global__ void __launch_bounds__(SP_THREADS, 3) spectrum_fused_kernel(const SpFusedArgs a) {
    __shared__ SpSmemA3 sm;
    __shared__ int s_ticket;
    const int tid = threadIdx.x;
    {
        float sn, cs;
        sincospif(-(float)((tid >> 4) * (tid & 15)) * (1.0f / 128.0f), &sn, &cs);
        sm.tw[tid] = make_float2(cs, sn);
    }
    int* a_done = a.ctrl + 1;
    int* b_done = a.ctrl + 1 + a.n_frames;
    const int per_step = a.avg * 16 + 16;
    const int n_steps = a.n_groups + a.lag;
    if (tid == 0) s_ticket = atomicAdd(a.ctrl, 1);
    for (;;) {
        __syncthreads();                 // s_ticket visible; previous item's shared-memory traffic finished
        const int ticket = s_ticket;
        __syncthreads();
        if (tid == 0) s_ticket = atomicAdd(a.ctrl, 1);   // next ticket is fetched while this item runs
        const int step = ticket / per_step, slot = ticket - step * per_step;
        if (step >= n_steps) break;
        if (slot < a.avg * 16) {
            // ---------------- pass A item: frame fr, column tile ct ----------------
            const int fr = step * a.avg + (slot >> 4);
            if (step >= a.n_groups || fr >= a.n_frames) continue;
            const int c0 = (slot & 15) * SP_COLS;
            if (fr >= a.R) {
                if (tid == 0) sp_spin_until(b_done + (fr - a.R) / a.avg, 16);
                __syncthreads();
            }
            const int t = tid >> 4, col = tid & 15;
            const int n2 = c0 + col;
            float2 base, stp;
            {
                float sn, cs;
                sincospif(-(float)(n2 * t) * (1.0f / 32768.0f), &sn, &cs);
                base = make_float2(cs, sn);
                sincospif(-(float)(n2 * 16) * (1.0f / 32768.0f), &sn, &cs);
                stp = make_float2(cs, sn);
            }
            u64* ex = sm.ex + col * SP_STRIDE;
            const u64* x = a.iq + (long long)fr * a.frame_stride;
            u64 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int n = (t + 16 * i) * SP_N2 + n2;
                v[i] = mul2(__ldcs(x + n), bc2(__ldg(a.window + n)));   // input is streamed once: evict-first
            }
            fft16(v);
#pragma unroll
            for (int ka = 0; ka < 16; ++ka) {
                u64 w = v[rev4(ka)];
                if (ka > 0) {
                    const float2 q = sm.tw[ka * 16 + t];
                    w = twid(w, q.x, -q.y);
                }
                ex[t * 17 + ka] = w;
            }
            __syncthreads();
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = ex[m * 17 + t];
            fft16(v);
            u64* T = a.ring + (long long)(fr % a.R) * SP_N;
            float2 w = base;
#pragma unroll
            for (int kb = 0; kb < 16; ++kb) {
                const int k1 = t + 16 * kb;
                T[k1 * SP_N2 + n2] = twid(v[rev4(kb)], w.x, -w.y);
                w = cmul(w, stp);
            }
            __syncthreads();
            if (tid == 0) {
                __threadfence();         // cumulative: orders the whole CTA's ring stores (observed via the barrier)
                atomicAdd(a_done + fr, 1);
            }
        } else {
            // ---------------- pass B item: group grp, row tile rt ----------------
            const int grp = step - a.lag;
            if (grp < 0 || grp >= a.n_groups) continue;
            const int r0 = (slot - a.avg * 16) * SP_COLS;
            const int f0 = grp * a.avg;
            const int cnt = min(a.avg, a.n_frames - f0);
            const int g = tid >> 4, t = tid & 15;
            float acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.f;
            u64* ex = sm.ex + g * SP_STRIDE;
            for (int f = 0; f < cnt; ++f) {
                if (tid == 0) sp_spin_until(a_done + f0 + f, 16);
                __syncthreads();
                const u64* T = a.ring + (long long)((f0 + f) % a.R) * SP_N + (long long)(r0 + g) * SP_N2;
                u64 v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __ldcg(T + t + 16 * i);
                fft256_core(ex, sm.tw, t, v);
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float re = lo2(v[rev4(q)]), im = hi2(v[rev4(q)]);
                    const float mag = sqrtf(fmaf(re, re, im * im));
                    acc[q] += 6.02059991327962f * __log2f(mag + 1e-10f);
                }
                __syncwarp();
            }
            __syncthreads();
            float* so = reinterpret_cast<float*>(sm.ex);
            const float inv = 1.0f / (float)cnt;
#pragma unroll
            for (int q = 0; q < 16; ++q) so[(t + 16 * q) * 17 + g] = acc[q] * inv;
            __syncthreads();
            float* o = a.out + (long long)grp * SP_N;
#pragma unroll 4
            for (int j = 0; j < 16; ++j) {
                const int id = tid + SP_THREADS * j;
                const int k2 = id >> 4, row = id & 15;
                const int k = (r0 + row) + SP_N1 * k2;
                o[k ^ (SP_N / 2)] = so[k2 * 17 + row];  // fftshift
            }
            __syncthreads();   // ring reads of this item are complete (they were consumed into registers above)
            if (tid == 0) {
                __threadfence();
                atomicAdd(b_done + grp, 1);
            }
        }
    }
}
#endif  // WC_DEV

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
    void* d_ring = nullptr;
    size_t ring_bytes = 0;
    void* d_ctrl = nullptr;
    size_t ctrl_bytes = 0;
    void* d_in = nullptr;
    size_t in_bytes = 0;
    void* d_out = nullptr;
    size_t out_bytes = 0;
    cudaStream_t stream = nullptr;
    // variant 5: pass A and pass B on two internal streams over a ring of small slabs
    static constexpr int PIPE_Q = 3;
    cudaStream_t s_a = nullptr, s_b = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_a[PIPE_Q] = {}, ev_b[PIPE_Q] = {}, ev_join_a = nullptr, ev_join_b = nullptr;
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
#ifdef WC_DEV
    cudaFuncSetAttribute(spectrum_pass_a, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(spectrum_pass_b, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(spectrum_pass_a2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpSmemA));
    cudaFuncSetAttribute(spectrum_pass_b2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpSmemB));
    cudaFuncSetAttribute(spectrum_pass_a2, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(spectrum_pass_a3, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(spectrum_fused_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(spectrum_pass_b3, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(spectrum_pass_b2, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
#endif
    *out = h;
    return 0;
}

void wc_spectrum_destroy(wc_spectrum* h) {
    if (!h) return;
    cudaFree(h->d_window);
    if (h->d_scratch) cudaFree(h->d_scratch);
    if (h->d_ring) cudaFree(h->d_ring);
    if (h->d_ctrl) cudaFree(h->d_ctrl);
    if (h->d_in) cudaFree(h->d_in);
    if (h->d_out) cudaFree(h->d_out);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->s_a) cudaStreamDestroy(h->s_a);
    if (h->s_b) cudaStreamDestroy(h->s_b);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join_a) cudaEventDestroy(h->ev_join_a);
    if (h->ev_join_b) cudaEventDestroy(h->ev_join_b);
    for (int q = 0; q < wc_spectrum::PIPE_Q; ++q) {
        if (h->ev_a[q]) cudaEventDestroy(h->ev_a[q]);
        if (h->ev_b[q]) cudaEventDestroy(h->ev_b[q]);
    }
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
    // process in slabs of <= 512 MB of scratch. Measured on B200 (profiles/r01_spectrum_notes.md): the scratch does not
    // stay L2-resident at any useful slab size (ncu: pass A writes it to DRAM), small slabs only starve the grid
    // (16 MB: 79 GS/s, 64 MB: 124, 512 MB: 150), so slabs are sized for parallelism.
    size_t slab_bytes = 512ull << 20;
    slab_bytes = (size_t)env_int("WC_SPECTRUM_SLAB_MB", 512) << 20;
    int slab = (int)(slab_bytes / (sizeof(float2) * (size_t)n));
    if (slab < avg) slab = avg;
    slab -= slab % avg;
    // 3 = two register-direct passes on the caller's stream, 4 = fused ring kernel, 5 = the same two passes pipelined over
    // two internal streams (default once a call spans more than one 128 MB slab: 210 -> 229 GS/s at 4096 frames)
    const size_t pipe_slab_bytes = (size_t)env_int("WC_SPECTRUM_PIPE_SLAB_MB", 128) << 20;
    // (a call of two or three slabs gains nothing from the pipeline: 368 frames 200 GS/s single-stream, 179 pipelined)
    const int variant = env_int("WC_SPECTRUM_VARIANT", (sizeof(float2) * (size_t)n * n_frames >= 4 * pipe_slab_bytes) ? 5 : 3);
#ifdef WC_DEV
    if (n == SP_N && variant == 6) {
        // one cooperative launch: groups of 16 CTAs, scratch = their private L2-resident rings
        int occ = 0;
        WC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spectrum_group_kernel, SP_THREADS, 0));
        int n_cg = (occ * sm_count()) / 16;
        const int groups = (n_frames + avg - 1) / avg;
        if (n_cg > groups) n_cg = groups;
        WC_REQUIRE(n_cg >= 1, "wc_spectrum_execute: the group kernel does not fit this device");
        if (sp_ensure(&h->d_ring, &h->ring_bytes, sizeof(float2) * (size_t)n * SPG_SLOTS * n_cg)) return -2;
        if (sp_ensure(&h->d_ctrl, &h->ctrl_bytes, sizeof(unsigned) * 1024)) return -2;
        WC_CUDA(cudaMemsetAsync(h->d_ctrl, 0, sizeof(unsigned) * SPG_SLOTS * n_cg, st));
        SpGroupArgs ga;
        ga.iq = reinterpret_cast<const u64*>(iq);
        ga.frame_stride = frame_stride;
        ga.window = h->d_window;
        ga.ring = reinterpret_cast<u64*>(h->d_ring);
        ga.cnt = reinterpret_cast<unsigned*>(h->d_ctrl);
        ga.avg = avg;
        ga.n_frames = n_frames;
        ga.n_groups = groups;
        ga.out = power_db_dev;
        void* args[] = {&ga};
        WC_CUDA(cudaLaunchCooperativeKernel((const void*)spectrum_group_kernel, dim3(16 * n_cg), dim3(SP_THREADS), args, 0, st));
        return 0;
    }
#endif
    if (n == SP_N && variant == 5) {
        // Two-stream pipeline: pass A of slab k+1 runs while pass B of slab k drains, over a ring of PIPE_Q slabs: the
        // tail of one launch is filled by the head of the other instead of idling the SMs. Measured on B200 (4096 frames,
        // profiles/r01_spectrum_notes.md): 8 MB slabs 89 GS/s ... 64 MB 225, 128 MB 229, 512 MB 218; single-stream 210.
        constexpr int Q = wc_spectrum::PIPE_Q;
        int sl = (int)(pipe_slab_bytes / (sizeof(float2) * (size_t)n));
        if (sl < avg) sl = avg;
        sl -= sl % avg;
        {   // equal slabs: a short last slab would end the pipeline on an unoverlapped tail
            const int n_slabs = (n_frames + sl - 1) / sl;
            int even = (n_frames + n_slabs - 1) / n_slabs;
            even = ((even + avg - 1) / avg) * avg;
            if (even < sl) sl = even;
        }
        if (sp_ensure(&h->d_scratch, &h->scratch_bytes, sizeof(float2) * (size_t)n * sl * Q)) return -2;
        if (!h->s_a) {
            WC_CUDA(cudaStreamCreateWithFlags(&h->s_a, cudaStreamNonBlocking));
            WC_CUDA(cudaStreamCreateWithFlags(&h->s_b, cudaStreamNonBlocking));
            WC_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
            WC_CUDA(cudaEventCreateWithFlags(&h->ev_join_a, cudaEventDisableTiming));
            WC_CUDA(cudaEventCreateWithFlags(&h->ev_join_b, cudaEventDisableTiming));
            for (int q = 0; q < Q; ++q) {
                WC_CUDA(cudaEventCreateWithFlags(&h->ev_a[q], cudaEventDisableTiming));
                WC_CUDA(cudaEventCreateWithFlags(&h->ev_b[q], cudaEventDisableTiming));
            }
        }
        WC_CUDA(cudaEventRecord(h->ev_fork, st));
        WC_CUDA(cudaStreamWaitEvent(h->s_a, h->ev_fork, 0));
        WC_CUDA(cudaStreamWaitEvent(h->s_b, h->ev_fork, 0));
        int k = 0;
        for (int f0 = 0; f0 < n_frames; f0 += sl, ++k) {
            const int cnt = (n_frames - f0 < sl) ? n_frames - f0 : sl;
            const int groups = (cnt + avg - 1) / avg;
            float* o = power_db_dev + (long long)(f0 / avg) * n;
            const float2* x = iq + (long long)f0 * frame_stride;
            u64* T = reinterpret_cast<u64*>(h->d_scratch) + (size_t)(k % Q) * (size_t)n * sl;
            if (k >= Q) WC_CUDA(cudaStreamWaitEvent(h->s_a, h->ev_b[k % Q], 0));   // pass B of slab k - Q has left the slot
            int fy = (6 * sm_count()) / 16;
            if (fy > cnt) fy = cnt;
            spectrum_pass_a5<2><<<dim3(SP_N2 / SP_COLS, fy), SP_THREADS, 0, h->s_a>>>(x, frame_stride, h->d_window, T, cnt);
            WC_CUDA(cudaEventRecord(h->ev_a[k % Q], h->s_a));
            WC_CUDA(cudaStreamWaitEvent(h->s_b, h->ev_a[k % Q], 0));
            int gy = (4 * sm_count()) / 16;
            if (gy > groups) gy = groups;
            spectrum_pass_b5<<<dim3(SP_N1 / SP_COLS, gy), SP_THREADS, 0, h->s_b>>>(T, avg, cnt, groups, o);
            WC_CUDA(cudaEventRecord(h->ev_b[k % Q], h->s_b));
            WC_CUDA(cudaGetLastError());
        }
        WC_CUDA(cudaEventRecord(h->ev_join_a, h->s_a));
        WC_CUDA(cudaEventRecord(h->ev_join_b, h->s_b));
        WC_CUDA(cudaStreamWaitEvent(st, h->ev_join_a, 0));
        WC_CUDA(cudaStreamWaitEvent(st, h->ev_join_b, 0));
        return 0;
    }
#ifdef WC_DEV
    const bool fused = (n == SP_N && variant == 4);
#else
    const bool fused = false;
#endif
    if (fused) slab = 1 << 20;   // the fused kernel's scratch is a ring: no slab limit (ctrl is 4 bytes per frame)
    if (slab > n_frames) slab = ((n_frames + avg - 1) / avg) * avg;
    if (!fused) {
        const size_t need = sizeof(float2) * (size_t)n * slab * (n == SP_N ? 1 : 2);
        if (sp_ensure(&h->d_scratch, &h->scratch_bytes, need)) return -2;
    }
    for (int f0 = 0; f0 < n_frames; f0 += slab) {
        const int cnt = (n_frames - f0 < slab) ? n_frames - f0 : slab;
        const int groups = (cnt + avg - 1) / avg;
        float* o = power_db_dev + (long long)(f0 / avg) * n;
        const float2* x = iq + (long long)f0 * frame_stride;
        if (n == SP_N) {
            u64* T = reinterpret_cast<u64*>(h->d_scratch);
#ifdef WC_DEV
            if (variant == 4) {
                // fused persistent kernel: ring of R frames, pass B lags pass A by `lag` groups
                int lagf = 32;
                lagf = env_int("WC_SPECTRUM_LAG", lagf);
                int lag = lagf / avg;
                if (lag < 1) lag = 1;
                int R = 2 * (lag + 1) * avg;
                if (R < 64) R = 64;
                R = env_int("WC_SPECTRUM_RING", R);
                if (R < (lag + 2) * avg) R = (lag + 2) * avg;
                if (sp_ensure(&h->d_ring, &h->ring_bytes, sizeof(float2) * (size_t)n * R)) return -2;
                const size_t ctrl_need = sizeof(int) * (size_t)(1 + cnt + groups);
                if (sp_ensure(&h->d_ctrl, &h->ctrl_bytes, ctrl_need)) return -2;
                WC_CUDA(cudaMemsetAsync(h->d_ctrl, 0, ctrl_need, st));
                SpFusedArgs fa;
                fa.iq = reinterpret_cast<const u64*>(x);
                fa.frame_stride = frame_stride;
                fa.window = h->d_window;
                fa.ring = reinterpret_cast<u64*>(h->d_ring);
                fa.R = R;
                fa.avg = avg;
                fa.n_frames = cnt;
                fa.n_groups = groups;
                fa.lag = lag;
                fa.out = o;
                fa.ctrl = reinterpret_cast<int*>(h->d_ctrl);
                spectrum_fused_kernel<<<3 * sm_count(), SP_THREADS, 0, st>>>(fa);
            } else if (variant == 1) {
                spectrum_pass_a<<<dim3(SP_N2 / SP_COLS, cnt), SP_THREADS, 0, st>>>(x, frame_stride, h->d_window, T);
                spectrum_pass_b<<<dim3(SP_N1 / SP_COLS, groups), SP_THREADS, 0, st>>>(T, avg, cnt, o);
            } else if (variant == 2) {
                // persistent grids: 2 (pass A) / 3 (pass B) CTAs per SM, each double-buffering its tiles
                int ga = (2 * sm_count()) / 16;
                if (ga > cnt) ga = cnt;
                if (ga < 1) ga = 1;
                spectrum_pass_a2<<<16 * ga, SP_THREADS, sizeof(SpSmemA), st>>>(x, frame_stride, h->d_window, T, cnt);
                int gb = 3 * sm_count();
                if (gb > groups * 16) gb = groups * 16;
                spectrum_pass_b2<<<gb, SP_THREADS, sizeof(SpSmemB), st>>>(T, avg, cnt, groups, o);
            } else
#endif
            {
                // two register-direct passes on the caller's stream
                int fy = (6 * sm_count()) / 16;   // frames in flight: ~6 CTAs per SM over the 16 column tiles
                fy = env_int("WC_SPECTRUM_FY", fy);
                if (fy > cnt) fy = cnt;
                if (fy < 1) fy = 1;
                // measured on B200, 4096 frames: A3 153 GS/s, A5<3> 166, A5<2> 200
#ifdef WC_DEV
                const int pav = env_int("WC_SPECTRUM_PASS_A", 6);
                if (pav == 5) spectrum_pass_a5<3><<<dim3(SP_N2 / SP_COLS, fy), SP_THREADS, 0, st>>>(x, frame_stride, h->d_window, T, cnt);
                else if (pav != 6) spectrum_pass_a3<<<dim3(SP_N2 / SP_COLS, fy), SP_THREADS, 0, st>>>(x, frame_stride, h->d_window, T, cnt);
                else
#endif
                spectrum_pass_a5<2><<<dim3(SP_N2 / SP_COLS, fy), SP_THREADS, 0, st>>>(x, frame_stride, h->d_window, T, cnt);
                // measured on B200, 4096 frames with pass A5<2>: B3 200 GS/s, B5 210
#ifdef WC_DEV
                if (env_int("WC_SPECTRUM_PASS_B", 5) != 5) {
                    spectrum_pass_b3<<<dim3(SP_N1 / SP_COLS, groups), SP_THREADS, 0, st>>>(T, avg, cnt, o);
                } else
#endif
                {
                    int gy = (4 * sm_count()) / 16;   // two waves of the 2 resident CTAs per SM over the 16 row tiles
                    gy = env_int("WC_SPECTRUM_GY", gy);
                    if (gy > groups) gy = groups;
                    if (gy < 1) gy = 1;
                    spectrum_pass_b5<<<dim3(SP_N1 / SP_COLS, gy), SP_THREADS, 0, st>>>(T, avg, cnt, groups, o);
                }
            }
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
