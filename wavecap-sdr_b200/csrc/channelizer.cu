// Polyphase channelizer (NMDPFB, 2x oversampled) — B200-native.
//
// Replaces the per-frame Python loop of wavecapsdr/dsp/channelizer.py:91-137
// (np.roll + einsum + np.fft.fft per frame) with one grid-wide pass:
//
//   u_b[k] = sum_{j=0..T-1} h[k + j*M] * blk_{b-j}[k],   blk_i[k] = x[i*M/2 + k]
//   y_b    = FFT_M(u_b)                                  (forward, unnormalised, FFT bin order)
//   fused FM: d_b[k] = angle(y_b[k] * conj(y_{b-1}[k])) * scale, d_0 = 0   (dsp/fm.py:65-97)
//
// Fast path (M = 256, T = 9 — the 125 MS/s / 256-channel workload): each CTA owns a contiguous
// run of frames. Per sub-tile of 8 frames:
//   phase 1  FIR: thread r keeps a 10-sample sliding window x[q*128 + r] in registers (both
//            u_b[r] and u_b[r+128] come from the same residue class), new rows arrive in shared
//            memory through a double-buffered 1-D bulk async copy (TMA engine) so every input
//            sample is read from HBM exactly once per run.
//   phase 2  256-point FFT as 16 x 16: 16 threads per frame, two in-register radix-16 passes,
//            one padded shared-memory exchange.
//   phase 3  (FM mode) per-bin discriminator with the previous frame's bin kept in registers,
//            coalesced float2 stores.
// Generic path (any even M): FIR kernel + direct O(M^2) DFT kernel; correct, not tuned.
#include <math.h>
#include <stdlib.h>
#include <vector>

#include "../../include/wcsdr_b200.h"
#include "common.cuh"
#include "fft16.cuh"

namespace wc {

constexpr int CH_M = 256;
constexpr int CH_H = 128;
constexpr int CH_T = 9;
constexpr int CH_S = 8;          // frames per sub-tile
constexpr int CH_THREADS = 128;  // one thread per residue class
constexpr int CH_REGION = 280;   // complex words per frame region (16 x 17 padded exchange + bank offset)

// input sample formats: complex64, or interleaved int16 I,Q scaled by 1/32768 like cli.py:449-453 (an exact float32 value)
constexpr int IN_CF32 = 0, IN_CS16 = 1;
constexpr float CS16_SCALE = 1.0f / 32768.0f;
__device__ __forceinline__ u64 cs16_to_pair(uint32_t w) {
    return pk2((float)(short)(w & 0xffffu) * CS16_SCALE, (float)(short)(w >> 16) * CS16_SCALE);
}
// sample `idx` of an input array (global or shared) of format FMT as a packed (re, im) float pair
template <int FMT>
__device__ __forceinline__ u64 in_pair(const void* base, long long idx) {
    if (FMT == IN_CS16) return cs16_to_pair(reinterpret_cast<const uint32_t*>(base)[idx]);
    return reinterpret_cast<const u64*>(base)[idx];
}
template <int FMT>
__device__ __forceinline__ const void* in_offset(const void* base, long long samples) {
    return reinterpret_cast<const char*>(base) + samples * (FMT == IN_CS16 ? 4 : 8);
}

struct ChanArgs {
    const void* x;         // [n_chunks][chunk_stride] complex64 (IN_CF32) or int16 pairs (IN_CS16)
    long long chunk_stride;
    int F;                 // frames per chunk
    int R;                 // frames per CTA run (multiple of 8, >= 16)
    const float* taps;     // [M*T] prototype, h[k + j*M], zero padded
    const float2* carried; // [9][256] blk_{-9..-1} for chunk 0
    void* out;             // MODE 0: float2 [n_chunks*F][256]; MODE 1: float [n_chunks*F][256]
    float scale;           // FM discriminator scale
    AtanScaled at;         // scale folded into the atan2 polynomial (fused FM mode)
};

constexpr int CH_REGION_W = 2 * CH_REGION;  // 32-bit words per frame region
static_assert(CH_REGION_W % 32 == 16, "frame regions must be offset by 16 banks (planar Y stores)");
constexpr int CH_YIM = 272;                 // word offset of the Im plane inside a frame region

struct __align__(128) ChanSmem {
    u64 stage[2][CH_S * CH_H];   // TMA landing buffers: 8 rows of 128 cf32 samples each
    u64 u[CH_S * CH_REGION];     // FIR output -> FFT exchange -> planar FFT output (in place)
    float2 tw[16 * 16];          // tw[k1*16 + t] = exp(-2*pi*i*k1*t/256)
    uint64_t full[2];
};

// block value for the slow prologue path: blk_i[k] of chunk c, i may be negative (history)
template <int FMT = IN_CF32>
__device__ __forceinline__ float2 blk_fetch(const ChanArgs& a, const void* xc, int c, int i, int k) {
    u64 v;
    if (i >= 0) v = in_pair<FMT>(xc, (long long)i * CH_H + k);
    else if (c > 0) v = in_pair<FMT>(xc, (long long)(a.F + i) * CH_H + k - a.chunk_stride);
    else return a.carried[(i + CH_T) * CH_M + k];
    return make_float2(lo2(v), hi2(v));
}

#ifdef WC_DEV   // phase-serial predecessor of chan256p_kernel: dev builds only, for A/B timing (WC_CHAN_VAR=0)
// ABL (dev builds only, -DWC_DEV_ABLATE): timing ablations that skip one phase; results are then wrong by design.
template <int MODE, int MINB, int ABL = 0>
__global__ void __launch_bounds__(CH_THREADS, MINB) chan256_kernel(const ChanArgs a) {
    __shared__ ChanSmem sm;
    const int tid = threadIdx.x;
    const int r = tid;
    const int c = blockIdx.y;
    // run i emits frames [f0, f1); in FM mode runs i > 0 also compute frame f0-1 (needs y_{f0-1}),
    // so they emit R-1 frames and every run computes a multiple of 8 frames.
    const int step = (MODE == 1) ? a.R - 1 : a.R;
    const int f0 = (blockIdx.x == 0) ? 0 : a.R + (blockIdx.x - 1) * step;
    if (f0 >= a.F) return;
    const int f1 = min(a.F, (blockIdx.x == 0) ? a.R : f0 + step);
    const float2* __restrict__ xc = reinterpret_cast<const float2*>(a.x) + (long long)c * a.chunk_stride;   // cf32 input only
    const long long out_base = (long long)c * a.F;

    // one warm-up frame so the discriminator has y_{f0-1}
    const int fe = (MODE == 1 && f0 > 0) ? f0 - 1 : f0;
    const int fast_start = (f0 == 0) ? 8 : fe;
    const int n_fast = (fast_start < f1) ? (f1 - fast_start + CH_S - 1) / CH_S : 0;

    if (tid == 0) {
        mbar_init(&sm.full[0], 1);
        mbar_init(&sm.full[1], 1);
        mbar_fence_init();
    }
    for (int i = tid; i < 256; i += CH_THREADS) {
        float s, co;
        sincospif(-(float)((i >> 4) * (i & 15)) * (1.0f / 128.0f), &s, &co);
        sm.tw[i] = make_float2(co, s);
    }
    __syncthreads();

    auto issue = [&](int n) {
        const int fs = fast_start + CH_S * n;
        const int nrows = min(CH_S, a.F - fs);  // rows fs+1 .. fs+nrows (row F is the last one)
        const uint32_t bytes = (uint32_t)nrows * CH_H * sizeof(float2);
        mbar_expect_tx(&sm.full[n & 1], bytes);
        bulk_g2s(sm.stage[n & 1], xc + (long long)(fs + 1) * CH_H, bytes, &sm.full[n & 1]);
    };
    if (tid == 0) {
        if (n_fast > 0) issue(0);
        if (n_fast > 1) issue(1);
    }

    float hlo[CH_T], hhi[CH_T];
#pragma unroll
    for (int j = 0; j < CH_T; ++j) {
        hlo[j] = __ldg(a.taps + r + CH_M * j);
        hhi[j] = __ldg(a.taps + r + CH_H + CH_M * j);
    }

    // discriminator state: previous frame's bins 2*tid, 2*tid+1 as (re0,re1) / (im0,im1) pairs
    u64 pre = 0ull, pim = 0ull;
    float* const sw = reinterpret_cast<float*>(sm.u);

    // phases 2 + 3 for the sub-tile whose FIR outputs sit in sm.u (frames fs .. fs+nv-1)
    auto fft_and_emit = [&](int fs, int nv) {
        if (ABL != 2) {
            const int g = tid >> 4, t = tid & 15;
            u64* reg = sm.u + g * CH_REGION;
            u64 v[16];
            if (g < nv) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = reg[t + 16 * i];
            }
            __syncwarp();
            if (g < nv) {
                fft16(v);
#pragma unroll
                for (int k1 = 0; k1 < 16; ++k1) {
                    u64 w = v[rev4(k1)];
                    if (k1 > 0) {
                        const float2 tw = sm.tw[k1 * 16 + t];
                        w = twid(w, tw.x, -tw.y);
                    }
                    reg[t * 17 + k1] = w;
                }
            }
            __syncwarp();
            if (g < nv) {
#pragma unroll
                for (int n2 = 0; n2 < 16; ++n2) v[n2] = reg[n2 * 17 + t];
            }
            __syncwarp();
            if (g < nv) {
                fft16(v);
                if (MODE == 0) {
                    const int b = fs + g;
                    if (b >= f0) {
                        u64* o = reinterpret_cast<u64*>(a.out) + (out_base + b) * CH_M + t;
#pragma unroll
                        for (int k2 = 0; k2 < 16; ++k2) o[16 * k2] = v[rev4(k2)];
                    }
                } else {
                    // planar: Re plane at words [0,256), Im plane at [CH_YIM, CH_YIM+256)
                    float* w = sw + g * CH_REGION_W + t;
#pragma unroll
                    for (int k2 = 0; k2 < 16; ++k2) {
                        w[16 * k2] = lo2(v[rev4(k2)]);
                        w[CH_YIM + 16 * k2] = hi2(v[rev4(k2)]);
                    }
                }
            }
        }
        __syncthreads();
        if (MODE == 1) {
            float* o = reinterpret_cast<float*>(a.out) + (out_base + fs) * CH_M + 2 * tid;
            const AtanScaled at = a.at;
            auto one = [&](int i) {
                const float* w = sw + i * CH_REGION_W + 2 * tid;
                const u64 yre = *reinterpret_cast<const u64*>(w);
                const u64 yim = *reinterpret_cast<const u64*>(w + CH_YIM);
                // p = y * conj(prev)
                const u64 px = fma2(yre, pre, mul2(yim, pim));
                const u64 py = sub2(mul2(yim, pre), mul2(yre, pim));
                const u64 d = (ABL == 3) ? add2(py, px) : scaled_atan2f_x2(py, px, at);
                if (ABL == 4) {
                    if (lo2(d) == 123.456f) *reinterpret_cast<u64*>(o + (long long)i * CH_M) = d;
                } else if (fs + i >= f0) *reinterpret_cast<u64*>(o + (long long)i * CH_M) = d;
                pre = yre;
                pim = yim;
            };
            if (nv == CH_S) {
#pragma unroll
                for (int i = 0; i < CH_S; ++i) one(i);
            } else {
                for (int i = 0; i < nv; ++i) one(i);
            }
            __syncthreads();
        }
    };

    // ---- prologue sub-tile: frames 0..7 of a chunk reach into the previous call/chunk ----
    if (f0 == 0) {
        const int nv = min(CH_S, f1);
        for (int b = 0; b < nv; ++b) {
            float2 lo = make_float2(0.f, 0.f), hi = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < CH_T; ++j) {
                const float2 vl = blk_fetch(a, xc, c, b - j, r);
                const float2 vh = blk_fetch(a, xc, c, b - j, r + CH_H);
                lo.x = fmaf(hlo[j], vl.x, lo.x);
                lo.y = fmaf(hlo[j], vl.y, lo.y);
                hi.x = fmaf(hhi[j], vh.x, hi.x);
                hi.y = fmaf(hhi[j], vh.y, hi.y);
            }
            sm.u[b * CH_REGION + r] = pk2(lo.x, lo.y);
            sm.u[b * CH_REGION + r + CH_H] = pk2(hi.x, hi.y);
        }
        __syncthreads();
        fft_and_emit(0, nv);
    }
    if (n_fast == 0) return;

    // ---- fast path: sliding window over rows fast_start-8 .. ----
    u64 w[10];
    {
        const u64* xr = reinterpret_cast<const u64*>(xc);
#pragma unroll
        for (int m = 0; m < 9; ++m) w[m] = __ldg(xr + (long long)(fast_start - 8 + m) * CH_H + r);
    }

    for (int n = 0; n < n_fast; ++n) {
        const int fs = fast_start + CH_S * n;
        const int nv = min(CH_S, f1 - fs);
        mbar_wait(&sm.full[n & 1], (n >> 1) & 1);
        const u64* st = sm.stage[n & 1];
#pragma unroll
        for (int i = 0; i < CH_S; ++i) {
            w[9] = st[i * CH_H + r];
            u64 lo = mul2(w[8], bc2(hlo[0])), hi = mul2(w[9], bc2(hhi[0]));
            if (ABL != 1) {
#pragma unroll
                for (int j = 1; j < CH_T; ++j) {
                    lo = fma2(w[8 - j], bc2(hlo[j]), lo);
                    hi = fma2(w[9 - j], bc2(hhi[j]), hi);
                }
            }
            sm.u[i * CH_REGION + r] = lo;
            sm.u[i * CH_REGION + r + CH_H] = hi;
#pragma unroll
            for (int m = 0; m < 9; ++m) w[m] = w[m + 1];
        }
        __syncthreads();
        if (tid == 0 && n + 2 < n_fast) issue(n + 2);
        fft_and_emit(fs, nv);
    }
}
#endif  // WC_DEV

// ---------------------------------------------------------------------------------------------
// Software-pipelined variant ("P"): the FIR of sub-tile n+1 is issued inside the FFT of sub-tile n.
// MODE 0: complex frames; 1: fused FM discriminator (degree-9 atan2); 2: the same with the degree-13 atan2 (audio mode).
//
// Why: in chan256_kernel the FIR phase is a burst of FFMA2 (FMA-pipe bound, issue slots idle) and the FFT
// phase is bound by shared-memory latency (short-scoreboard stalls, FMA pipe idle) — ncu attributes 30 % /
// 46 % of the warp samples to them (profiles/r01_chan_fm_notes.md). Both live in the same 128 threads, so
// placing them in ONE instruction stream lets every warp fill its LDS/exchange latency with independent
// FFMA2 work instead of relying on the 4 resident CTAs being in different phases. The FIR output buffer is
// double-buffered (u[2]); one CTA-wide barrier per sub-tile disappears.
//
//   pre-loop : FIR(0) -> u[0]
//   iter n   : wait stage[n+1] | FFT(n) in u[n&1]  ++  FIR(n+1) -> u[(n+1)&1] | barrier | TMA issue(n+3)
//              | discriminator(n) from u[n&1] | barrier
// ---------------------------------------------------------------------------------------------
struct __align__(128) ChanSmemP {
    u64 stage[2][CH_S * CH_H];       // TMA landing buffers: 8 rows of 128 cf32 samples each
    u64 u[2][CH_S * CH_REGION];      // FIR output -> FFT exchange -> planar FFT output (in place), double buffered
    float2 tw[16 * 16];
    uint64_t full[2];
    uint64_t drained;                // split barrier: all discriminator reads of a sub-tile's u[] are done
};

// per-thread FIR state: sliding window of residue class r (9 history rows + up to 2 new rows), taps of u[r] and u[r+128]
struct FirState {
    u64 w[11];
    float hlo[CH_T], hhi[CH_T];
};

// FIR rows [I0, I1) of one sub-tile: new row from the staged tile, two packed accumulations, window shift
template <int I0, int I1, int FMT>
__device__ __forceinline__ void fir_rows(FirState& f, const u64* __restrict__ st, u64* __restrict__ ub, int r) {
#pragma unroll
    for (int i = I0; i < I1; ++i) {
        f.w[9] = in_pair<FMT>(st, i * CH_H + r);
        u64 lo = mul2(f.w[8], bc2(f.hlo[0])), hi = mul2(f.w[9], bc2(f.hhi[0]));
#pragma unroll
        for (int j = 1; j < CH_T; ++j) {
            lo = fma2(f.w[8 - j], bc2(f.hlo[j]), lo);
            hi = fma2(f.w[9 - j], bc2(f.hhi[j]), hi);
        }
        ub[i * CH_REGION + r] = lo;
        ub[i * CH_REGION + r + CH_H] = hi;
#pragma unroll
        for (int m = 0; m < 9; ++m) f.w[m] = f.w[m + 1];
    }
}

// The same two rows split into load / math / store so that the caller can keep every shared-memory load of a
// segment ahead of every shared-memory store (ptxas cannot prove that the stage, exchange and twiddle regions do
// not alias, so it never moves an LDS above an earlier STS).
template <int I, int FMT>
__device__ __forceinline__ void fir2_load(FirState& f, const u64* __restrict__ st, int r) {
    f.w[9] = in_pair<FMT>(st, I * CH_H + r);
    f.w[10] = in_pair<FMT>(st, (I + 1) * CH_H + r);
}
__device__ __forceinline__ void fir2_math(FirState& f, u64 (&o)[4]) {
    o[0] = mul2(f.w[8], bc2(f.hlo[0]));
    o[1] = mul2(f.w[9], bc2(f.hhi[0]));
    o[2] = mul2(f.w[9], bc2(f.hlo[0]));
    o[3] = mul2(f.w[10], bc2(f.hhi[0]));
#pragma unroll
    for (int j = 1; j < CH_T; ++j) {
        o[0] = fma2(f.w[8 - j], bc2(f.hlo[j]), o[0]);
        o[1] = fma2(f.w[9 - j], bc2(f.hhi[j]), o[1]);
        o[2] = fma2(f.w[9 - j], bc2(f.hlo[j]), o[2]);
        o[3] = fma2(f.w[10 - j], bc2(f.hhi[j]), o[3]);
    }
#pragma unroll
    for (int m = 0; m < 9; ++m) f.w[m] = f.w[m + 2];
}
template <int I>
__device__ __forceinline__ void fir2_store(const u64 (&o)[4], u64* __restrict__ ub, int r) {
    ub[I * CH_REGION + r] = o[0];
    ub[I * CH_REGION + r + CH_H] = o[1];
    ub[(I + 1) * CH_REGION + r] = o[2];
    ub[(I + 1) * CH_REGION + r + CH_H] = o[3];
}

// FFT-256 of the sub-tile in `uc` (16 threads per frame), optionally with the FIR of the next sub-tile
// (stage tile `st` -> `un`) interleaved two rows per segment. FULL: all 8 frames valid (no guards).
// Every segment (between warp barriers) is written loads -> math -> stores.
template <int MODE, bool FULL, bool FIR, int FMT>
__device__ __forceinline__ void fft_fir_phase(const ChanArgs& a, u64* __restrict__ uc, const float2* __restrict__ tws,
                                              FirState& f, const u64* __restrict__ st, u64* __restrict__ un,
                                              int tid, int fs, int nv, int f0, long long out_base,
                                              uint64_t* drained, int drained_parity) {
    const int g = tid >> 4, t = tid & 15;
    u64* reg = uc + g * CH_REGION;
    const bool on = FULL || (g < nv);
    u64 v[16];
    u64 o[4];
    // segment 1: pass-1 operands + FIR rows 0,1
    if (on) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = reg[t + 16 * i];
    }
    if (FIR) {
        fir2_load<0, FMT>(f, st, tid);
        fir2_math(f, o);
        // `un` was read by the discriminator of the sub-tile before last: wait (late) for every thread's reads
        if (drained_parity >= 0) mbar_wait(drained, (uint32_t)drained_parity);
        fir2_store<0>(o, un, tid);
    }
    __syncwarp();
    // segment 2: radix-16 pass 1, four-step twiddle in place, exchange store + FIR rows 2,3
    if (FIR) fir2_load<2, FMT>(f, st, tid);
    if (on) {
        fft16(v);
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) {
            const float2 tw = tws[k1 * 16 + t];
            v[rev4(k1)] = twid(v[rev4(k1)], tw.x, -tw.y);
        }
    }
    if (FIR) fir2_math(f, o);
    if (on) {
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) reg[t * 17 + k1] = v[rev4(k1)];
    }
    if (FIR) fir2_store<2>(o, un, tid);
    __syncwarp();
    // segment 3: pass-2 operands + FIR rows 4,5
    if (on) {
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) v[n2] = reg[n2 * 17 + t];
    }
    if (FIR) {
        fir2_load<4, FMT>(f, st, tid);
        fir2_math(f, o);
        fir2_store<4>(o, un, tid);
    }
    __syncwarp();
    // segment 4: radix-16 pass 2, output + FIR rows 6,7
    if (FIR) fir2_load<6, FMT>(f, st, tid);
    if (on) fft16(v);
    if (FIR) fir2_math(f, o);
    if (on) {
        if (MODE == 0) {
            const int b = fs + g;
            if (b >= f0) {
                u64* og = reinterpret_cast<u64*>(a.out) + (out_base + b) * CH_M + t;
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) og[16 * k2] = v[rev4(k2)];
            }
        } else {
            // bins in natural order, complex, in place over the frame's exchange area (all of its reads precede the warp
            // barrier above): one 64-bit store per bin, and the discriminator fetches two neighbouring bins per 128-bit load
            u64* w = uc + g * CH_REGION + t;
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) w[16 * k2] = v[rev4(k2)];
        }
    }
    if (FIR) fir2_store<6>(o, un, tid);
}

template <int MODE, int FMT>
__global__ void __launch_bounds__(CH_THREADS, 4) chan256p_kernel(const ChanArgs a) {
    extern __shared__ __align__(128) unsigned char chan_smem_raw[];
    ChanSmemP& sm = *reinterpret_cast<ChanSmemP*>(chan_smem_raw);
    const int tid = threadIdx.x;
    const int r = tid;
    const int c = blockIdx.y;
    const int step = (MODE >= 1) ? a.R - 1 : a.R;
    const int f0 = (blockIdx.x == 0) ? 0 : a.R + (blockIdx.x - 1) * step;
    if (f0 >= a.F) return;
    const int f1 = min(a.F, (blockIdx.x == 0) ? a.R : f0 + step);
    const void* __restrict__ xc = in_offset<FMT>(a.x, (long long)c * a.chunk_stride);
    const long long out_base = (long long)c * a.F;

    const int fe = (MODE >= 1 && f0 > 0) ? f0 - 1 : f0;
    const int fast_start = (f0 == 0) ? 8 : fe;
    const int n_fast = (fast_start < f1) ? (f1 - fast_start + CH_S - 1) / CH_S : 0;

    if (tid == 0) {
        mbar_init(&sm.full[0], 1);
        mbar_init(&sm.full[1], 1);
        mbar_init(&sm.drained, CH_THREADS);
        mbar_fence_init();
    }
    for (int i = tid; i < 256; i += CH_THREADS) {
        float s, co;
        sincospif(-(float)((i >> 4) * (i & 15)) * (1.0f / 128.0f), &s, &co);
        sm.tw[i] = make_float2(co, s);
    }
    __syncthreads();

    auto issue = [&](int n) {
        const int fs = fast_start + CH_S * n;
        const int nrows = min(CH_S, a.F - fs);
        const uint32_t bytes = (uint32_t)nrows * CH_H * (FMT == IN_CS16 ? 4u : 8u);
        mbar_expect_tx(&sm.full[n & 1], bytes);
        bulk_g2s(sm.stage[n & 1], in_offset<FMT>(xc, (long long)(fs + 1) * CH_H), bytes, &sm.full[n & 1]);
    };
    if (tid == 0) {
        if (n_fast > 0) issue(0);
        if (n_fast > 1) issue(1);
    }

    FirState f;
#pragma unroll
    for (int j = 0; j < CH_T; ++j) {
        f.hlo[j] = __ldg(a.taps + r + CH_M * j);
        f.hhi[j] = __ldg(a.taps + r + CH_H + CH_M * j);
    }

    u64 p0 = 0ull, p1 = 0ull;   // previous frame's bins 2*tid and 2*tid + 1

    // discriminator of the sub-tile whose FFT output sits in `ub` (frames fs .. fs+nv-1): y_b * conj(y_{b-1}) =
    // Re(p) * y + Im(p) * (-j y), two packed operations per bin
    auto disc = [&](const u64* ub, int fs, int nv) {
        float* o = reinterpret_cast<float*>(a.out) + (out_base + fs) * CH_M + 2 * tid;
        const AtanScaled at = a.at;
        auto one = [&](int i, bool guard) {
            const ulonglong2 yy = *reinterpret_cast<const ulonglong2*>(ub + i * CH_REGION + 2 * tid);
            const u64 z0 = fma2(mulmj(yy.x), bc2(hi2(p0)), mul2(yy.x, bc2(lo2(p0))));
            const u64 z1 = fma2(mulmj(yy.y), bc2(hi2(p1)), mul2(yy.y, bc2(lo2(p1))));
            const u64 px = pk2(lo2(z0), lo2(z1)), py = pk2(hi2(z0), hi2(z1));
            const u64 d = (MODE == 2) ? scaled_atan2f_hi_x2(py, px, at) : scaled_atan2f_x2(py, px, at);
            if (!guard || fs + i >= f0) *reinterpret_cast<u64*>(o + (long long)i * CH_M) = d;
            p0 = yy.x;
            p1 = yy.y;
        };
        if (nv == CH_S && fs >= f0) {
#pragma unroll
            for (int i = 0; i < CH_S; ++i) one(i, false);
        } else if (nv == CH_S) {
#pragma unroll
            for (int i = 0; i < CH_S; ++i) one(i, true);
        } else {
            for (int i = 0; i < nv; ++i) one(i, true);
        }
    };

    // ---- prologue sub-tile: frames 0..7 of a chunk reach into the previous call/chunk ----
    if (f0 == 0) {
        const int nv = min(CH_S, f1);
        for (int b = 0; b < nv; ++b) {
            float2 lo = make_float2(0.f, 0.f), hi = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < CH_T; ++j) {
                const float2 vl = blk_fetch<FMT>(a, xc, c, b - j, r);
                const float2 vh = blk_fetch<FMT>(a, xc, c, b - j, r + CH_H);
                lo.x = fmaf(f.hlo[j], vl.x, lo.x);
                lo.y = fmaf(f.hlo[j], vl.y, lo.y);
                hi.x = fmaf(f.hhi[j], vh.x, hi.x);
                hi.y = fmaf(f.hhi[j], vh.y, hi.y);
            }
            sm.u[0][b * CH_REGION + r] = pk2(lo.x, lo.y);
            sm.u[0][b * CH_REGION + r + CH_H] = pk2(hi.x, hi.y);
        }
        __syncthreads();
        fft_fir_phase<MODE, false, false, FMT>(a, sm.u[0], sm.tw, f, nullptr, nullptr, tid, 0, nv, f0, out_base, nullptr, -1);
        __syncthreads();
        if (MODE >= 1) {
            disc(sm.u[0], 0, nv);
            __syncthreads();
        }
    }
    if (n_fast == 0) return;

#pragma unroll
    for (int m = 0; m < 9; ++m) f.w[m] = in_pair<FMT>(xc, (long long)(fast_start - 8 + m) * CH_H + r);
    // pre-loop: FIR of fast sub-tile 0
    mbar_wait(&sm.full[0], 0);
    fir_rows<0, 8, FMT>(f, sm.stage[0], sm.u[0], r);
    __syncthreads();
    if (tid == 0 && 2 < n_fast) issue(2);

    for (int n = 0; n < n_fast; ++n) {
        const int fs = fast_start + CH_S * n;
        const int nv = min(CH_S, f1 - fs);
        u64* uc = sm.u[n & 1];
        if (n + 1 < n_fast) {
            // nv == 8 here: only the last sub-tile of a run can be ragged
            mbar_wait(&sm.full[(n + 1) & 1], ((n + 1) >> 1) & 1);
            fft_fir_phase<MODE, true, true, FMT>(a, uc, sm.tw, f, sm.stage[(n + 1) & 1], sm.u[(n + 1) & 1], tid, fs, nv, f0,
                                            out_base, &sm.drained, (MODE >= 1 && n > 0) ? ((n - 1) & 1) : -1);
        } else {
            fft_fir_phase<MODE, false, false, FMT>(a, uc, sm.tw, f, nullptr, nullptr, tid, fs, nv, f0, out_base, nullptr, -1);
        }
        __syncthreads();
        if (tid == 0 && n + 3 < n_fast) issue(n + 3);
        if (MODE >= 1) {
            disc(uc, fs, nv);
            // split barrier instead of __syncthreads: u[n&1] is next written by the FIR stores of iteration n+1,
            // which wait on this phase (parity n&1) only after their own loads and math
            mbar_arrive(&sm.drained);
        }
    }
}

// ---- carried-history update: blk_{-8..-1} for the next call (double buffered) ----
template <int FMT>
__global__ void chan_carry_kernel(const void* x_last, int F, int M, int T1, const float2* old_c, float2* new_c) {
    // new_c[m][k], m = 0..T1-1 <-> block index F + m - T1 of the last chunk (or older history); T1 = T
    const int m = blockIdx.x;
    const int H = M / 2;
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
        const int i = F + m - T1;
        float2 v;
        if (i >= 0) {
            const u64 p = in_pair<FMT>(x_last, (long long)i * H + k);
            v = make_float2(lo2(p), hi2(p));
        } else {
            v = old_c[(i + T1) * M + k];
        }
        new_c[m * M + k] = v;
    }
}

// ---- audio mode: nbfm_demod(extract_channel(k), demod_rate, audio_rate) for every channel ----------------------
// (dsp/fm.py:317-406 with its defaults: quadrature_demod -> rms_normalize -> resample_poly -> soft_clip.) The fused FM
// kernel leaves the discriminator rows d[chunk][frame][channel] in a workspace; this kernel is scipy's resample_poly
// for an integer decimation D (up = 1: taps firwin(2*10*D+1, 1/D, kaiser 5.0), y[m] = sum_j h[j] d[D m + 10 D - j],
// zero extension at the chunk ends) along the frame axis of every channel, plus the per-(chunk, channel) sum of
// squares rms_normalize needs. The RMS scale commutes with the (linear) resampler, so it is applied afterwards by
// chan_audio_finish_kernel together with the tanh soft clip. Lanes run along channels (coalesced rows); a warp owns AO
// consecutive outputs: per tap phase p it holds the 21 taps h[p + D k] and walks the inputs d[D q + p], every input
// feeding up to AO accumulators — 2*10*AO*... FMAs for ~(AO + 20) D loads.
constexpr int AR = 32;         // consecutive outputs per warp (run length): accumulators live in registers
constexpr int AU_WARPS = 4;    // warps per CTA
constexpr int AU_K = 21;       // taps per phase: h has 2 * 10 * D + 1 = 20 D + 1 entries
constexpr int AU_PF = 12;       // input rows in flight ahead of their use
struct AudioArgs {
    const float* d;        // [n_chunks][F][M]
    float* audio;          // [n_chunks][n_out][M] (unscaled)
    double* sumsq;         // [n_chunks][M]
    const float* taps;     // [2*10*D + 1]
    int F, M, D, n_out;
};
// Polyphase by component: y[m] = sum_p sum_k T[p][k] x_p[m + 10 - k], x_p[u] = d[D u + p], T[p][k] = h[D k - p] (0 outside
// the filter) — D short FIRs of 21 taps over the decimated sequences x_p. A warp (lanes = 32 channels, coalesced rows) owns a
// run of AR consecutive outputs: per phase it slides a 21-sample register window along the run, so every input row is
// loaded ONCE per run (+20 of lead-in per phase: (AR + 20) / AR = 1.6 loads per output and phase, against 3.5 for the
// 8-output block form this replaces, which was L2-bound), 21 FMAs per load, the AR accumulators and the window in registers
// with every index a compile-time constant (the run loop is fully unrolled, the window rotates by renaming).
// the phases p = warp, warp + AU_WARPS, ... of one run: sliding 21-sample register window per phase, AR accumulators
template <bool EDGE>
__device__ __forceinline__ void audio_phases(const AudioArgs& a, const float* __restrict__ dc, const float* __restrict__ au_taps,
                                             int warp, int m0, float (&acc)[AR], float& ss) {
    const int D = a.D;
    const int step = D * a.M;                         // elements between consecutive rows of one phase (fits 32 bits: F * M < 2^31)
    for (int p = warp; p < D; p += AU_WARPS) {
        float t[AU_K];
#pragma unroll
        for (int k = 0; k < AU_K; ++k) t[k] = au_taps[p * AU_K + k];
        // x_p[u] for u = m0 - 10 .. m0 + AR + 9; window slot of u: (u - (m0 - 10)) % 21. Loads are volatile asm so that ptxas
        // keeps them where they are written — AU_PF steps ahead of their use — instead of hoisting the whole run's loads to
        // the top of the phase; the row index is made opaque per phase, otherwise every row address of the run is computed
        // before the phase loop (loop-invariant in p) and kept live: 36 x 64-bit registers.
        int u_next = m0 - 10;
        asm volatile("" : "+r"(u_next));
        int n_next = D * u_next + p;                  // row index (may be negative in front of the chunk)
        int off = n_next * a.M;
        auto ld = [&]() -> float {
            float v = 0.f;
            if (!EDGE || (n_next >= 0 && n_next < a.F)) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(dc + off));
            off += step;
            if (EDGE) n_next += D;
            return v;
        };
        float w[AU_K];
#pragma unroll
        for (int j = 0; j < AU_K - 1; ++j) w[j] = ld();                  // lead-in: u = m0 - 10 .. m0 + 9
        float pre[AU_PF];
#pragma unroll
        for (int j = 0; j < AU_PF; ++j) pre[j] = ld();
#pragma unroll
        for (int i = 0; i < AR; ++i) {
            // output m = m0 + i needs u = m + 10 - k, k = 0..20: newest u = m0 + i + 10 -> slot (i + 20) % 21
            w[(i + 20) % AU_K] = pre[i % AU_PF];
            if (i + AU_PF < AR) pre[i % AU_PF] = ld();
            // own range of the sum of squares: inputs D (m0 + i) + p, i.e. u = m0 + i, which entered the window 10 steps ago
            const float xo = w[(i + 10) % AU_K];
            ss = fmaf(xo, xo, ss);
#pragma unroll
            for (int k = 0; k < AU_K; ++k) acc[i] = fmaf(t[k], w[(i + 20 - k + AU_K) % AU_K], acc[i]);
        }
    }
}

// The AU_WARPS warps of a CTA share ONE run and split its phases (warp w takes p = w, w + AU_WARPS, ...): together they sweep
// the run's input region once, at the same time, so the rows are fetched from HBM as a near-contiguous stream instead of
// being revisited phase by phase by a lone warp long after they left L2 (measured: 1.7 ms -> 1.1 ms per 32 chunks for the
// register-window form alone, DRAM-pattern-bound; the partial sums of the warps are added in a fixed order through shared
// memory).
__global__ void __launch_bounds__(32 * AU_WARPS, 4) chan_audio_kernel(const AudioArgs a) {
    extern __shared__ float au_smem[];   // [D][AU_K] taps: au_taps[p * AU_K + k] = h[D k - p] | [AU_WARPS][AR + 1][32] partial sums
    float* au_taps = au_smem;
    const int D = a.D, half = 10 * D;
    float* red = au_smem + ((D * AU_K + 31) & ~31);
    for (int i = threadIdx.x; i < D * AU_K; i += blockDim.x) {
        const int p = i / AU_K, k = i % AU_K, j = D * k - p;
        au_taps[i] = (j >= 0 && j <= 2 * half) ? a.taps[j] : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = blockIdx.x * 32 + lane;
    const int c = blockIdx.z;
    const int m0 = blockIdx.y * AR;
    const bool live = ch < a.M;
    const float* dc = a.d + (long long)c * a.F * a.M + (live ? ch : 0);
    float acc[AR];
#pragma unroll
    for (int i = 0; i < AR; ++i) acc[i] = 0.f;
    float ss = 0.f;
    // interior runs (all but the first and the last one or two of a chunk) never touch the zero extension: their loads carry
    // no bounds test, one 32-bit offset add and one address formation each (the tested form costs ~10 instructions per load)
    const bool interior = (m0 >= 10) && ((long long)D * (m0 + AR + 9) + (D - 1) < a.F);
    if (interior) audio_phases<false>(a, dc, au_taps, warp, m0, acc, ss);
    else audio_phases<true>(a, dc, au_taps, warp, m0, acc, ss);
    // partial sums of the warps -> warp 0 adds them in warp order and stores
    float* mine = red + warp * (AR + 1) * 32;
#pragma unroll
    for (int i = 0; i < AR; ++i) mine[i * 32 + lane] = acc[i];
    mine[AR * 32 + lane] = ss;
    __syncthreads();
    if (warp != 0 || !live) return;
    float* o = a.audio + ((long long)c * a.n_out + m0) * a.M + ch;
#pragma unroll 4
    for (int i = 0; i < AR; ++i) {
        float v = red[i * 32 + lane];
        for (int w2 = 1; w2 < AU_WARPS; ++w2) v += red[w2 * (AR + 1) * 32 + i * 32 + lane];
        if (m0 + i < a.n_out) o[(long long)i * a.M] = v;
    }
    float sv = red[AR * 32 + lane];
    for (int w2 = 1; w2 < AU_WARPS; ++w2) sv += red[w2 * (AR + 1) * 32 + AR * 32 + lane];
    atomicAdd(a.sumsq + (long long)c * a.M + ch, (double)sv);
}

// rms_normalize (dsp/fm.py:42-62: rms = sqrt(mean(x^2)) in float32, scale = target/rms when rms > min_rms) applied after the
// resampler, then dsp.fm.soft_clip (:26-39)
__global__ void chan_audio_finish_kernel(float* __restrict__ audio, const double* __restrict__ sumsq, int F, int M, int n_out,
                                         long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % M);
        const long long c = i / ((long long)n_out * M);
        const float rms = sqrtf((float)(sumsq[c * M + ch] / (double)F));
        const float scale = (rms > 1e-4f) ? (float)(0.18 / (double)rms) : 1.0f;
        audio[i] = tanhf(audio[i] * scale * 1.5f) * 1.1047914f * 0.95f;
    }
}

// ---- generic path (any even M): FIR then direct DFT ----
struct GenArgs {
    const float2* x;
    long long chunk_stride;
    int F, M, T;
    const float* taps;      // [M*T]
    const float2* carried;  // [T][M] blk_{-T..-1}
    float2* u;              // [n_chunks*F][M] workspace
};

__global__ void chan_generic_fir_kernel(const GenArgs a) {
    const int c = blockIdx.z;
    const int b = blockIdx.x;
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    if (k >= a.M) return;
    const int H = a.M / 2, T1 = a.T;
    const float2* xc = a.x + (long long)c * a.chunk_stride;
    float2 acc = make_float2(0.f, 0.f);
    for (int j = 0; j < a.T; ++j) {
        const int i = b - j;
        float2 v;
        if (i >= 0) v = xc[(long long)i * H + k];
        else if (c > 0) v = (xc - a.chunk_stride)[(long long)(a.F + i) * H + k];
        else v = a.carried[(i + T1) * a.M + k];
        const float h = a.taps[k + j * a.M];
        acc.x = fmaf(h, v.x, acc.x);
        acc.y = fmaf(h, v.y, acc.y);
    }
    a.u[((long long)c * a.F + b) * a.M + k] = acc;
}

// one CTA per frame; smem = u[M] + twiddle[M]
__global__ void chan_generic_dft_kernel(const float2* __restrict__ u, float2* __restrict__ y, int M) {
    extern __shared__ float2 gsm[];
    float2* su = gsm;
    float2* sw = gsm + M;
    const long long fr = blockIdx.x;
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
        su[k] = u[fr * M + k];
        double s, co;
        sincospi(-2.0 * (double)k / (double)M, &s, &co);
        sw[k] = make_float2((float)co, (float)s);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < M; q += blockDim.x) {
        float2 acc = make_float2(0.f, 0.f);
        int idx = 0;
        for (int k = 0; k < M; ++k) {
            const float2 p = cmul(su[k], sw[idx]);
            acc.x += p.x;
            acc.y += p.y;
            idx += q;
            if (idx >= M) idx -= M;
        }
        y[fr * M + q] = acc;
    }
}

// discriminator over frames for the generic path: y [n_chunks][F][M] -> d
__global__ void chan_generic_disc_kernel(const float2* __restrict__ y, float* __restrict__ d, int F, int M, float scale) {
    const int c = blockIdx.z;
    const int b = blockIdx.x;
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    if (k >= M) return;
    const long long o = ((long long)c * F + b) * M + k;
    if (b == 0) {
        d[o] = 0.f;
        return;
    }
    const float2 p = cmulc(y[o], y[o - M]);
    d[o] = fast_atan2f(p.y, p.x) * scale;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

static double bessel_i0(double x) {
    // power series; converges quickly for the beta values used by Kaiser windows
    double sum = 1.0, term = 1.0;
    const double q = x * x / 4.0;
    for (int k = 1; k < 200; ++k) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < 1e-18 * sum) break;
    }
    return sum;
}

// scipy.signal.firwin(numtaps, cutoff, window=("kaiser", beta)) for a single low-pass band,
// cutoff normalised to Nyquist, scaled to unit DC gain.
void firwin_kaiser_lowpass(int numtaps, double cutoff, double beta, std::vector<double>& h) {
    h.resize(numtaps);
    const double alpha = 0.5 * (numtaps - 1);
    const double i0b = bessel_i0(beta);
    double s = 0.0;
    for (int n = 0; n < numtaps; ++n) {
        const double m = n - alpha;
        const double xx = cutoff * m;
        const double sinc = (xx == 0.0) ? 1.0 : sin(M_PI * xx) / (M_PI * xx);
        double rr = (numtaps > 1) ? (2.0 * n / (numtaps - 1) - 1.0) : 0.0;
        double arg = 1.0 - rr * rr;
        if (arg < 0) arg = 0;
        const double win = bessel_i0(beta * sqrt(arg)) / i0b;
        h[n] = cutoff * sinc * win;
        s += h[n];
    }
    for (int n = 0; n < numtaps; ++n) h[n] /= s;
}

}  // namespace wc

using namespace wc;

struct wc_chan {
    double sample_rate;
    int channel_bandwidth;
    int T;
    int M;
    std::vector<double> arms;  // [M][T] float64, as the reference exposes
    float* d_taps = nullptr;   // [T][M] float32: h[k + j*M]
    float2* d_carried[2] = {nullptr, nullptr};
    int cur = 0;
    int run_frames = 0;        // wc_chan_set_run_frames: frames per CTA run, 0 = sized from the grid target
    // audio mode (wc_chan_audio_config): nbfm_demod(extract_channel(k), demod_rate, audio_rate), integer decimation D
    int audio_D = 0, audio_demod_rate = 0;
    float* d_audio_taps = nullptr;                     // [2*10*D + 1] firwin(.., 1/D, kaiser 5.0)
    float* d_disc = nullptr;  size_t disc_bytes = 0;   // discriminator rows [n_chunks][F][M]
    double* d_sumsq = nullptr; size_t sumsq_bytes = 0; // [n_chunks][M]
    // workspaces
    float2* d_ws = nullptr;   size_t ws_bytes = 0;     // generic path u / y
    void* d_in = nullptr;     size_t in_bytes = 0;     // host API staging
    void* d_out = nullptr;    size_t out_bytes = 0;
    cudaStream_t stream = nullptr;                     // compute stream of the *_host entry points
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;     // copy streams of the pipelined host path
    cudaEvent_t ev_in[2] = {}, ev_comp[2] = {}, ev_out[2] = {};
};

static int ensure(void** p, size_t* cap, size_t need) {
    if (*cap >= need) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    WC_CUDA(cudaMalloc(p, need));
    *cap = need;
    return 0;
}

extern "C" {

int wc_chan_create(double sample_rate, int channel_bandwidth, int taps_per_channel, wc_chan** out) {
    WC_REQUIRE(out != nullptr, "wc_chan_create: out is null");
    WC_REQUIRE(sample_rate > 0 && channel_bandwidth > 0 && taps_per_channel >= 1 && taps_per_channel <= 64,
               "wc_chan_create: bad parameters");
    int M = (int)(sample_rate / (double)channel_bandwidth);  // channelizer.py:53-55
    if (M % 2 != 0) M -= 1;
    WC_REQUIRE(M >= 2 && M <= 8192, "wc_chan_create: channel count %d out of range [2, 8192]", M);
    wc_chan* h = new wc_chan();
    h->sample_rate = sample_rate;
    h->channel_bandwidth = channel_bandwidth;
    h->T = taps_per_channel;
    h->M = M;
    // channelizer.py:69-89 prototype + polyphase split
    std::vector<double> proto;
    const int L = M * taps_per_channel - 1;
    const double cutoff = ((double)channel_bandwidth * 0.9) / (sample_rate / 2.0);
    firwin_kaiser_lowpass(L, cutoff, 8.0, proto);
    h->arms.assign((size_t)M * taps_per_channel, 0.0);
    std::vector<float> taps((size_t)M * taps_per_channel, 0.f);
    for (int k = 0; k < M; ++k)
        for (int j = 0; j < taps_per_channel; ++j) {
            const int idx = k + j * M;
            const double v = idx < L ? proto[idx] : 0.0;
            h->arms[(size_t)k * taps_per_channel + j] = v;
            taps[(size_t)j * M + k] = (float)v;
        }
    const size_t carry_bytes = sizeof(float2) * (size_t)M * (size_t)taps_per_channel;
    if (cudaMalloc(&h->d_taps, taps.size() * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&h->d_carried[0], carry_bytes) != cudaSuccess ||
        cudaMalloc(&h->d_carried[1], carry_bytes) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("wc_chan_create: CUDA allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete h;
        return -2;
    }
    cudaMemcpy(h->d_taps, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemset(h->d_carried[0], 0, carry_bytes);
    cudaMemset(h->d_carried[1], 0, carry_bytes);
    *out = h;
    return 0;
}

void wc_chan_destroy(wc_chan* h) {
    if (!h) return;
    cudaFree(h->d_taps);
    cudaFree(h->d_carried[0]);
    cudaFree(h->d_carried[1]);
    if (h->d_audio_taps) cudaFree(h->d_audio_taps);
    if (h->d_disc) cudaFree(h->d_disc);
    if (h->d_sumsq) cudaFree(h->d_sumsq);
    if (h->d_ws) cudaFree(h->d_ws);
    if (h->d_in) cudaFree(h->d_in);
    if (h->d_out) cudaFree(h->d_out);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->s_h2d) {
        cudaStreamDestroy(h->s_h2d);
        cudaStreamDestroy(h->s_d2h);
        for (int i = 0; i < 2; ++i) {
            cudaEventDestroy(h->ev_in[i]);
            cudaEventDestroy(h->ev_comp[i]);
            cudaEventDestroy(h->ev_out[i]);
        }
    }
    delete h;
}

int wc_chan_info(const wc_chan* h, int* channel_count, double* channel_sample_rate, int* taps_per_channel) {
    WC_REQUIRE(h != nullptr, "wc_chan_info: null handle");
    if (channel_count) *channel_count = h->M;
    if (channel_sample_rate) *channel_sample_rate = (h->sample_rate / h->M) * 2.0;  // channelizer.py:58
    if (taps_per_channel) *taps_per_channel = h->T;
    return 0;
}

int wc_chan_get_arms(const wc_chan* h, double* arms) {
    WC_REQUIRE(h && arms, "wc_chan_get_arms: null argument");
    for (size_t i = 0; i < h->arms.size(); ++i) arms[i] = h->arms[i];
    return 0;
}

long long wc_chan_frames_for(const wc_chan* h, long long n_samples) {
    if (!h || n_samples < h->M) return 0;
    return (n_samples - h->M) / (h->M / 2) + 1;  // range(0, len - M + 1, M/2), channelizer.py:114
}

int wc_chan_reset(wc_chan* h) {
    WC_REQUIRE(h != nullptr, "wc_chan_reset: null handle");
    const size_t carry_bytes = sizeof(float2) * (size_t)h->M * (size_t)h->T;
    WC_CUDA(cudaMemsetAsync(h->d_carried[0], 0, carry_bytes, h->stream));
    WC_CUDA(cudaMemsetAsync(h->d_carried[1], 0, carry_bytes, h->stream));
    WC_CUDA(cudaStreamSynchronize(h->stream));
    h->cur = 0;
    return 0;
}

// arm_history[k][j] = blk_{last-j}[k] (channelizer.py:64,122-126), complex64 [M][T].
int wc_chan_get_history(wc_chan* h, void* arm_history_host) {
    WC_REQUIRE(h && arm_history_host, "wc_chan_get_history: null argument");
    const int T = h->T;
    std::vector<float2> c((size_t)h->M * T);
    WC_CUDA(cudaStreamSynchronize(h->stream));
    WC_CUDA(cudaMemcpy(c.data(), h->d_carried[h->cur], c.size() * sizeof(float2), cudaMemcpyDeviceToHost));
    float2* o = reinterpret_cast<float2*>(arm_history_host);
    for (int k = 0; k < h->M; ++k)
        for (int j = 0; j < T; ++j) o[(size_t)k * T + j] = c[(size_t)(T - 1 - j) * h->M + k];
    return 0;
}

static int chan_audio_stage(wc_chan* h, long long F, int n_chunks, float* audio_out, cudaStream_t stream);

int wc_chan_process_ex(wc_chan* h, const void* iq_dev, int in_fmt, long long n_samples, int n_chunks, long long chunk_stride,
                       int mode, float fm_scale, void* out_dev, void* stream_v) {
    WC_REQUIRE(h && iq_dev && out_dev, "wc_chan_process: null argument");
    WC_REQUIRE(mode == WC_CHAN_OUT_COMPLEX || mode == WC_CHAN_OUT_FM || mode == WC_CHAN_OUT_AUDIO, "wc_chan_process: bad mode %d", mode);
    WC_REQUIRE(in_fmt == WC_CHAN_IN_CF32 || in_fmt == WC_CHAN_IN_CS16, "wc_chan_process: bad input format %d", in_fmt);
    WC_REQUIRE(n_chunks >= 1, "wc_chan_process: n_chunks must be >= 1");
    void* audio_out = nullptr;
    if (mode == WC_CHAN_OUT_AUDIO) {
        // discriminator rows go to the handle's workspace, the decimating audio stage writes the caller's buffer
        WC_REQUIRE(h->audio_D > 0, "wc_chan_process: audio mode needs wc_chan_audio_config first");
        const long long Fa = wc_chan_frames_for(h, n_samples);
        if (Fa == 0) return 0;
        if (ensure((void**)&h->d_disc, &h->disc_bytes, sizeof(float) * (size_t)Fa * n_chunks * h->M)) return -2;
        audio_out = out_dev;
        out_dev = h->d_disc;
        mode = WC_CHAN_OUT_FM;
        fm_scale = (float)(h->audio_demod_rate / (2.0 * M_PI * 75000.0));   // dsp/fm.py:94
    }
    cudaStream_t stream = (cudaStream_t)stream_v;  // taken literally: NULL = the CUDA default stream
    const long long F = wc_chan_frames_for(h, n_samples);
    if (F == 0) return 0;
    WC_REQUIRE(F < (1LL << 30), "wc_chan_process: chunk too long");
    const int T1 = h->T;
    WC_REQUIRE(n_chunks == 1 || F >= T1, "wc_chan_process: batched chunks need >= %d frames each", T1);
    WC_REQUIRE(n_chunks == 1 || chunk_stride >= n_samples, "wc_chan_process: chunk_stride < n_samples");
    const float2* x = reinterpret_cast<const float2*>(iq_dev);
    const int align_samples = (in_fmt == WC_CHAN_IN_CS16) ? 4 : 2;   // bulk async copies need 16-byte aligned row bases

    const bool fast = (h->M == CH_M && h->T == CH_T && ((uintptr_t)iq_dev % 16 == 0) &&
                       (n_chunks == 1 || chunk_stride % align_samples == 0));
    WC_REQUIRE(fast || in_fmt == WC_CHAN_IN_CF32,
               "wc_chan_process: int16 input is built for the 256-channel / 9-tap grid with 16-byte aligned chunks only");
    if (fast) {
        ChanArgs a;
        a.x = iq_dev;
        a.chunk_stride = chunk_stride;
        a.F = (int)F;
        a.taps = h->d_taps;
        a.carried = h->d_carried[h->cur];
        a.out = out_dev;
        a.scale = fm_scale;
        int R;
        int occ = 4;  // resident CTAs per SM of the fused-FM kernel; measured on B200: 4 -> 181 GS/s, 5 -> 165, 6 -> 159 (profiles/r01_chan_sweep2.jsonl)
        // dev switches (dev builds only; env_int is the inlined default in the product library)
        const int occ_env = env_int("WC_CHAN_OCC", 0), r_env = env_int("WC_CHAN_R", 0);
        if (occ_env > 0) occ = occ_env;
        if (mode == WC_CHAN_OUT_COMPLEX) occ = 4;
        if (h->run_frames > 0) {
            R = h->run_frames;
        } else if (r_env > 0) {
            R = r_env;
        } else {
            const long long total = F * n_chunks;
            const long long target = (long long)sm_count() * occ * 6;
            R = (int)((total + target - 1) / target);
        }
        R = ((R + 7) / 8) * 8;
        if (R < 16) R = 16;
        if (R > 256) R = 256;   // measured: R = 256 is the optimum (128: -1 %, 512: -1 %, 2048: -8 %)
        a.R = R;
        const int step = (mode == WC_CHAN_OUT_FM) ? R - 1 : R;
        const unsigned runs = (F > R) ? 1u + (unsigned)((F - R + step - 1) / step) : 1u;
        dim3 grid(runs, (unsigned)n_chunks);
        {
            // scale * (degree-9 minimax coefficients of atan(t)/t in t^2), tools/fit_atan.py
            const double c[5] = {0.999970019, -0.331700921, 0.185215309, -0.0919253752, 0.0238626394};
            a.at.k0 = (float)(c[0] * fm_scale);
            a.at.k1 = (float)(c[1] * fm_scale);
            a.at.k2 = (float)(c[2] * fm_scale);
            a.at.k3 = (float)(c[3] * fm_scale);
            a.at.k4 = (float)(c[4] * fm_scale);
            a.at.hp = (float)(1.5707963267948966 * fm_scale);
            a.at.pi = (float)(3.141592653589793 * fm_scale);
            const double ch[7] = {0.9999994039535522, -0.3332701623439789, 0.198873370885849, -0.13512229919433594,
                                  0.08435501158237457, -0.037443727254867554, 0.008007131516933441};
            for (int i = 0; i < 7; ++i) a.at.h[i] = (float)(ch[i] * fm_scale);
        }
#ifdef WC_DEV
        if (env_int("WC_CHAN_VAR", 1) == 0) {   // phase-serial kernel
#ifdef WC_DEV_ABLATE
            const int abl_env = env_int("WC_CHAN_ABL", 0);
#endif
            if (mode == WC_CHAN_OUT_COMPLEX) chan256_kernel<0, 4><<<grid, CH_THREADS, 0, stream>>>(a);
#ifdef WC_DEV_ABLATE
            else if (abl_env == 1) chan256_kernel<1, 4, 1><<<grid, CH_THREADS, 0, stream>>>(a);
            else if (abl_env == 2) chan256_kernel<1, 4, 2><<<grid, CH_THREADS, 0, stream>>>(a);
            else if (abl_env == 3) chan256_kernel<1, 4, 3><<<grid, CH_THREADS, 0, stream>>>(a);
            else if (abl_env == 4) chan256_kernel<1, 4, 4><<<grid, CH_THREADS, 0, stream>>>(a);
#endif
            else if (occ == 4) chan256_kernel<1, 4><<<grid, CH_THREADS, 0, stream>>>(a);
            else if (occ == 6) chan256_kernel<1, 6><<<grid, CH_THREADS, 0, stream>>>(a);
            else chan256_kernel<1, 5><<<grid, CH_THREADS, 0, stream>>>(a);
        } else
#endif
        {
            // software-pipelined kernel (FIR of sub-tile n+1 inside the FFT of sub-tile n). The dynamic shared-memory
            // opt-in is a per-(function, device) attribute: set once per device this process launches on.
            static std::atomic<unsigned long long> done[6] = {};
#define WC_CHAN_LAUNCH(MODE, FMT, SLOT)                                                                    \
    do {                                                                                                   \
        WC_CUDA(smem_optin(chan256p_kernel<MODE, FMT>, (int)sizeof(ChanSmemP), done[SLOT]));               \
        chan256p_kernel<MODE, FMT><<<grid, CH_THREADS, sizeof(ChanSmemP), stream>>>(a);                    \
    } while (0)
            if (mode == WC_CHAN_OUT_COMPLEX) {
                if (in_fmt == WC_CHAN_IN_CS16) WC_CHAN_LAUNCH(0, IN_CS16, 0);
                else WC_CHAN_LAUNCH(0, IN_CF32, 1);
            } else if (audio_out) {
                if (in_fmt == WC_CHAN_IN_CS16) WC_CHAN_LAUNCH(2, IN_CS16, 4);
                else WC_CHAN_LAUNCH(2, IN_CF32, 5);
            } else {
                if (in_fmt == WC_CHAN_IN_CS16) WC_CHAN_LAUNCH(1, IN_CS16, 2);
                else WC_CHAN_LAUNCH(1, IN_CF32, 3);
            }
#undef WC_CHAN_LAUNCH
        }
        WC_CUDA(cudaGetLastError());
    } else {
        const size_t need = sizeof(float2) * (size_t)F * n_chunks * h->M * (mode == WC_CHAN_OUT_FM ? 2 : 1);
        if (ensure((void**)&h->d_ws, &h->ws_bytes, need)) return -2;
        GenArgs g;
        g.x = x;
        g.chunk_stride = chunk_stride;
        g.F = (int)F;
        g.M = h->M;
        g.T = h->T;
        g.taps = h->d_taps;
        g.carried = h->d_carried[h->cur];
        g.u = h->d_ws;
        dim3 fgrid((unsigned)F, (h->M + 127) / 128, (unsigned)n_chunks);
        chan_generic_fir_kernel<<<fgrid, 128, 0, stream>>>(g);
        float2* y = (mode == WC_CHAN_OUT_FM) ? h->d_ws + (size_t)F * n_chunks * h->M
                                             : reinterpret_cast<float2*>(out_dev);
        const size_t smem = sizeof(float2) * 2 * h->M;
        if (smem > 48 * 1024)
            WC_CUDA(cudaFuncSetAttribute(chan_generic_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        chan_generic_dft_kernel<<<(unsigned)(F * n_chunks), 256, smem, stream>>>(h->d_ws, y, h->M);
        if (mode == WC_CHAN_OUT_FM)
            chan_generic_disc_kernel<<<fgrid, 128, 0, stream>>>(y, reinterpret_cast<float*>(out_dev), (int)F, h->M, fm_scale);
        WC_CUDA(cudaGetLastError());
    }
    {
        if (in_fmt == WC_CHAN_IN_CS16) {
            const void* x_last = reinterpret_cast<const uint32_t*>(iq_dev) + (long long)(n_chunks - 1) * chunk_stride;
            chan_carry_kernel<IN_CS16><<<T1, 256, 0, stream>>>(x_last, (int)F, h->M, T1, h->d_carried[h->cur], h->d_carried[h->cur ^ 1]);
        } else {
            const float2* x_last = x + (long long)(n_chunks - 1) * chunk_stride;
            chan_carry_kernel<IN_CF32><<<T1, 256, 0, stream>>>(x_last, (int)F, h->M, T1, h->d_carried[h->cur], h->d_carried[h->cur ^ 1]);
        }
        WC_CUDA(cudaGetLastError());
        h->cur ^= 1;
    }
    if (audio_out) return chan_audio_stage(h, F, n_chunks, reinterpret_cast<float*>(audio_out), stream);
    return 0;
}

int wc_chan_process(wc_chan* h, const void* iq_dev, long long n_samples, int n_chunks, long long chunk_stride,
                    int mode, float fm_scale, void* out_dev, void* stream_v) {
    return wc_chan_process_ex(h, iq_dev, WC_CHAN_IN_CF32, n_samples, n_chunks, chunk_stride, mode, fm_scale, out_dev, stream_v);
}

// nbfm_demod(ch_iq, demod_rate, audio_rate) with the reference's defaults (dsp/fm.py:317-406) for every channel:
// resample_poly reduces audio_rate/demod_rate by their gcd; built for up == 1 (integer decimation D = down).
int wc_chan_audio_config(wc_chan* h, int demod_rate, int audio_rate) {
    WC_REQUIRE(h && demod_rate > 0 && audio_rate > 0, "wc_chan_audio_config: bad argument");
    long long a = demod_rate, b = audio_rate;
    while (b) {
        const long long t = a % b;
        a = b;
        b = t;
    }
    const long long up = audio_rate / a, down = demod_rate / a;
    WC_REQUIRE(up == 1 && down >= 2 && down <= 64,
               "wc_chan_audio_config: only integer decimation 2..64 is built (got up/down = %lld/%lld)", up, down);
    const int D = (int)down, ntaps = 2 * 10 * D + 1;
    std::vector<double> hd;
    firwin_kaiser_lowpass(ntaps, 1.0 / D, 5.0, hd);      // scipy resample_poly's design (x up = 1)
    std::vector<float> hf(ntaps);
    for (int i = 0; i < ntaps; ++i) hf[i] = (float)hd[i];
    if (h->d_audio_taps) cudaFree(h->d_audio_taps);
    h->d_audio_taps = nullptr;
    WC_CUDA(cudaMalloc(&h->d_audio_taps, sizeof(float) * ntaps));
    WC_CUDA(cudaMemcpy(h->d_audio_taps, hf.data(), sizeof(float) * ntaps, cudaMemcpyHostToDevice));
    h->audio_D = D;
    h->audio_demod_rate = demod_rate;
    return 0;
}

// audio samples per channel per chunk: ceil(frames / D) (scipy resample_poly output length for up = 1)
long long wc_chan_audio_len(const wc_chan* h, long long n_samples) {
    if (!h || h->audio_D <= 0) return 0;
    const long long F = wc_chan_frames_for(h, n_samples);
    return (F + h->audio_D - 1) / h->audio_D;
}

static int chan_audio_stage(wc_chan* h, long long F, int n_chunks, float* audio_out, cudaStream_t stream) {
    const int D = h->audio_D, M = h->M;
    const long long n_out = (F + D - 1) / D;
    if (ensure((void**)&h->d_sumsq, &h->sumsq_bytes, sizeof(double) * (size_t)n_chunks * M)) return -2;
    WC_CUDA(cudaMemsetAsync(h->d_sumsq, 0, sizeof(double) * (size_t)n_chunks * M, stream));
    AudioArgs a;
    a.d = h->d_disc;
    a.audio = audio_out;
    a.sumsq = h->d_sumsq;
    a.taps = h->d_audio_taps;
    a.F = (int)F;
    a.M = M;
    a.D = D;
    a.n_out = (int)n_out;
    dim3 grid((M + 31) / 32, (unsigned)((n_out + AR - 1) / AR), (unsigned)n_chunks);
    const size_t au_smem = sizeof(float) * (((size_t)D * AU_K + 31) / 32 * 32 + (size_t)AU_WARPS * (AR + 1) * 32);
    chan_audio_kernel<<<grid, 32 * AU_WARPS, au_smem, stream>>>(a);
    const long long total = (long long)n_chunks * n_out * M;
    long long blocks = (total + 255) / 256;
    if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
    chan_audio_finish_kernel<<<(unsigned)blocks, 256, 0, stream>>>(audio_out, h->d_sumsq, (int)F, M, (int)n_out, total);
    WC_CUDA(cudaGetLastError());
    return 0;
}


// Advance the carried history as if process() had just been called on these n_samples, without computing
// any output: used by time-sharded runs where another rank emitted the tail of the call.
int wc_chan_set_run_frames(wc_chan* h, int run_frames) {
    WC_REQUIRE(h && run_frames >= 0, "wc_chan_set_run_frames: bad argument");
    h->run_frames = run_frames;
    return 0;
}

// Set the carried history from the last (T + 1) hop rows of the previous process() call — samples
// [(F - T) * M/2, (F + 1) * M/2) of that call, F its frame count — wherever they now live (e.g. copied out of another
// GPU's slab of a striped capture): identical to what wc_chan_carry_from(prev_call) would leave behind when F >= T.
int wc_chan_carry_tail(wc_chan* h, const void* tail_dev, void* stream_v) {
    WC_REQUIRE(h && tail_dev, "wc_chan_carry_tail: null argument");
    // chan_carry_kernel with F = T reads block i = m (m = 0 .. T-1) at x_last[i * M/2 + k], k < M: exactly the tail layout
    chan_carry_kernel<IN_CF32><<<h->T, 256, 0, (cudaStream_t)stream_v>>>(tail_dev, h->T, h->M, h->T, h->d_carried[h->cur],
                                                                        h->d_carried[h->cur ^ 1]);
    WC_CUDA(cudaGetLastError());
    h->cur ^= 1;
    return 0;
}

int wc_chan_carry_from(wc_chan* h, const void* iq_dev, long long n_samples, void* stream_v) {
    WC_REQUIRE(h && iq_dev, "wc_chan_carry_from: null argument");
    const long long F = wc_chan_frames_for(h, n_samples);
    if (F == 0) return 0;
    WC_REQUIRE(F < (1LL << 30), "wc_chan_carry_from: chunk too long");
    chan_carry_kernel<IN_CF32><<<h->T, 256, 0, (cudaStream_t)stream_v>>>(iq_dev, (int)F, h->M, h->T,
                                                                        h->d_carried[h->cur], h->d_carried[h->cur ^ 1]);
    WC_CUDA(cudaGetLastError());
    h->cur ^= 1;
    return 0;
}

int wc_chan_process_host_ex(wc_chan* h, const void* iq_host, int in_fmt, long long n_samples, int n_chunks, int mode,
                            float fm_scale, void* out_host) {
    WC_REQUIRE(h && iq_host && out_host, "wc_chan_process_host: null argument");
    WC_REQUIRE(n_chunks >= 1, "wc_chan_process_host: n_chunks must be >= 1");
    WC_REQUIRE(in_fmt == WC_CHAN_IN_CF32 || in_fmt == WC_CHAN_IN_CS16, "wc_chan_process_host: bad input format %d", in_fmt);
    const long long F = wc_chan_frames_for(h, n_samples);
    if (F == 0) return 0;
    WC_REQUIRE(n_chunks == 1 || F >= h->T, "wc_chan_process_host: batched chunks need >= %d frames each", h->T);
    // Software pipeline over sub-batches of g chunks: H2D (copy stream) | kernels (compute stream) |
    // D2H (copy stream), double-buffered on the device. Chunk bases stay 16-byte aligned.
    const size_t isz = (in_fmt == WC_CHAN_IN_CS16) ? 4 : 8;
    const long long al = (in_fmt == WC_CHAN_IN_CS16) ? 3 : 1;
    const long long stride = (n_samples + al) & ~al;
    int g = (int)((4LL << 20) / n_samples);
    if (g < 1) g = 1;
    if (g > n_chunks) g = n_chunks;
    // rows of the result per chunk and bytes per row element: frames x complex64 / float32, or audio samples x float32
    const long long rows = (mode == WC_CHAN_OUT_AUDIO) ? wc_chan_audio_len(h, n_samples) : F;
    WC_REQUIRE(mode != WC_CHAN_OUT_AUDIO || h->audio_D > 0, "wc_chan_process_host: audio mode needs wc_chan_audio_config first");
    const size_t esz = (mode == WC_CHAN_OUT_COMPLEX) ? sizeof(float2) : sizeof(float);
    const size_t in_sub = isz * (size_t)stride * g;
    const size_t out_sub = esz * (size_t)rows * g * h->M;
    if (ensure(&h->d_in, &h->in_bytes, 2 * in_sub)) return -2;
    if (ensure(&h->d_out, &h->out_bytes, 2 * out_sub)) return -2;
    if (!h->s_h2d) {
        WC_CUDA(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
        WC_CUDA(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            WC_CUDA(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
            WC_CUDA(cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming));
            WC_CUDA(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
        }
    }
    const char* src = reinterpret_cast<const char*>(iq_host);
    char* dst = reinterpret_cast<char*>(out_host);
    int it = 0;
    for (int c0 = 0; c0 < n_chunks; c0 += g, ++it) {
        const int gc = (n_chunks - c0 < g) ? (n_chunks - c0) : g;
        const int b = it & 1;
        char* din = reinterpret_cast<char*>(h->d_in) + (size_t)b * in_sub;
        char* dout = reinterpret_cast<char*>(h->d_out) + (size_t)b * out_sub;
        if (it >= 2) WC_CUDA(cudaStreamWaitEvent(h->s_h2d, h->ev_comp[b], 0));  // d_in[b] free again
        const char* hs = src + isz * (size_t)n_samples * c0;
        if (stride == n_samples) {
            WC_CUDA(cudaMemcpyAsync(din, hs, isz * (size_t)n_samples * gc, cudaMemcpyHostToDevice, h->s_h2d));
        } else {
            WC_CUDA(cudaMemcpy2DAsync(din, isz * stride, hs, isz * n_samples, isz * n_samples, gc, cudaMemcpyHostToDevice,
                                      h->s_h2d));
        }
        WC_CUDA(cudaEventRecord(h->ev_in[b], h->s_h2d));
        WC_CUDA(cudaStreamWaitEvent(h->stream, h->ev_in[b], 0));
        if (it >= 2) WC_CUDA(cudaStreamWaitEvent(h->stream, h->ev_out[b], 0));  // d_out[b] drained
        int rc = wc_chan_process_ex(h, din, in_fmt, n_samples, gc, stride, mode, fm_scale, dout, h->stream);
        if (rc) return rc;
        WC_CUDA(cudaEventRecord(h->ev_comp[b], h->stream));
        WC_CUDA(cudaStreamWaitEvent(h->s_d2h, h->ev_comp[b], 0));
        const size_t obytes = esz * (size_t)rows * gc * h->M;
        WC_CUDA(cudaMemcpyAsync(dst + esz * (size_t)rows * c0 * h->M, dout, obytes, cudaMemcpyDeviceToHost, h->s_d2h));
        WC_CUDA(cudaEventRecord(h->ev_out[b], h->s_d2h));
    }
    WC_CUDA(cudaStreamSynchronize(h->s_d2h));
    WC_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int wc_chan_process_host(wc_chan* h, const void* iq_host, long long n_samples, int n_chunks, int mode,
                         float fm_scale, void* out_host) {
    return wc_chan_process_host_ex(h, iq_host, WC_CHAN_IN_CF32, n_samples, n_chunks, mode, fm_scale, out_host);
}

}  // extern "C"
