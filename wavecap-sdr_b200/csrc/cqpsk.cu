// P25 CQPSK / LSM (pi/4-DQPSK) symbol recovery, batched over channels (config C4, second half).
//
// Replaces wavecapsdr/decoders/p25.py:190-669 (CQPSKDemodulator.demodulate) for C independent,
// stateful channels per call:
//
//   Q1 cqpsk_agc_kernel    block AGC: mean |x| of the call, gain smoothing and clip (:436-455), NCO phase
//                          accumulator advance (:457-466). One CTA per channel.
//   Q2 cqpsk_front_kernel  x * gain, NCO exp(-j(phase_acc + freq_offset*n)) in float64 when it is active,
//                          63-tap Hamming low-pass as np.convolve(..., "same") per rail per call (:468-471,
//                          zero padded at both ends, no state), float64 accumulation, complex64 result.
//   Q3 cqpsk_sync_kernel   one thread per channel replays the per-sample symbol clock, the 8-tap
//                          Hann-sinc MMSE interpolation, the differential slicer, the frequency loop
//                          and the Gardner TED (:481-669). The reference's 32-entry history ring is
//                          "the last 32 filtered samples": it is indexed straight out of the filtered
//                          call extended by a carried 32-sample tail.
//
// Scalar precision mirrors the reference as executed under NumPy 2 (re-inspected on
// the live object): complex64 symbols with float32 arithmetic and separately rounded products,
// complex-by-real division as multiplication by the float32 reciprocal (numpy scalarmath), float32
// symbol clock / period after the first TED update (Python floats = float64 before it), float64
// frequency offset and phase accumulator, and the float64 first-symbol path (np.conj(0j) is
// complex128). Compiled with -fmad=false.
#include <math.h>
#include <string.h>
#include <vector>

#include "../../include/wcsdr_b200.h"
#include "common.cuh"
#include "seqloop.cuh"

namespace wc {

constexpr int CQ_HIST = 32;     // MMSE_NTAPS (decoders/p25.py:221)
constexpr int CQ_LPF = 63;

__constant__ float c_mmse[129][8];   // _generate_mmse_taps (:289-323), filled at create

struct CqState {
    double freq_offset;   // float64 in the reference
    double phase_acc;
    double clock_d;       // symbol clock while it is still a Python float (before the first TED update)
    float clock_f, sym_time_f, omega_f;
    float agc_gain;
    float2 prev;          // previous symbol (complex64)
    int first;            // 1 until the first symbol has been produced (prev is Python 0j)
    int clock_is_f32;     // 0 until the first TED update
    float2 tail[CQ_HIST]; // last 32 filtered samples, oldest first
};

struct CqChunk {          // per channel, per call: what Q2 needs from Q1
    double fo0, pa0;
    float gain;
    int nco_on;
};

struct CqConst {
    double sps;
    double sym_time0;     // 1.0 / sps (Python float)
    int half_sps, full_sps;
};

// ---- Q1 ----
__global__ void __launch_bounds__(256) cqpsk_agc_kernel(const float2* __restrict__ x, long long stride, int n, CqState* st,
                                                        CqChunk* ck) {
    __shared__ double red[8];
    const int ch = blockIdx.x;
    const float2* xc = x + (long long)ch * stride;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float2 v = xc[i];
        s += (double)(float)sqrt((double)v.x * (double)v.x + (double)v.y * (double)v.y);  // np.abs -> float32
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += red[w];
        CqState& S = st[ch];
        const float mean = (float)(tot / (double)n);
        float g = S.agc_gain;
        if (mean > 1e-8f) {
            const float tg = __fdiv_rn(1.0f, mean);
            g = __fadd_rn(__fmul_rn(g, (float)(1.0 - 0.005)), __fmul_rn(tg, 0.005f));
            g = fminf(fmaxf(g, 0.01f), 500.0f);
        }
        S.agc_gain = g;
        CqChunk c;
        c.gain = g;
        c.fo0 = S.freq_offset;
        c.pa0 = S.phase_acc;
        c.nco_on = fabs(S.freq_offset) > 1e-7 ? 1 : 0;
        if (c.nco_on) {
            double pa = S.phase_acc + S.freq_offset * (double)n;
            double sn, cs;
            sincos(pa, &sn, &cs);
            S.phase_acc = atan2(sn, cs);  // np.angle(np.exp(1j * phase_acc))
        }
        ck[ch] = c;
    }
}

// ---- Q2 ----
constexpr int CQF_THREADS = 256;
constexpr int CQF_PER = 4;
constexpr int CQF_TILE = CQF_THREADS * CQF_PER;

__global__ void __launch_bounds__(CQF_THREADS) cqpsk_front_kernel(const float2* __restrict__ x, long long stride, int n,
                                                                  const CqChunk* __restrict__ ck, const float* __restrict__ lpf,
                                                                  float2* __restrict__ y) {
    __shared__ double2 xs[CQF_TILE + CQ_LPF - 1];
    __shared__ double hs[CQ_LPF];   // firwin(63, 7250/(fs/2), hamming) as float32 (:375-388), promoted
    if (threadIdx.x < CQ_LPF) hs[threadIdx.x] = (double)lpf[threadIdx.x];
    const int ch = blockIdx.y;
    const int n0 = blockIdx.x * CQF_TILE;
    const CqChunk c = ck[ch];
    const float2* xc = x + (long long)ch * stride;
    const bool filt = n >= CQ_LPF;
    const int halo = filt ? (CQ_LPF - 1) / 2 : 0;  // 31 on each side
    const int total = CQF_TILE + 2 * halo;
    for (int i = threadIdx.x; i < total; i += CQF_THREADS) {
        const int g = n0 - halo + i;
        double2 v = make_double2(0.0, 0.0);
        if (g >= 0 && g < n) {
            const float2 s = xc[g];
            const float xr = __fmul_rn(s.x, c.gain), xi = __fmul_rn(s.y, c.gain);  // complex64 * float32
            if (c.nco_on) {
                const double th = c.pa0 + c.fo0 * (double)g;
                double sn, cs;
                sincos(th, &sn, &cs);           // exp(-j th) = cs - j sn
                const double er = cs, ei = -sn;
                v.x = (double)xr * er - (double)xi * ei;
                v.y = (double)xr * ei + (double)xi * er;
            } else {
                v.x = (double)xr;
                v.y = (double)xi;
            }
        }
        xs[i] = v;
    }
    __syncthreads();
    float2* yc = y + (long long)ch * n;
#pragma unroll
    for (int r = 0; r < CQF_PER; ++r) {
        const int o = threadIdx.x + r * CQF_THREADS;
        const int g = n0 + o;
        if (g >= n) continue;
        if (!filt) {
            yc[g] = make_float2((float)xs[o].x, (float)xs[o].y);
            continue;
        }
        // out[m] = sum_k taps[k] * x[m + 31 - k]; xs[o + j] = x[g - 31 + j]  ->  x[g + 31 - k] = xs[o + 62 - k]
        double ar = 0.0, ai = 0.0;
#pragma unroll 9
        for (int k = 0; k < CQ_LPF; ++k) {
            const double h = hs[k];
            const double2 v = xs[o + (CQ_LPF - 1) - k];
            ar = fma(h, v.x, ar);
            ai = fma(h, v.y, ai);
        }
        yc[g] = make_float2((float)ar, (float)ai);
    }
}

// ---- Q3 ----
struct CqSyncArgs {
    int C, n, max_sym;
    CqConst k;
    CqState* st;
    const float2* filt;      // [C][n]
    unsigned char* dibits;   // [C][max_sym]
    int* n_sym;              // [C]
    float2* sym;             // [C][max_sym] interpolated symbols of this call (timing kernel -> slicer kernel)
    double* delta;           // [C][max_sym] frequency-loop increments (slicer kernel -> frequency scan)
};

constexpr int CQ_CH = 8;                // channels per warp: the loop is latency-bound per symbol whatever the lane count, and a
                                        // warp stages the rows of all its channels every tile, so fewer channels per warp = less
                                        // staging per symbol and more SMs in use (64 channels: 8 warps instead of 2)
constexpr int CQ_TILE = 96;             // new samples staged per step
constexpr int CQ_RING = 256;            // ring slots per channel: the tile being walked + its CQ_HIST of look-back + the next
                                        // tile landing meanwhile (samples 256 apart share a slot: 2 * CQ_TILE + CQ_HIST <= 256)
constexpr int CQ_THREADS = 64;          // warp 0 walks the symbols, warp 1 stages the next tile
constexpr int CQ_MIRROR = CQ_HIST;      // slots 0..31 are kept twice (also at 128..159) so that look-back reads never wrap
constexpr int CQ_PITCH = CQ_RING + CQ_MIRROR + 1;   // row pitch in float2: rows of different lanes start in different banks

struct CqView {
    const float2* ring;   // this lane's row of the shared-memory ring: sample j of the call sits in slot (j + 32) & 255,
                          // the 32 carried samples of the previous call in slots 0..31
    const float* mmse;    // shared-memory copy of the 129 x 8 table, rows padded to 9 floats: every thread (channel)
                          // indexes its own row, which would serialise on the constant cache
};
constexpr int CQ_ROW = 9;

// sample `back` positions before sample index m of this call (m - back may reach into the tail); generic, wrapping form
__device__ __forceinline__ float2 cq_hist(const CqView& v, int m, int back) {
    return v.ring[(m - back + CQ_HIST) & (CQ_RING - 1)];
}

// _mmse_interpolate_at_offset (decoders/p25.py:325-359) with the tap row already in registers (the symbol, mid-point and
// previous-symbol interpolations of one symbol share it: same mu). `p` points at the ring slot of the sample `back`
// positions before the current one, in the unwrapped half of the mirrored ring: p[-(tap - 3)] is tap's sample, a load
// with an immediate offset. Products rounded separately, summed left to right in float32 like the reference's loop.
// FULL: all 8 taps lie inside the 32-sample history (3 <= back, back + 4 < 32); otherwise taps whose offset is negative
// are skipped (back = 0: taps 3..7).
template <bool FULL>
__device__ __forceinline__ float2 cq_interp(const float2* __restrict__ p, int back, const float (&t)[8]) {
    float rr = 0.f, ri = 0.f;
    bool any = false;
#pragma unroll
    for (int tap = 0; tap < 8; ++tap) {
        const int off = back + (tap - 3);
        if (FULL || (off >= 0 && off < CQ_HIST)) {
            const float2 h = p[-(tap - 3)];
            const float pr = __fmul_rn(t[tap], h.x), pi = __fmul_rn(t[tap], h.y);
            if (!any) {
                rr = pr;
                ri = pi;
                any = true;
            } else {
                rr = __fadd_rn(rr, pr);
                ri = __fadd_rn(ri, pi);
            }
        }
    }
    return make_float2(rr, ri);
}

__device__ __forceinline__ float cq_abs(float2 z) {
    return (float)sqrt((double)z.x * (double)z.x + (double)z.y * (double)z.y);
}

__global__ void __launch_bounds__(CQ_THREADS) cqpsk_sync_kernel(const CqSyncArgs a) {
    __shared__ float s_mmse[129 * CQ_ROW];
    __shared__ float2 s_ring[CQ_CH * CQ_PITCH];
    for (int i = threadIdx.x; i < 129 * 8; i += blockDim.x) s_mmse[(i >> 3) * CQ_ROW + (i & 7)] = c_mmse[i >> 3][i & 7];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * CQ_CH;
    const int ch = c0 + lane;
    const bool live = warp == 0 && lane < CQ_CH && ch < a.C;
    // the 32 carried samples of every row
    if (warp == 0) {
        for (int r = 0; r < CQ_CH && c0 + r < a.C; ++r) {
            const float2 tv = a.st[c0 + r].tail[lane];
            s_ring[r * CQ_PITCH + lane] = tv;
            s_ring[r * CQ_PITCH + CQ_RING + lane] = tv;
        }
    }
    // tile t of every row -> ring slots, 8-byte LDGSTS copies, lanes along the sample axis (warp 1 only)
    auto stage = [&](int t) {
        const int base = t * CQ_TILE;
        const int lim = min(CQ_TILE, a.n - base);
        for (int r = 0; r < CQ_CH && c0 + r < a.C; ++r) {
            const float2* xr = a.filt + (long long)(c0 + r) * a.n + base;
            float2* row = s_ring + r * CQ_PITCH;
#pragma unroll
            for (int k = 0; k < CQ_TILE / 32; ++k) {
                const int i = lane + 32 * k;
                if (i < lim) {
                    const int slot = (base + i + CQ_HIST) & (CQ_RING - 1);
                    cp_async8(&row[slot], xr + i);
                    if (slot < CQ_MIRROR) cp_async8(&row[slot + CQ_RING], xr + i);
                }
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
    };
    const int T = (a.n + CQ_TILE - 1) / CQ_TILE;
    if (warp == 1 && T > 0) stage(0);
    __syncthreads();
    CqState S;
    if (live) S = a.st[ch];
    else memset(&S, 0, sizeof(S));
    CqView v;
    v.mmse = s_mmse;
    v.ring = s_ring + (lane < CQ_CH ? lane : 0) * CQ_PITCH;
    float2* sym_out = a.sym + (long long)(live ? ch : 0) * a.max_sym;
    const float omega_lo = (float)(a.k.sps * 0.95), omega_hi = (float)(a.k.sps * 1.05);
    const bool full_taps = a.k.half_sps >= 3 && a.k.full_sps + 4 < CQ_HIST;   // every Gardner tap inside the history
    int nsym = 0;
    // Tile loop: warp 1 stages tile t + 1 into the ring while warp 0 walks tile t (one CTA barrier per tile; with the copies
    // and their wait on the walking warp they were 19 % of the kernel). Inside a tile the loop is symbol-driven: every lane
    // first advances its own sample clock to its next firing sample, then the lanes that fired run the expensive symbol body
    // together — a plain per-sample loop would execute the body at almost every sample index because the channels' clocks
    // are not aligned.
    int m = -1;
    for (int t = 0; t < T; ++t) {
        const int tile_end = min((t + 1) * CQ_TILE, a.n);
        if (warp == 1) {
            if (t + 1 < T) stage(t + 1);
        } else {
    for (;;) {
        bool fire = false;
        if (live && S.clock_is_f32) {
            // the clock advances by one rounded float32 addition per sample (exactly the reference's sequence); eight
            // samples per trip in straight-line code (clock_run8): the additions stay a dependent chain, the compare /
            // branch per sample goes (a single warp serves its channels, so this loop's latency is the demodulator's;
            // the four-wide predecessor spent 48 % of the kernel here, five branches per trip in the ncu source view)
            float cf = S.clock_f;
            while (m + 1 < tile_end) {
                m += clock_run<true>(cf, S.sym_time_f, tile_end - 1 - m, fire);
                if (fire) break;
            }
            S.clock_f = cf;
        } else {
            while (live && m + 1 < tile_end) {
                ++m;
                S.clock_d += a.k.sym_time0;
                fire = S.clock_d >= 1.0;
                if (fire) break;
            }
        }
        if (!__any_sync(0xffffffffu, fire)) break;
        if (!fire) continue;
        int imu;
        if (S.clock_is_f32) {
            S.clock_f = __fsub_rn(S.clock_f, 1.0f);
            float mu = __fdiv_rn(S.clock_f, S.sym_time_f);
            mu = fminf(fmaxf(mu, 0.0f), (float)(1.0 - 1e-6));
            imu = (int)rintf(__fmul_rn(mu, 128.0f));
        } else {
            S.clock_d -= 1.0;
            double mu = S.clock_d / a.k.sym_time0;
            mu = fmin(fmax(mu, 0.0), 1.0 - 1e-6);
            imu = (int)rint(mu * 128.0);
        }
        imu = min(imu, 128);
        float taps[8];
#pragma unroll
        for (int tap = 0; tap < 8; ++tap) taps[tap] = v.mmse[imu * CQ_ROW + tap];
        // the three interpolations are independent: issue them together, the Gardner error needs all of them
        int slot = (m + CQ_HIST) & (CQ_RING - 1);
        if (slot < CQ_MIRROR) slot += CQ_RING;                    // mirrored copy: slot - 31 .. slot are all valid, no wrap
        const float2* pc = v.ring + slot;
        const float2 curr = cq_interp<false>(pc, 0, taps);
        float2 mid, prv;
        if (full_taps) {
            mid = cq_interp<true>(pc - a.k.half_sps, a.k.half_sps, taps);
            prv = cq_interp<true>(pc - a.k.full_sps, a.k.full_sps, taps);
        } else {
            mid = cq_interp<false>(pc - a.k.half_sps, a.k.half_sps, taps);
            prv = cq_interp<false>(pc - a.k.full_sps, a.k.full_sps, taps);
        }
        // The differential slicer and the frequency loop (decoders/p25.py:540-585) read this symbol and the previous one
        // but feed nothing back into the symbol clock within a call (the frequency offset only steers the NEXT call's
        // NCO): they run afterwards, in parallel over all symbols (cqpsk_slice_kernel) with a short per-channel scan for
        // the clamped accumulation (cqpsk_freq_kernel). What stays here is the genuinely sequential part: clock,
        // interpolation, Gardner error. Measured on B200, 64 channels x 72 000 samples: 14.5 ms with the slicer inline.
        if (nsym < a.max_sym) sym_out[nsym] = curr;
        ++nsym;
        // Gardner TED (decoders/p25.py:587-608); full_sps + 4 < 32 is guaranteed at create
        {
            const float er = __fsub_rn(curr.x, prv.x), ei = __fsub_rn(curr.y, prv.y);
            // real((e) * conj(mid)) = er*mr - ei*(-mi)
            const float ted = __fsub_rn(__fmul_rn(er, mid.x), __fmul_rn(ei, -mid.y));
            const float adj = __fmul_rn(0.015f, ted);
            const bool was_f64 = !S.clock_is_f32;
            if (!S.clock_is_f32) {
                S.clock_f = (float)S.clock_d;   // Python float + np.float32 -> float32
                S.omega_f = (float)a.k.sps;
                S.clock_is_f32 = 1;
            }
            S.clock_f = __fadd_rn(S.clock_f, adj);
            S.omega_f = __fadd_rn(S.omega_f, __fmul_rn(0.0f, ted));
            const float om = fminf(fmaxf(S.omega_f, omega_lo), omega_hi);
            if (om != S.omega_f || was_f64) {   // gain_omega is 0 in the reference (:274): omega only moves when ted is not finite
                S.omega_f = om;
                S.sym_time_f = __fdiv_rn(1.0f, S.omega_f);
            }
        }
        while (S.clock_f >= 1.0f) S.clock_f = __fsub_rn(S.clock_f, 1.0f);
        while (S.clock_f < 0.0f) S.clock_f = __fadd_rn(S.clock_f, 1.0f);
    }
        }
        __syncthreads();
    }
    if (!live) return;
    // carry the ring: last 32 of (tail ++ x)
    float2 nt[CQ_HIST];
    for (int i = 0; i < CQ_HIST; ++i) nt[i] = cq_hist(v, a.n - 1, CQ_HIST - 1 - i);
    CqState* G = &a.st[ch];
    for (int i = 0; i < CQ_HIST; ++i) G->tail[i] = nt[i];
    G->clock_d = S.clock_d;
    G->clock_f = S.clock_f;
    G->sym_time_f = S.sym_time_f;
    G->omega_f = S.omega_f;
    G->clock_is_f32 = S.clock_is_f32;
    a.n_sym[ch] = min(nsym, a.max_sym);
}

// Differential slicer (decoders/p25.py:540-585), one thread per (channel, symbol): normalised curr * conj(prev), the
// correctly rounded float32 phase, dibit, and the frequency-loop increment 0.0005 * phase_error * |curr| as float64.
// Symbol 0 of a call pairs with the state's previous symbol; the very first symbol after a reset pairs with Python's 0j
// and follows the reference's float64 path.
__global__ void cqpsk_slice_kernel(const CqSyncArgs a) {
    const int ch = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n_sym[ch]) return;
    const float HALF_PI_F = (float)1.5707963267948966, PI_F = (float)3.141592653589793, TWO_PI_F = (float)6.283185307179586;
    const float Q_PI_F = (float)0.7853981633974483, TQ_PI_F = (float)2.356194490192345;
    const double PI_D = 3.141592653589793;
    const float2* sym = a.sym + (long long)ch * a.max_sym;
    const float2 curr = sym[k];
    const float cm = cq_abs(curr);
    int dibit;
    double delta;
    if (k == 0 && a.st[ch].first) {
        // prev is Python 0j: diff = curr * np.conj(0j) is complex128 (+-0); phase = arctan2(+-0, +-0) in float64
        const double re0 = (double)curr.x * 0.0 - (double)curr.y * (-0.0);
        const double im0 = (double)curr.x * (-0.0) + (double)curr.y * 0.0;
        const double phase = atan2(im0, re0);
        double expected;
        if (phase >= 1.5707963267948966) {
            dibit = 1;
            expected = 2.356194490192345;
        } else if (phase >= 0.0) {
            dibit = 0;
            expected = 0.7853981633974483;
        } else if (phase >= -1.5707963267948966) {
            dibit = 2;
            expected = -0.7853981633974483;
        } else {
            dibit = 3;
            expected = -2.356194490192345;
        }
        double pe = phase - expected;
        if (pe > PI_D) pe -= 2.0 * PI_D;
        else if (pe < -PI_D) pe += 2.0 * PI_D;
        delta = (0.0005 * pe) * (double)cm;
    } else {
        const float2 prev = (k == 0) ? a.st[ch].prev : sym[k - 1];
        const float pm = cq_abs(prev);
        float dr, di;
        if (cm > 1e-6f && pm > 1e-6f) {
            const float sc = __fdiv_rn(1.0f, cm), sp = __fdiv_rn(1.0f, pm);
            const float ar = __fmul_rn(curr.x, sc), ai = __fmul_rn(curr.y, sc);
            const float br = __fmul_rn(prev.x, sp), bi = -__fmul_rn(prev.y, sp);  // conj
            dr = __fsub_rn(__fmul_rn(ar, br), __fmul_rn(ai, bi));
            di = __fadd_rn(__fmul_rn(ar, bi), __fmul_rn(ai, br));
        } else {
            const float br = prev.x, bi = -prev.y;
            dr = __fsub_rn(__fmul_rn(curr.x, br), __fmul_rn(curr.y, bi));
            di = __fadd_rn(__fmul_rn(curr.x, bi), __fmul_rn(curr.y, br));
        }
        const float phase = (float)atan2((double)di, (double)dr);
        float expected;
        if (phase >= HALF_PI_F) {
            dibit = 1;
            expected = TQ_PI_F;
        } else if (phase >= 0.0f) {
            dibit = 0;
            expected = Q_PI_F;
        } else if (phase >= -HALF_PI_F) {
            dibit = 2;
            expected = -Q_PI_F;
        } else {
            dibit = 3;
            expected = -TQ_PI_F;
        }
        float pe = __fsub_rn(phase, expected);
        if (pe > PI_F) pe = __fsub_rn(pe, TWO_PI_F);
        else if (pe < -PI_F) pe = __fadd_rn(pe, TWO_PI_F);
        delta = (double)__fmul_rn(__fmul_rn(0.0005f, pe), cm);
    }
    a.dibits[(long long)ch * a.max_sym + k] = (unsigned char)dibit;
    a.delta[(long long)ch * a.max_sym + k] = delta;
}

// Frequency loop state: freq_offset = clip(freq_offset + delta_k, +-0.02) symbol by symbol (:583-585) — the clip makes it
// order dependent, so one thread per channel replays the additions (two dependent float64 operations per symbol);
// previous-symbol / first-symbol state for the next call.
__global__ void __launch_bounds__(32) cqpsk_freq_kernel(const CqSyncArgs a) {
    const int ch = blockIdx.x, lane = threadIdx.x;
    const int n = a.n_sym[ch];
    if (n <= 0) return;
    const double* d = a.delta + (long long)ch * a.max_sym;
    double f = a.st[ch].freq_offset;
    // one warp per channel: 32 increments per coalesced load, broadcast one by one; every lane replays the same additions
    for (int k0 = 0; k0 < n; k0 += 32) {
        const double mine = (k0 + lane < n) ? d[k0 + lane] : 0.0;
        const int cnt = min(32, n - k0);
        for (int k = 0; k < cnt; ++k) f = fmin(fmax(f + __shfl_sync(0xffffffffu, mine, k), -0.02), 0.02);
    }
    if (lane == 0) {
        a.st[ch].freq_offset = f;
        a.st[ch].prev = a.sym[(long long)ch * a.max_sym + n - 1];
        a.st[ch].first = 0;
    }
}

__global__ void cqpsk_reset_kernel(CqState* st, int lo, int hi) {
    const int ch = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= hi) return;
    CqState s;
    memset(&s, 0, sizeof(s));
    s.agc_gain = 1.0f;
    s.first = 1;
    st[ch] = s;
}

}  // namespace wc

using namespace wc;

struct wc_cqpsk {
    int C = 0, sample_rate = 0, symbol_rate = 0;
    CqConst k;
    CqState* d_state = nullptr;
    CqChunk* d_chunk = nullptr;
    float* d_lpf = nullptr;
    float2* d_filt = nullptr;  size_t filt_cap = 0;
    void* d_in = nullptr;      size_t in_cap = 0;
    unsigned char* d_dib = nullptr; size_t dib_cap = 0;
    int* d_nsym = nullptr;
    float2* d_sym = nullptr;   size_t sym_cap = 0;     // [C][max_sym] symbols of the call
    double* d_delta = nullptr;
    cudaStream_t stream = nullptr;
};

static void hamming_lowpass63(double cutoff_norm, float* out) {
    double h[CQ_LPF], s = 0.0;
    const double alpha = 0.5 * (CQ_LPF - 1);
    for (int n = 0; n < CQ_LPF; ++n) {
        const double m = n - alpha, xx = cutoff_norm * m;
        const double sinc = (xx == 0.0) ? 1.0 : sin(M_PI * xx) / (M_PI * xx);
        h[n] = cutoff_norm * sinc * (0.54 - 0.46 * cos(2.0 * M_PI * n / (CQ_LPF - 1)));
        s += h[n];
    }
    for (int n = 0; n < CQ_LPF; ++n) out[n] = (float)(h[n] / s);
}

extern "C" {

int wc_cqpsk_create(int n_channels, int sample_rate, int symbol_rate, const float* lpf_taps63, const float* mmse_taps_129x8,
                    wc_cqpsk** out) {
    WC_REQUIRE(out != nullptr, "wc_cqpsk_create: out is null");
    WC_REQUIRE(n_channels >= 1 && sample_rate > 0 && symbol_rate > 0, "wc_cqpsk_create: bad parameters");
    const double sps = (double)sample_rate / (double)symbol_rate;
    const int full_sps = (int)rint(sps), half_sps = (int)rint(sps / 2.0);  // Python round(): half to even
    WC_REQUIRE(sps >= 2.0 && full_sps + 4 < CQ_HIST, "wc_cqpsk_create: samples per symbol %.3f outside [2, 27]", sps);
    wc_cqpsk* h = new wc_cqpsk();
    h->C = n_channels;
    h->sample_rate = sample_rate;
    h->symbol_rate = symbol_rate;
    h->k.sps = sps;
    h->k.sym_time0 = 1.0 / sps;
    h->k.half_sps = half_sps;
    h->k.full_sps = full_sps;
    float lpf[CQ_LPF];
    if (lpf_taps63) memcpy(lpf, lpf_taps63, sizeof(lpf));
    else {
        double c = 7250.0 / (sample_rate / 2.0);
        c = c > 0.99 ? 0.99 : (c < 0.01 ? 0.01 : c);
        hamming_lowpass63(c, lpf);
    }
    std::vector<float> mm(129 * 8);
    if (mmse_taps_129x8) memcpy(mm.data(), mmse_taps_129x8, sizeof(float) * 129 * 8);
    else {
        // _generate_mmse_taps (decoders/p25.py:289-323)
        for (int step = 0; step <= 128; ++step) {
            const double mu = (double)step / 128.0;
            float row[8];
            for (int tap = 0; tap < 8; ++tap) {
                const double t = tap - 3 - mu;
                if (fabs(t) < 1e-6) row[tap] = 1.0f;
                else {
                    const double sv = sin(M_PI * t) / (M_PI * t);
                    const double w = fabs(t) < 4 ? 0.5 * (1 + cos(M_PI * t / 4)) : 0.0;
                    row[tap] = (float)(sv * w);
                }
            }
            float s = 0.f;
            for (int tap = 0; tap < 8; ++tap) s += row[tap];
            for (int tap = 0; tap < 8; ++tap) mm[step * 8 + tap] = (fabsf(s) > 1e-6f) ? row[tap] / s : row[tap];
        }
    }
    bool ok = cudaMalloc(&h->d_state, sizeof(CqState) * n_channels) == cudaSuccess &&
              cudaMalloc(&h->d_chunk, sizeof(CqChunk) * n_channels) == cudaSuccess &&
              cudaMalloc(&h->d_nsym, sizeof(int) * n_channels) == cudaSuccess &&
              cudaMalloc(&h->d_lpf, sizeof(lpf)) == cudaSuccess &&
              cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMemcpy(h->d_lpf, lpf, sizeof(lpf), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpyToSymbol(c_mmse, mm.data(), sizeof(float) * 129 * 8) == cudaSuccess;
    if (!ok) {
        set_error("wc_cqpsk_create: CUDA setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete h;
        return -2;
    }
    cqpsk_reset_kernel<<<(n_channels + 63) / 64, 64, 0, h->stream>>>(h->d_state, 0, n_channels);
    WC_CUDA(cudaStreamSynchronize(h->stream));
    *out = h;
    return 0;
}

void wc_cqpsk_destroy(wc_cqpsk* h) {
    if (!h) return;
    cudaFree(h->d_state);
    cudaFree(h->d_chunk);
    cudaFree(h->d_nsym);
    cudaFree(h->d_lpf);
    if (h->d_filt) cudaFree(h->d_filt);
    if (h->d_in) cudaFree(h->d_in);
    if (h->d_dib) cudaFree(h->d_dib);
    if (h->d_sym) cudaFree(h->d_sym);
    if (h->d_delta) cudaFree(h->d_delta);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int wc_cqpsk_max_symbols(const wc_cqpsk* h, int n_samples) {
    if (!h || n_samples <= 0) return 0;
    const long long m = (long long)((double)n_samples / floor(h->k.sps) * 1.25) + 16;
    return (int)(m < n_samples ? m : n_samples);
}

int wc_cqpsk_reset(wc_cqpsk* h, int channel) {
    WC_REQUIRE(h != nullptr, "wc_cqpsk_reset: null handle");
    WC_REQUIRE(channel >= -1 && channel < h->C, "wc_cqpsk_reset: channel %d out of range", channel);
    const int lo = channel < 0 ? 0 : channel, hi = channel < 0 ? h->C : channel + 1;
    cqpsk_reset_kernel<<<(hi - lo + 63) / 64, 64, 0, h->stream>>>(h->d_state, lo, hi);
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int wc_cqpsk_demod(wc_cqpsk* h, const void* iq_dev, long long chan_stride, int n_samples, unsigned char* dibits_dev,
                   int* n_sym_dev, int max_sym, void* stream_v) {
    WC_REQUIRE(h && iq_dev && dibits_dev && n_sym_dev, "wc_cqpsk_demod: null argument");
    WC_REQUIRE(n_samples >= 0 && chan_stride >= n_samples, "wc_cqpsk_demod: bad sizes");
    cudaStream_t s = (cudaStream_t)stream_v;
    const int C = h->C;
    if (n_samples == 0) {
        WC_CUDA(cudaMemsetAsync(n_sym_dev, 0, sizeof(int) * C, s));
        return 0;
    }
    WC_REQUIRE(max_sym >= wc_cqpsk_max_symbols(h, n_samples), "wc_cqpsk_demod: max_sym %d < %d", max_sym,
               wc_cqpsk_max_symbols(h, n_samples));
    const size_t need = (size_t)C * n_samples;
    if (h->filt_cap < need) {
        if (h->d_filt) cudaFree(h->d_filt);
        h->d_filt = nullptr;
        h->filt_cap = 0;
        WC_CUDA(cudaMalloc((void**)&h->d_filt, need * sizeof(float2)));
        h->filt_cap = need;
    }
    const size_t sneed = (size_t)C * max_sym;
    if (h->sym_cap < sneed) {
        if (h->d_sym) cudaFree(h->d_sym);
        if (h->d_delta) cudaFree(h->d_delta);
        h->d_sym = nullptr;
        h->d_delta = nullptr;
        h->sym_cap = 0;
        WC_CUDA(cudaMalloc((void**)&h->d_sym, sneed * sizeof(float2)));
        WC_CUDA(cudaMalloc((void**)&h->d_delta, sneed * sizeof(double)));
        h->sym_cap = sneed;
    }
    const float2* x = reinterpret_cast<const float2*>(iq_dev);
    cqpsk_agc_kernel<<<C, 256, 0, s>>>(x, chan_stride, n_samples, h->d_state, h->d_chunk);
    dim3 fg((n_samples + CQF_TILE - 1) / CQF_TILE, C);
    cqpsk_front_kernel<<<fg, CQF_THREADS, 0, s>>>(x, chan_stride, n_samples, h->d_chunk, h->d_lpf, h->d_filt);
    CqSyncArgs a;
    a.C = C;
    a.n = n_samples;
    a.max_sym = max_sym;
    a.k = h->k;
    a.st = h->d_state;
    a.filt = h->d_filt;
    a.dibits = dibits_dev;
    a.n_sym = n_sym_dev;
    a.sym = h->d_sym;
    a.delta = h->d_delta;
    cqpsk_sync_kernel<<<(C + CQ_CH - 1) / CQ_CH, CQ_THREADS, 0, s>>>(a);
    cqpsk_slice_kernel<<<dim3((max_sym + 127) / 128, C), 128, 0, s>>>(a);
    cqpsk_freq_kernel<<<C, 32, 0, s>>>(a);
    WC_CUDA(cudaGetLastError());
    return 0;
}

int wc_cqpsk_demod_host(wc_cqpsk* h, const void* iq_host, int n_samples, unsigned char* dibits_host, int* n_sym_host,
                        int max_sym) {
    WC_REQUIRE(h && iq_host && dibits_host && n_sym_host, "wc_cqpsk_demod_host: null argument");
    const int C = h->C;
    if (n_samples <= 0) {
        for (int c = 0; c < C; ++c) n_sym_host[c] = 0;
        return 0;
    }
    const size_t in_bytes = sizeof(float2) * (size_t)C * n_samples;
    if (h->in_cap < in_bytes) {
        if (h->d_in) cudaFree(h->d_in);
        h->d_in = nullptr;
        h->in_cap = 0;
        WC_CUDA(cudaMalloc(&h->d_in, in_bytes));
        h->in_cap = in_bytes;
    }
    const size_t ob = (size_t)C * max_sym;
    if (h->dib_cap < ob) {
        if (h->d_dib) cudaFree(h->d_dib);
        h->d_dib = nullptr;
        h->dib_cap = 0;
        WC_CUDA(cudaMalloc((void**)&h->d_dib, ob));
        h->dib_cap = ob;
    }
    WC_CUDA(cudaMemcpyAsync(h->d_in, iq_host, in_bytes, cudaMemcpyHostToDevice, h->stream));
    int rc = wc_cqpsk_demod(h, h->d_in, n_samples, n_samples, h->d_dib, h->d_nsym, max_sym, h->stream);
    if (rc) return rc;
    WC_CUDA(cudaMemcpyAsync(dibits_host, h->d_dib, ob, cudaMemcpyDeviceToHost, h->stream));
    WC_CUDA(cudaMemcpyAsync(n_sym_host, h->d_nsym, sizeof(int) * C, cudaMemcpyDeviceToHost, h->stream));
    WC_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

/* state6 = {freq_offset, phase_acc, symbol_clock, symbol_time, agc_gain, omega} */
int wc_cqpsk_get_state(wc_cqpsk* h, int channel, double* state6) {
    WC_REQUIRE(h && state6, "wc_cqpsk_get_state: null argument");
    WC_REQUIRE(channel >= 0 && channel < h->C, "wc_cqpsk_get_state: channel %d out of range", channel);
    CqState s;
    WC_CUDA(cudaDeviceSynchronize());
    WC_CUDA(cudaMemcpy(&s, h->d_state + channel, sizeof(s), cudaMemcpyDeviceToHost));
    state6[0] = s.freq_offset;
    state6[1] = s.phase_acc;
    state6[2] = s.clock_is_f32 ? (double)s.clock_f : s.clock_d;
    state6[3] = s.clock_is_f32 ? (double)s.sym_time_f : h->k.sym_time0;
    state6[4] = s.agc_gain;
    state6[5] = s.clock_is_f32 ? (double)s.omega_f : h->k.sps;
    return 0;
}

}  // extern "C"
