// Voice-channel discriminator path, batched over channels:
//   wc_fm_discriminator   trunking/system.py:708-717   np.diff(np.unwrap([last_phase | np.angle(iq)])), last phase carried
//   wc_discdemod_*        decoders/p25.py:1105-1345    DiscriminatorDemodulator.demodulate: auto gain, DC tracker,
//                                                      65-tap low-pass (np.convolve 'same' per call), MMSE interpolation
//                                                      with timing / spread / frequency loops, 4-level slicer
//
// The reference mixes Python floats with np.float32 values; under the NumPy >= 2 promotion rules every state variable
// that has met a float32 IS a float32 (oracle/discriminator.py documents the flow and reports it via state_dtypes()).
// Two variables keep a Python-float phase that matters numerically and is modelled explicitly:
//   clock   a Python float (float64 accumulation of symbol_time) from construction / reset() until the first symbol,
//           float32 afterwards;
//   spread  the literal 1.6 / 2.4 whenever max(1.6, min(2.4, spread)) clamps — then `1.5 * spread` is a float64 product
//           rounded once instead of a float32 product.
// Compiled with -fmad=false: every float32 operation of the reference rounds on its own.
//
// Kernels: peak (parallel), gain + DC tracker (sequential per channel: DD_CH channels per CTA, one lane each on the walking
// warp, tiles moved by helper warps), low-pass (parallel, float64 accumulation, rounded once), MMSE loop (sequential per
// channel, same layout; profiles/r03_seq_notes.md has the measurements behind that layout).
#include <math.h>
#include <string.h>
#include <vector>

#include "../../include/wcsdr_b200.h"
#include "common.cuh"
#include "seqloop.cuh"

namespace wc {

constexpr int DD_TAPS = 8, DD_STEPS = 128, DD_LPF = 65;
constexpr int DD_CH = 8;              // channels per CTA (one lane each on the walking warp): the loops are latency-bound per symbol
                                      // whatever the lane count, and fewer lanes = fewer shared-memory bank conflicts between
                                      // rows read at different offsets, fewer lanes to wait for at a tile end, more SMs in use
constexpr int DD_TILE = 512;          // samples staged per step (MMSE loop)

struct DDState {
    double clock_d;     // valid while clock_py
    float clock_f;
    int clock_py;
    float spread;       // float32 value (1.6f / 2.4f while spread_py)
    int spread_py;      // 0 = float32, 1 = literal 1.6, 2 = literal 2.4
    float fine, coarse, dc, gain;
    float hist[DD_TAPS];
    int hidx;
    long long symbols;
};

struct DDConst {
    double symbol_time;    // symbol_rate / sample_rate
    float symbol_time_f;
};

__global__ void dd_peak_kernel(const float* __restrict__ x, long long stride, int n, float* __restrict__ peak) {
    __shared__ float red[32];
    const int ch = blockIdx.x;
    const float* xc = x + (long long)ch * stride;
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(xc[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) peak[ch] = m;
    }
}

// auto gain (p25.py:1213-1222) and DC removal (:1224-1229); y = x*gain - dc, float32 throughout
// Layout of the two sequential kernels: warp 0 walks the rows (lane = channel); three helper warps move the tiles between
// global and shared memory (4-byte LDGSTS in, coalesced stores out), one CTA barrier per tile. With the copies on the
// walking warp they were 80 % of its instructions (ncu source view, 36 instructions per sample against 6 of arithmetic).
constexpr int DD_HELPERS = 3;
constexpr int DD_THREADS = 32 * (1 + DD_HELPERS);
constexpr int DC_TILE = 496, DC_ROW = DC_TILE + 1;   // three tiles of DD_CH rows inside the 48 KB static limit; odd pitch

template <int TILE, int PITCH, int COL0>
__device__ __forceinline__ void dd_helper_load(float (*tile)[PITCH], const float* __restrict__ x, long long stride, int c0, int C,
                                               int base, int n, int hw, int lane) {
    const int lim = min(TILE, n - base);
    for (int r = hw; r < DD_CH && c0 + r < C; r += DD_HELPERS) {
        const float* xr = x + (long long)(c0 + r) * stride + base;
#pragma unroll
        for (int k = 0; k < (TILE + 31) / 32; ++k) {
            const int i = lane + 32 * k;
            if (i < lim) cp_async4(&tile[r][COL0 + i], xr + i);
        }
    }
    cp_async_commit();
}

__global__ void __launch_bounds__(DD_THREADS) dd_dc_kernel(const float* __restrict__ x, long long stride, int n, int C,
                                                           const float* __restrict__ peak, DDState* __restrict__ st,
                                                           float* __restrict__ y) {
    __shared__ float tile[3][DD_CH][DC_ROW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * DD_CH;
    const int c = c0 + lane;
    const bool live = warp == 0 && lane < DD_CH && c < C;
    float gain = 1.f, dc = 0.f;
    if (live) {
        gain = st[c].gain;
        dc = st[c].dc;
        const float pk = peak[c];
        if (n > 100 && pk > 0.01f)
            gain = __fadd_rn(__fmul_rn(gain, 0.9f), __fmul_rn(__fdiv_rn(3.0f, pk), 0.1f));
    }
    const int T = (n + DC_TILE - 1) / DC_TILE;
    if (warp > 0) {
        dd_helper_load<DC_TILE, DC_ROW, 0>(tile[0], x, stride, c0, C, 0, n, warp - 1, lane);
        cp_async_wait<0>();
    }
    __syncthreads();
    for (int t = 0; t <= T; ++t) {
        if (warp == 0) {
            if (live && t < T) {
                float* row = tile[t % 3][lane];   // live lanes only: lane < DD_CH
                const int lim = min(DC_TILE, n - t * DC_TILE);
                int i = 0;
                for (; i + 4 <= lim; i += 4) {
                    const float v0 = __fmul_rn(row[i], gain), v1 = __fmul_rn(row[i + 1], gain);
                    const float v2 = __fmul_rn(row[i + 2], gain), v3 = __fmul_rn(row[i + 3], gain);
                    const float w0 = __fmul_rn(v0, 0.001f), w1 = __fmul_rn(v1, 0.001f);
                    const float w2 = __fmul_rn(v2, 0.001f), w3 = __fmul_rn(v3, 0.001f);
                    const float d0 = __fadd_rn(__fmul_rn(dc, 0.999f), w0);
                    const float d1 = __fadd_rn(__fmul_rn(d0, 0.999f), w1);
                    const float d2 = __fadd_rn(__fmul_rn(d1, 0.999f), w2);
                    dc = __fadd_rn(__fmul_rn(d2, 0.999f), w3);
                    row[i] = __fsub_rn(v0, d0);
                    row[i + 1] = __fsub_rn(v1, d1);
                    row[i + 2] = __fsub_rn(v2, d2);
                    row[i + 3] = __fsub_rn(v3, dc);
                }
                for (; i < lim; ++i) {
                    const float v = __fmul_rn(row[i], gain);
                    dc = __fadd_rn(__fmul_rn(dc, 0.999f), __fmul_rn(v, 0.001f));
                    row[i] = __fsub_rn(v, dc);
                }
            }
        } else {
            const int hw = warp - 1;
            if (t + 1 < T) dd_helper_load<DC_TILE, DC_ROW, 0>(tile[(t + 1) % 3], x, stride, c0, C, (t + 1) * DC_TILE, n, hw, lane);
            if (t >= 1) {   // tile t-1 is finished: out, coalesced
                const int base = (t - 1) * DC_TILE;
                const int lim = min(DC_TILE, n - base);
                for (int r = hw; r < DD_CH && c0 + r < C; r += DD_HELPERS) {
                    float* yr = y + (long long)(c0 + r) * n + base;
                    const float* tr = tile[(t - 1) % 3][r];
#pragma unroll
                    for (int k = 0; k < (DC_TILE + 31) / 32; ++k) {
                        const int i = lane + 32 * k;
                        if (i < lim) yr[i] = tr[i];
                    }
                }
            }
            cp_async_wait<0>();
        }
        __syncthreads();
    }
    if (live) {
        st[c].gain = gain;
        st[c].dc = dc;
    }
}

// np.convolve(x, h, 'same') for the odd-length kernel: out[i] = sum_k h[k] x[i + 32 - k], zero outside the call;
// skipped (copy) when the call is shorter than the filter (p25.py:1231-1233)
__global__ void __launch_bounds__(256) dd_lpf_kernel(const float* __restrict__ x, int n, const double* __restrict__ h,
                                                     float* __restrict__ y) {
    __shared__ double sh[DD_LPF];
    if (threadIdx.x < DD_LPF) sh[threadIdx.x] = h[threadIdx.x];
    __syncthreads();
    const int ch = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* xc = x + (long long)ch * n;
    if (n < DD_LPF) {
        y[(long long)ch * n + i] = xc[i];
        return;
    }
    double acc = 0.0;
#pragma unroll 5
    for (int k = 0; k < DD_LPF; ++k) {
        const int j = i + 32 - k;
        if (j >= 0 && j < n) acc += sh[k] * (double)xc[j];
    }
    y[(long long)ch * n + i] = (float)acc;
}

constexpr int DD_PAD = DD_TAPS;                 // carried history in front of every staged row
constexpr int DD_ROWH = DD_PAD + DD_TILE + 1;   // 521 floats: odd pitch, the channel rows start in different banks
constexpr int DD_TROW = DD_TAPS + 1;            // interpolator rows padded to 9 floats (lanes index different rows)

__device__ __forceinline__ float dd_interp(const float* __restrict__ taps, const float* __restrict__ w, int imu) {
    // p25.py:1238-1248: float32 products summed in order into a float32 accumulator (0.0 + first product = that product);
    // w = the last 8 samples, oldest first (the reference's ring read from hist_idx onwards)
    float acc = __fmul_rn(taps[imu * DD_TROW], w[0]);
#pragma unroll
    for (int i = 1; i < DD_TAPS; ++i) acc = __fadd_rn(acc, __fmul_rn(taps[imu * DD_TROW + i], w[i]));
    return acc;
}

// One symbol of _mmse_timing_recovery (p25.py:1262-1333) for a lane whose clock just crossed 1: interpolate at mu and
// mu + 1/128, slicer error, spread / clock / frequency loop updates. Straight-line: every alternative of the reference's
// if-chains is computed and selected. PY = some lane may still carry the Python-float clock of a fresh / reset demodulator.
template <bool PY>
__device__ __forceinline__ void dd_symbol(DDState& s, const DDConst& k, const float* __restrict__ taps, const float* __restrict__ w8,
                                          unsigned char* __restrict__ dout, float* __restrict__ sout, int max_sym, int& count) {
    int imu, imu1;
    if (PY && s.clock_py) {
        s.clock_d -= 1.0;
        double mu = s.clock_d / k.symbol_time;
        if (1.0 < mu) mu = 1.0;
        imu = min((int)rint(mu * 128.0), DD_STEPS);
        double m1 = mu + 0.0078125;
        if (1.0 < m1) m1 = 1.0;
        imu1 = min((int)rint(m1 * 128.0), DD_STEPS);
    } else {
        s.clock_f = __fsub_rn(s.clock_f, 1.0f);
        const float mu = __fdiv_rn(s.clock_f, k.symbol_time_f);
        const float m1 = __fadd_rn(mu, 0.0078125f);
        const int ia = min((int)rintf(__fmul_rn(mu, 128.0f)), DD_STEPS);
        const int ib = min((int)rintf(__fmul_rn(m1, 128.0f)), DD_STEPS);
        imu = (1.0f < mu) ? DD_STEPS : ia;                  // min() returned the Python float 1.0
        imu1 = (1.0f < mu || 1.0f < m1) ? DD_STEPS : ib;
    }
    float w[DD_TAPS];
#pragma unroll
    for (int j = 0; j < DD_TAPS; ++j) w[j] = w8[j];
    float y = dd_interp(taps, w, imu);
    float y1 = dd_interp(taps, w, imu1);
    y = __fsub_rn(y, s.fine);
    y1 = __fsub_rn(y1, s.fine);
    const float sp = s.spread;
    const float soft = __fdiv_rn(__fmul_rn(2.0f, y), sp);
    const float c15 = (s.spread_py == 1) ? (float)(1.5 * 1.6) : (s.spread_py == 2) ? (float)(1.5 * 2.4) : __fmul_rn(1.5f, sp);
    const float c05 = __fmul_rn(0.5f, sp);
    const bool lo = y < -sp, neg = y < 0.0f, mid = y < sp;
    const float e_lo = __fadd_rn(y, c15), e_neg = __fadd_rn(y, c05), e_mid = __fsub_rn(y, c05), e_hi = __fsub_rn(y, c15);
    const float err = lo ? e_lo : neg ? e_neg : mid ? e_mid : e_hi;
    const float e01 = __fmul_rn(err, 0.01f);
    const float n_out = __fsub_rn(sp, __fmul_rn(__fmul_rn(err, 0.5f), 0.01f));
    const float n_neg = __fsub_rn(sp, e01), n_pos = __fadd_rn(sp, e01);
    const float ns = (lo || !mid) ? n_out : neg ? n_neg : n_pos;
    // max(1.6, min(2.4, spread)): the literals win when the float32 value reaches them
    const bool top = !(ns < 2.4f), bot = !(ns > 1.6f);
    s.spread = top ? 2.4f : bot ? 1.6f : ns;
    s.spread_py = top ? 2 : bot ? 1 : 0;
    const float cf = (PY && s.clock_py) ? (float)s.clock_d : s.clock_f;
    const float tadj = __fmul_rn(err, 0.025f);
    s.clock_f = (y1 < y) ? __fadd_rn(cf, tadj) : __fsub_rn(cf, tadj);
    s.clock_py = 0;
    s.coarse = __fadd_rn(s.coarse, __fmul_rn(__fsub_rn(s.fine, s.coarse), 0.00125f));
    s.fine = __fadd_rn(s.fine, __fmul_rn(err, 0.125f));
    if (count < max_sym) {
        dout[count] = (soft < -2.0f) ? 3 : (soft < 0.0f) ? 2 : (soft < 2.0f) ? 0 : 1;
        if (sout) sout[count] = soft;
    }
    ++count;
}

// One staged tile of one warp's 32 channels. The channels' symbol clocks are not aligned, so each lane walks its own row:
// it advances its clock (clock_run8: the reference's rounded float32 additions, eight per trip, no branch per sample) up
// to its next symbol instant or the end of the tile, then the warp runs the symbol update for every lane that reached one —
// once per symbol, not once per sample whenever any of the 32 channels ticks. The interpolator's 8-sample history is the
// staged row itself (8 carried samples in front of every tile): row[i .. i + 7] after i consumed samples.
template <bool PY>
__device__ __forceinline__ void dd_walk_tile(DDState& s, const DDConst& k, const float* __restrict__ taps, const float* __restrict__ row,
                                             int lim, bool live, unsigned char* __restrict__ dout, float* __restrict__ sout,
                                             int max_sym, int& count) {
    int i = 0;   // samples of this tile consumed by this lane
    while (true) {
        bool tick = false;
        if (live) {
            if (PY && s.clock_py) {
                while (i < lim) {
                    ++i;
                    s.clock_d += k.symbol_time;
                    if (s.clock_d > 1.0) {
                        tick = true;
                        break;
                    }
                }
            } else {
                float cf = s.clock_f;
                while (i < lim) {
                    i += clock_run<false>(cf, k.symbol_time_f, lim - i, tick);
                    if (tick) break;
                }
                s.clock_f = cf;
            }
        }
        if (!__any_sync(0xffffffffu, tick)) break;
        if (tick) dd_symbol<PY>(s, k, taps, row + i, dout, sout, max_sym, count);
    }
}

// _mmse_timing_recovery (p25.py:1250-1333): warp 0 walks (lane = channel), the helper warps stage the next tile.
__global__ void __launch_bounds__(DD_THREADS) dd_mmse_kernel(const float* __restrict__ x, int n, int C, DDConst k,
                                                             const float* __restrict__ taps_g, DDState* __restrict__ st,
                                                             unsigned char* __restrict__ dibits, float* __restrict__ soft_out,
                                                             int max_sym, int* __restrict__ n_sym) {
    __shared__ float tile[2][DD_CH][DD_ROWH];
    __shared__ float taps[(DD_STEPS + 1) * DD_TROW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (DD_STEPS + 1) * DD_TAPS; i += DD_THREADS) taps[(i >> 3) * DD_TROW + (i & 7)] = taps_g[i];
    const int c0 = blockIdx.x * DD_CH;
    const int c = c0 + lane;
    const bool live = warp == 0 && lane < DD_CH && c < C;
    const int rl = lane < DD_CH ? lane : 0;   // row of this lane (idle lanes alias row 0, read-only)
    DDState s;
    if (live) s = st[c];
    else memset(&s, 0, sizeof(s));
    const int hidx0 = s.hidx;
    // the ring, oldest first, in front of the first tile
    if (warp == 0 && lane < DD_CH)
        for (int j = 0; j < DD_PAD; ++j) tile[0][lane][j] = live ? st[c].hist[(hidx0 + j) & 7] : 0.f;
    int count = 0;
    unsigned char* dout = dibits + (long long)c * max_sym;
    float* sout = soft_out ? soft_out + (long long)c * max_sym : nullptr;
    const int T = (n + DD_TILE - 1) / DD_TILE;
    if (warp > 0) {
        dd_helper_load<DD_TILE, DD_ROWH, DD_PAD>(tile[0], x, n, c0, C, 0, n, warp - 1, lane);
        cp_async_wait<0>();
    }
    __syncthreads();
    for (int t = 0; t < T; ++t) {
        const int buf = t & 1;
        if (warp == 0) {
            const int lim = min(DD_TILE, n - t * DD_TILE);
            const float* row = tile[buf][rl];
            if (__any_sync(0xffffffffu, live && s.clock_py)) dd_walk_tile<true>(s, k, taps, row, lim, live, dout, sout, max_sym, count);
            else dd_walk_tile<false>(s, k, taps, row, lim, live, dout, sout, max_sym, count);
            // the last 8 samples (carried ones included when the tile is shorter) lead the next tile
#pragma unroll
            for (int j = 0; j < DD_PAD; ++j)
                if (lane < DD_CH) tile[buf ^ 1][lane][j] = row[lim + j];
        } else if (t + 1 < T) {
            dd_helper_load<DD_TILE, DD_ROWH, DD_PAD>(tile[buf ^ 1], x, n, c0, C, (t + 1) * DD_TILE, n, warp - 1, lane);
            cp_async_wait<0>();
        }
        __syncthreads();
    }
    if (live) {
        // buffer T & 1 holds, in its pad, the newest 8 samples, oldest first: back into ring order
        const int hend = (hidx0 + n) & 7;
        for (int j = 0; j < DD_PAD; ++j) st[c].hist[(hend + j) & 7] = tile[T & 1][lane][j];
        DDState* o = &st[c];
        o->clock_d = s.clock_d;
        o->clock_f = s.clock_f;
        o->clock_py = s.clock_py;
        o->spread = s.spread;
        o->spread_py = s.spread_py;
        o->fine = s.fine;
        o->coarse = s.coarse;
        o->hidx = hend;
        o->symbols = s.symbols + count;
        n_sym[c] = min(count, max_sym);
    }
}

__global__ void dd_reset_kernel(DDState* st, int C, int channel, int keep_gain) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C || (channel >= 0 && c != channel)) return;
    DDState& s = st[c];
    s.clock_d = 0.0;
    s.clock_f = 0.f;
    s.clock_py = 1;
    s.spread = 2.0f;
    s.spread_py = 0;
    s.fine = s.coarse = s.dc = 0.f;
    if (!keep_gain) s.gain = 1.0f;   // reset() leaves _input_gain alone (p25.py:1335-1345)
    for (int i = 0; i < DD_TAPS; ++i) s.hist[i] = 0.f;
    s.hidx = 0;
    s.symbols = 0;
}

// np.diff(np.unwrap([last | angle(iq)])): wrapped phase step per sample, float64 (trunking/system.py:708-717)
template <typename T2>
__global__ void fm_disc_kernel(const T2* __restrict__ iq, long long stride, int n, double* __restrict__ last, double* __restrict__ out,
                               double* __restrict__ new_last) {
    const int ch = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T2* x = iq + (long long)ch * stride;
    // np.angle of complex64 is float32 (correctly rounded here; numpy's SIMD arctan2 may be 1 ulp off), of complex128 float64
    const bool f32 = sizeof(x[0].x) == 4;
    double cur = atan2((double)x[i].y, (double)x[i].x);
    if (f32) cur = (double)(float)cur;
    double prev = last[ch];
    if (i > 0) {
        prev = atan2((double)x[i - 1].y, (double)x[i - 1].x);
        if (f32) prev = (double)(float)prev;
    }
    const double dd = cur - prev;
    // numpy.unwrap (period 2 pi, discont pi): ddmod = mod(dd + pi, 2 pi) - pi, with -pi -> +pi for positive steps;
    // the correction only applies where |dd| >= pi
    const double PI = 3.141592653589793, TWO_PI = 6.283185307179586;
    double r = fmod(dd + PI, TWO_PI);
    if (r < 0.0) r += TWO_PI;   // Python's mod has the sign of the divisor
    double ddmod = r - PI;
    if (ddmod == -PI && dd > 0.0) ddmod = PI;
    out[(long long)ch * n + i] = (fabs(dd) < PI) ? dd : ddmod;
    if (i == n - 1) new_last[ch] = cur;
}

// MMSE interpolation table of the reference (p25.py:1165-1186), float32
static void dd_build_taps(std::vector<float>& taps) {
    taps.assign((DD_STEPS + 1) * DD_TAPS, 0.f);
    for (int step = 0; step <= DD_STEPS; ++step) {
        const double mu = (double)step / DD_STEPS;
        for (int tap = 0; tap < DD_TAPS; ++tap) {
            const double t = tap - 3 - mu;
            float v;
            if (fabs(t) < 1e-6) v = 1.0f;
            else {
                const double sinc = sin(M_PI * t) / (M_PI * t);
                const double win = (fabs(t) < 4) ? 0.5 * (1 + cos(M_PI * t / 4)) : 0.0;
                v = (float)(sinc * win);
            }
            taps[step * DD_TAPS + tap] = v;
        }
        // np.sum of 8 float32 values (pairwise == sequential below 8 elements is NOT guaranteed; numpy adds the 8-element
        // row with its unrolled pairwise kernel: ((a0+a1)+(a2+a3)) + ((a4+a5)+(a6+a7)))
        const float* r = &taps[step * DD_TAPS];
        const float tot = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        if (tot > 0)
            for (int tap = 0; tap < DD_TAPS; ++tap) taps[step * DD_TAPS + tap] /= tot;
    }
}

}  // namespace wc

using namespace wc;

struct wc_discdemod {
    int C = 0;
    int sample_rate = 0, symbol_rate = 0;
    DDConst k;
    DDState* d_state = nullptr;
    float* d_taps = nullptr;
    double* d_lpf = nullptr;
    float* d_peak = nullptr;
    float* d_a = nullptr;  size_t a_cap = 0;   // [C][n] after DC removal
    float* d_b = nullptr;  size_t b_cap = 0;   // [C][n] after the low-pass
    // host-call staging
    float* d_in = nullptr; size_t in_cap = 0;
    unsigned char* d_dib = nullptr; size_t dib_cap = 0;
    float* d_soft = nullptr; size_t soft_cap = 0;
    int* d_nsym = nullptr;
    cudaStream_t stream = nullptr;
};

template <typename T>
static int dd_ensure(T** p, size_t* cap, size_t need) {
    if (*cap >= need) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    WC_CUDA(cudaMalloc((void**)p, need * sizeof(T)));
    *cap = need;
    return 0;
}

extern "C" {

int wc_discdemod_create(int n_channels, int sample_rate, int symbol_rate, const float* mmse_taps_129x8,
                        const float* lpf_taps65, wc_discdemod** out) {
    WC_REQUIRE(out != nullptr, "wc_discdemod_create: out is null");
    WC_REQUIRE(n_channels >= 1 && sample_rate > 0 && symbol_rate > 0 && symbol_rate < sample_rate,
               "wc_discdemod_create: bad parameters");
    WC_REQUIRE(lpf_taps65 != nullptr, "wc_discdemod_create: the 65-tap low-pass design is required (scipy firwin on the host)");
    wc_discdemod* h = new wc_discdemod();
    h->C = n_channels;
    h->sample_rate = sample_rate;
    h->symbol_rate = symbol_rate;
    h->k.symbol_time = (double)symbol_rate / (double)sample_rate;
    h->k.symbol_time_f = (float)h->k.symbol_time;
    std::vector<float> taps;
    if (mmse_taps_129x8) taps.assign(mmse_taps_129x8, mmse_taps_129x8 + (DD_STEPS + 1) * DD_TAPS);
    else dd_build_taps(taps);
    std::vector<double> lpf(DD_LPF);
    for (int i = 0; i < DD_LPF; ++i) lpf[i] = (double)lpf_taps65[i];
    const size_t C = (size_t)n_channels;
    bool ok = cudaMalloc(&h->d_state, sizeof(DDState) * C) == cudaSuccess &&
              cudaMalloc(&h->d_taps, sizeof(float) * taps.size()) == cudaSuccess &&
              cudaMalloc(&h->d_lpf, sizeof(double) * DD_LPF) == cudaSuccess &&
              cudaMalloc(&h->d_peak, sizeof(float) * C) == cudaSuccess &&
              cudaMalloc(&h->d_nsym, sizeof(int) * C) == cudaSuccess &&
              cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMemcpy(h->d_taps, taps.data(), sizeof(float) * taps.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(h->d_lpf, lpf.data(), sizeof(double) * DD_LPF, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        set_error("wc_discdemod_create: CUDA allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete h;
        return -2;
    }
    cudaMemset(h->d_state, 0, sizeof(DDState) * C);
    dd_reset_kernel<<<(n_channels + 127) / 128, 128, 0, h->stream>>>(h->d_state, n_channels, -1, 0);
    cudaStreamSynchronize(h->stream);
    *out = h;
    return 0;
}

void wc_discdemod_destroy(wc_discdemod* h) {
    if (!h) return;
    cudaFree(h->d_state);
    cudaFree(h->d_taps);
    cudaFree(h->d_lpf);
    cudaFree(h->d_peak);
    cudaFree(h->d_nsym);
    if (h->d_a) cudaFree(h->d_a);
    if (h->d_b) cudaFree(h->d_b);
    if (h->d_in) cudaFree(h->d_in);
    if (h->d_dib) cudaFree(h->d_dib);
    if (h->d_soft) cudaFree(h->d_soft);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int wc_discdemod_reset(wc_discdemod* h, int channel) {
    WC_REQUIRE(h != nullptr, "wc_discdemod_reset: null handle");
    WC_REQUIRE(channel >= -1 && channel < h->C, "wc_discdemod_reset: channel %d out of range", channel);
    dd_reset_kernel<<<(h->C + 127) / 128, 128, 0, h->stream>>>(h->d_state, h->C, channel, 1);
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int wc_discdemod_max_symbols(const wc_discdemod* h, int n_samples) {
    if (!h || n_samples <= 0) return 0;
    // the timing loop moves the clock by at most 0.025 * 3.6 per symbol: 1.2x the nominal rate is a safe bound
    return (int)(1.2 * (double)n_samples * h->k.symbol_time) + 8;
}

int wc_discdemod_get_taps(const wc_discdemod* h, float* mmse_taps_129x8) {
    WC_REQUIRE(h && mmse_taps_129x8, "wc_discdemod_get_taps: null argument");
    WC_CUDA(cudaMemcpy(mmse_taps_129x8, h->d_taps, sizeof(float) * (DD_STEPS + 1) * DD_TAPS, cudaMemcpyDeviceToHost));
    return 0;
}

int wc_discdemod_demod(wc_discdemod* h, const float* audio_dev, long long chan_stride, int n_samples,
                       unsigned char* dibits_dev, float* soft_dev, int* n_sym_dev, int max_sym, void* stream_v) {
    WC_REQUIRE(h && audio_dev && dibits_dev && n_sym_dev, "wc_discdemod_demod: null argument");
    WC_REQUIRE(n_samples >= 0 && chan_stride >= n_samples, "wc_discdemod_demod: bad sizes");
    cudaStream_t s = (cudaStream_t)stream_v;
    const int C = h->C;
    if (n_samples == 0) {
        WC_CUDA(cudaMemsetAsync(n_sym_dev, 0, sizeof(int) * C, s));
        return 0;
    }
    WC_REQUIRE(max_sym >= wc_discdemod_max_symbols(h, n_samples), "wc_discdemod_demod: max_sym %d < %d", max_sym,
               wc_discdemod_max_symbols(h, n_samples));
    if (dd_ensure(&h->d_a, &h->a_cap, (size_t)C * n_samples)) return -2;
    if (dd_ensure(&h->d_b, &h->b_cap, (size_t)C * n_samples)) return -2;
    dd_peak_kernel<<<C, 256, 0, s>>>(audio_dev, chan_stride, n_samples, h->d_peak);
    dd_dc_kernel<<<(C + DD_CH - 1) / DD_CH, DD_THREADS, 0, s>>>(audio_dev, chan_stride, n_samples, C, h->d_peak, h->d_state, h->d_a);
    dim3 lg((n_samples + 255) / 256, C);
    dd_lpf_kernel<<<lg, 256, 0, s>>>(h->d_a, n_samples, h->d_lpf, h->d_b);
    dd_mmse_kernel<<<(C + DD_CH - 1) / DD_CH, DD_THREADS, 0, s>>>(h->d_b, n_samples, C, h->k, h->d_taps, h->d_state, dibits_dev, soft_dev,
                                                max_sym, n_sym_dev);
    WC_CUDA(cudaGetLastError());
    return 0;
}

int wc_discdemod_demod_host(wc_discdemod* h, const float* audio_host, int n_samples, unsigned char* dibits_host,
                            float* soft_host, int* n_sym_host, int max_sym) {
    WC_REQUIRE(h && audio_host && dibits_host && n_sym_host, "wc_discdemod_demod_host: null argument");
    const int C = h->C;
    if (n_samples <= 0) {
        for (int c = 0; c < C; ++c) n_sym_host[c] = 0;
        return 0;
    }
    if (dd_ensure(&h->d_in, &h->in_cap, (size_t)C * n_samples)) return -2;
    if (dd_ensure(&h->d_dib, &h->dib_cap, (size_t)C * max_sym)) return -2;
    if (dd_ensure(&h->d_soft, &h->soft_cap, (size_t)C * max_sym)) return -2;
    WC_CUDA(cudaMemcpyAsync(h->d_in, audio_host, sizeof(float) * (size_t)C * n_samples, cudaMemcpyHostToDevice, h->stream));
    int rc = wc_discdemod_demod(h, h->d_in, n_samples, n_samples, h->d_dib, h->d_soft, h->d_nsym, max_sym, h->stream);
    if (rc) return rc;
    WC_CUDA(cudaMemcpyAsync(dibits_host, h->d_dib, (size_t)C * max_sym, cudaMemcpyDeviceToHost, h->stream));
    if (soft_host)
        WC_CUDA(cudaMemcpyAsync(soft_host, h->d_soft, sizeof(float) * (size_t)C * max_sym, cudaMemcpyDeviceToHost, h->stream));
    WC_CUDA(cudaMemcpyAsync(n_sym_host, h->d_nsym, sizeof(int) * C, cudaMemcpyDeviceToHost, h->stream));
    WC_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

/* state8 = {input gain, dc estimate, symbol clock, symbol spread, fine freq correction, coarse, clock is still the
 * Python float (1/0), symbols produced since reset} */
int wc_discdemod_get_state(wc_discdemod* h, int channel, double* state8) {
    WC_REQUIRE(h && state8, "wc_discdemod_get_state: null argument");
    WC_REQUIRE(channel >= 0 && channel < h->C, "wc_discdemod_get_state: channel %d out of range", channel);
    DDState s;
    WC_CUDA(cudaStreamSynchronize(h->stream));
    WC_CUDA(cudaMemcpy(&s, h->d_state + channel, sizeof(DDState), cudaMemcpyDeviceToHost));
    state8[0] = s.gain;
    state8[1] = s.dc;
    state8[2] = s.clock_py ? s.clock_d : (double)s.clock_f;
    state8[3] = (s.spread_py == 1) ? 1.6 : (s.spread_py == 2) ? 2.4 : (double)s.spread;
    state8[4] = s.fine;
    state8[5] = s.coarse;
    state8[6] = s.clock_py;
    state8[7] = (double)s.symbols;
    return 0;
}

/* trunking/system.py:708-717. iq complex64 (is_f64 = 0) or complex128 (1) [C][chan_stride]; last_phase_dev float64 [C] in/out
 * (0.0 initially); out float64 [C][n_samples]. */
int wc_fm_discriminator(const void* iq_dev, int is_f64, long long chan_stride, int n_samples, int n_channels,
                        double* last_phase_dev, double* out_dev, void* stream_v) {
    WC_REQUIRE(iq_dev && last_phase_dev && out_dev, "wc_fm_discriminator: null argument");
    WC_REQUIRE(n_channels >= 1 && n_samples >= 0 && chan_stride >= n_samples, "wc_fm_discriminator: bad sizes");
    if (n_samples == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream_v;
    double* tmp = nullptr;
    WC_CUDA(cudaMallocAsync((void**)&tmp, sizeof(double) * n_channels, s));
    dim3 g((n_samples + 255) / 256, n_channels);
    if (is_f64) fm_disc_kernel<double2><<<g, 256, 0, s>>>((const double2*)iq_dev, chan_stride, n_samples, last_phase_dev, out_dev, tmp);
    else fm_disc_kernel<float2><<<g, 256, 0, s>>>((const float2*)iq_dev, chan_stride, n_samples, last_phase_dev, out_dev, tmp);
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaMemcpyAsync(last_phase_dev, tmp, sizeof(double) * n_channels, cudaMemcpyDeviceToDevice, s));
    WC_CUDA(cudaFreeAsync(tmp, s));
    return 0;
}

}  // extern "C"
