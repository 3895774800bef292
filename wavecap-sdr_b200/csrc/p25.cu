// P25 Phase-1 C4FM symbol recovery, batched over channels (config C4).
//
// Replaces wavecapsdr/dsp/p25/c4fm.py:2379-2807 (C4FMDemodulator.demodulate and its helpers) for C
// independent, stateful channels per call:
//
//   K1 p25_fir_kernel    I/Q low-pass + RRC (c4fm.py:2570-2588: four scipy lfilter(b,1,x,zi) calls in
//                        float64) as ONE float64 FIR with the convolved taps and carried input history.
//   K2 c4fm_phase_kernel symbol-spaced differential demodulator (_FMDemodulator.demodulate, :324-395):
//                        8-tap fractional interpolation (f32 products, f64 sum, f32 result), conjugate
//                        product in f32, atan2 -> phase; phases go to a scratch row and to the
//                        channel's 65536-entry phase ring (the reference's self._buffer, :2492).
//   K3 c4fm_sync_kernel  one warp per channel: fixed-rate symbol extraction (_symbol_recovery_jit,
//                        :649-783), then the sync loop (:2596-2807): primary + lagging soft sync
//                        correlators (:2268-2321), threshold crossing, hill-climb timing optimiser
//                        (_timing_optimize_jit :543-644), PLL/gain correction (:260-272) and re-slicing
//                        of the following <=340 symbols (_resample_message_jit :795-869).
//
// Buffer model. The reference shifts its 65536-float buffer down by half whenever the write pointer
// reaches the end (:709-725). Here the buffer is a ring holding the last 65536 phases; virtual index v
// of the reference buffer maps to ring[(v + shift_mod) & 65535] and reads beyond the write pointer
// return 0 (the reference zeroes the upper half on a shift). No data moves.
//
// Arithmetic follows the reference AS EXECUTED under NumPy 2 / numba: every value
// the reference keeps in float32 is float32 here, every Python-float / numba-f64 expression is float64
// with the same operation order, and the file is compiled with -fmad=false so no multiply-add is
// contracted where the reference rounds twice.
#include <math.h>
#include <string.h>
#include <vector>

#include "../../include/wcsdr_b200.h"
#include "common.cuh"

namespace wc {

__constant__ float c_interp[129][8] = {
#include "interp_taps_129x8.inc"
};
// +-3 per dibit of the 48-bit sync word 0x5575F5FF77FF (c4fm.py:2279-2299)
__constant__ float c_sync[24];

constexpr int C4_RING = 65536;
constexpr int C4_HALF = 32768;
constexpr double C4_HALF_PI = 1.5707963267948966;
constexpr double C4_NORM = 1.2732395447351628;      // 4/pi
constexpr double C4_LOOP_GAIN = 0.15;               // c4fm.py:63
constexpr double C4_MAX_PLL = 1.0471975511965976;   // pi/3, :64
constexpr double C4_MAX_GAIN = 1.25;                // :65
constexpr double C4_INITIAL_GAIN = 1.219;           // :66
constexpr double C4_THRESH = 100.0;                 // :2408-2409
constexpr int C4_MSG_DIBITS = 340;                  // :792

struct C4State {
    double sample_point, pll, gain;
    int ptr;          // reference _buffer_pointer
    int shift_mod;    // (number of half shifts * 32768) mod 65536
    int eq_init, fine, since_sync, sync_count;
    float prev_phase; // phase of the last sample of the previous call (buffer[ptr])
    int n_events;     // sync events accepted during the last call (diagnostics)
    float det[24];    // primary soft-sync ring, oldest first
    float lag[24];    // lagging soft-sync ring, oldest first
};

struct C4Const {
    double sps;
    double lag_offset;   // sps / 2
    double lag_mu;       // 1 - frac(lag_offset)
    double max_fine_adj; // 0.2 * sps
    int lag_int;         // int(lag_offset)
    int ov;              // floor(sps) + 4
    int interp_off;      // max(0, floor(sps) - 4)
    int row;             // interpolator row of the differential demodulator
};

// ---------------------------------------------------------------------------------------------
// K1: streaming FIR, float64 accumulation, real taps on complex input
// ---------------------------------------------------------------------------------------------
constexpr int FIR_THREADS = 128;
constexpr int FIR_PER_THREAD = 8;
constexpr int FIR_TILE = FIR_THREADS * FIR_PER_THREAD;  // 1024 outputs per CTA
constexpr int FIR_MAX_TAPS = 512;

struct FirArgs {
    const float2* x;        // [C][stride]
    long long stride;
    int n;                  // samples per channel in this call
    const float2* hist;     // [C][ntp-1] previous inputs, oldest first
    const double* taps;     // [ntp], taps[t] multiplies x[n-t]; zero padded to a multiple of 8
    int ntp;
    float2* y;              // [C][n]
};

__global__ void __launch_bounds__(FIR_THREADS) p25_fir_kernel(const FirArgs a) {
    // plane-transposed tile: logical sample i lives at plane (i & 7), slot (i >> 3) so that the
    // sliding-window loads of a warp (stride 8 samples between lanes) are conflict free
    constexpr int P = (FIR_TILE + FIR_MAX_TAPS) / 8 + 1;
    __shared__ double2 xs[8 * P];
    __shared__ double hs[FIR_MAX_TAPS];
    const int ch = blockIdx.y;
    const int n0 = blockIdx.x * FIR_TILE;
    const int hl = a.ntp - 1;
    const float2* xc = a.x + (long long)ch * a.stride;
    const float2* hc = a.hist + (long long)ch * hl;
    const int total = FIR_TILE + hl;
    for (int i = threadIdx.x; i < total; i += FIR_THREADS) {
        const int g = n0 - hl + i;  // chunk-relative sample index
        float2 v = make_float2(0.f, 0.f);
        if (g < 0) {
            if (hl + g >= 0) v = hc[hl + g];
        } else if (g < a.n) {
            v = xc[g];
        }
        xs[(i & 7) * P + (i >> 3)] = make_double2((double)v.x, (double)v.y);
    }
    for (int i = threadIdx.x; i < a.ntp; i += FIR_THREADS) hs[i] = a.taps[i];
    __syncthreads();

    const int tid = threadIdx.x;
    double2 w[8];
    double accr[8], acci[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        w[r] = xs[r * P + tid];  // logical 8*tid + r
        accr[r] = 0.0;
        acci[r] = 0.0;
    }
    // step s (oldest tap first): output r uses xs[8*tid + s + r] * taps[hl - s]
    for (int s8 = 0; s8 < a.ntp; s8 += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int s = s8 + u;
            const double h = hs[hl - s];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const double2 v = w[(u + r) & 7];
                accr[r] = fma(h, v.x, accr[r]);
                acci[r] = fma(h, v.y, acci[r]);
            }
            w[u] = xs[u * P + tid + 1 + (s8 >> 3)];  // logical 8*tid + s + 8
        }
    }
    float2* yc = a.y + (long long)ch * a.n;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int g = n0 + 8 * tid + r;
        if (g < a.n) yc[g] = make_float2((float)accr[r], (float)acci[r]);
    }
}

// history update: new_hist = last (ntp-1) samples of (old_hist ++ x[0..n))
__global__ void p25_hist_kernel(const float2* x, long long stride, int n, const float2* old_h, float2* new_h, int hl) {
    const int ch = blockIdx.x;
    for (int i = threadIdx.x; i < hl; i += blockDim.x) {
        const int g = n - hl + i;
        new_h[(long long)ch * hl + i] = (g >= 0) ? x[(long long)ch * stride + g] : old_h[(long long)ch * hl + (hl + g)];
    }
}

// ---------------------------------------------------------------------------------------------
// K2: differential demodulator -> phases
// ---------------------------------------------------------------------------------------------
struct PhaseArgs {
    const float2* filt;   // [C][n] filtered I/Q (float32, c4fm.py:2593)
    const float2* tail;   // [C][ov] last `ov` filtered samples of the previous call (zeros after reset)
    float2* new_tail;     // [C][ov]
    int n;
    C4Const k;
    const C4State* st;
    float* ph;            // [C][n] scratch
    float* ring;          // [C][65536]
};

__device__ __forceinline__ float2 c4_bufget(const PhaseArgs& a, int ch, int j) {
    return (j < a.k.ov) ? a.tail[(long long)ch * a.k.ov + j] : a.filt[(long long)ch * a.n + (j - a.k.ov)];
}

__global__ void __launch_bounds__(256) c4fm_phase_kernel(const PhaseArgs a) {
    const int ch = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < a.k.ov) {
        // new overlap = last ov entries of the concatenated buffer [tail | filt] (length ov + n)
        a.new_tail[(long long)ch * a.k.ov + x] = c4_bufget(a, ch, a.n + x);
    }
    if (x >= a.n) return;
    const float2 prev = c4_bufget(a, ch, x);
    double si = 0.0, sq = 0.0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const float2 v = c4_bufget(a, ch, a.k.interp_off + x + t);
        const float tap = c_interp[a.k.row][t];
        si += (double)__fmul_rn(v.x, tap);   // float32 product, float64 running sum (c4fm.py:401-409)
        sq += (double)__fmul_rn(v.y, tap);
    }
    const float ic = (float)si, qc = (float)sq;   // NumPy-2: f64 meets np.float32 operands -> float32
    const float ip = prev.x, qpc = -prev.y;
    const float di = __fsub_rn(__fmul_rn(ip, ic), __fmul_rn(qpc, qc));
    const float dq = __fadd_rn(__fmul_rn(ip, qc), __fmul_rn(ic, qpc));
    const float phase = (float)atan2((double)dq, (double)di);
    a.ph[(long long)ch * a.n + x] = phase;
    if (x >= a.n - C4_RING) {
        const C4State& s = a.st[ch];
        const int slot = (s.ptr + s.shift_mod + 1 + x) & (C4_RING - 1);
        a.ring[(long long)ch * C4_RING + slot] = phase;
    }
}

// ---------------------------------------------------------------------------------------------
// K2': discriminator-audio front end of demodulate_discriminator (c4fm.py:2855-2874): RRC FIR in float64 with the
// state scipy's lfilter carries (first call: zi = lfilter_zi(rrc) * audio[0], i.e. the filter behaves as if audio[0]
// had been its input forever), phases = filtered * samples_per_symbol -> float32, into the scratch row and the ring.
// ---------------------------------------------------------------------------------------------
struct DiscArgs {
    const float* x;        // [C][stride] discriminator audio (float32, as the reference casts it)
    long long stride;
    int n;
    const double* hist;    // [C][hl] previous inputs, oldest first
    double* new_hist;
    const double* taps;    // [nt] float32 design promoted to float64
    int nt;
    int* init;             // [C] 0 until the first call
    const double* first;   // [C] audio[0] of the first call (float64: the caller's own dtype)
    double sps;
    const C4State* st;
    float* ph;             // [C][n]
    float* ring;           // [C][65536]
};

__global__ void __launch_bounds__(256) disc_fir_kernel(const DiscArgs a) {
    const int ch = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.n) return;
    const int hl = a.nt - 1;
    const float* xc = a.x + (long long)ch * a.stride;
    const double* hc = a.hist + (long long)ch * hl;
    const bool warm = a.init[ch] != 0;
    const double x0 = a.first[ch];
    double acc = 0.0;
    for (int t = 0; t < a.nt; ++t) {
        const int j = x - t;
        const double v = (j >= 0) ? (double)xc[j] : (warm ? hc[hl + j] : x0);
        acc = fma(a.taps[t], v, acc);
    }
    const float phase = (float)(acc * a.sps);
    a.ph[(long long)ch * a.n + x] = phase;
    if (x >= a.n - C4_RING) {
        const C4State& s = a.st[ch];
        const int slot = (s.ptr + s.shift_mod + 1 + x) & (C4_RING - 1);
        a.ring[(long long)ch * C4_RING + slot] = phase;
    }
}

// runs after disc_fir_kernel on the same stream
__global__ void disc_hist_kernel(const DiscArgs a) {
    const int ch = blockIdx.x;
    const int hl = a.nt - 1;
    const bool warm = a.init[ch] != 0;
    for (int i = threadIdx.x; i < hl; i += blockDim.x) {
        const int g = a.n - hl + i;
        double v;
        if (g >= 0) v = (double)a.x[(long long)ch * a.stride + g];
        else v = warm ? a.hist[(long long)ch * hl + (hl + g)] : a.first[ch];
        a.new_hist[(long long)ch * hl + i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) a.init[ch] = 1;
}

// ---------------------------------------------------------------------------------------------
// K3: symbol extraction + sync loop, one warp per channel
// ---------------------------------------------------------------------------------------------
struct SyncArgs {
    int n;                 // samples per channel in this call
    int max_sym;           // row length of the outputs
    C4Const k;
    C4State* st;
    const float* ph;       // [C][n]
    const float* ring;     // [C][65536]
    unsigned char* dibits; // [C][max_sym]
    float* soft;           // [C][max_sym]
    int* idx;              // [C][max_sym] scratch: reference buffer index of each symbol, -1 if shifted out
    int* n_sym;            // [C]
};

struct SyncCtx {
    const float* ring;
    const float* taps;    // shared-memory copy of the 129 x 8 interpolator table, rows padded to 9 floats (the lanes of
                          // a score evaluation index 24 different rows: that would serialise on the constant cache)
    const float* win;     // shared-memory copy of the buffer values win_lo .. win_lo + win_n - 1 around the sync word
    int win_lo, win_n;    // being timed (one fill per event instead of eight L2 round trips per score); 0 = none
    int ptr, shift_mod;
    double sps, pll, gain;
    float my_sync;        // c_sync[lane]
};
constexpr int C4_ROW = 9;
constexpr int C4_WIN = 1024;        // floats; a sync word spans 24 * sps samples
constexpr int C4_SOFTWIN = 256;     // symbols of the soft / index window the sync loop reads from shared memory

__device__ __forceinline__ float c4_buf(const SyncCtx& c, int v) {
    if (v < 0 || v > c.ptr) return 0.0f;
    return c.ring[(v + c.shift_mod) & (C4_RING - 1)];
}

// the same value through the per-event window when it covers v
__device__ __forceinline__ float c4_get(const SyncCtx& c, int v) {
    const unsigned u = (unsigned)(v - c.win_lo);
    if (u < (unsigned)c.win_n) return c.win[u];
    return c4_buf(c, v);
}

__device__ __forceinline__ int c4_slice(double sr) {
    if (sr >= C4_HALF_PI) return 1;
    if (sr >= 0.0) return 0;
    if (sr >= -C4_HALF_PI) return 2;
    return 3;
}

__device__ __forceinline__ double shfl_d(double v, int src) {
    return __shfl_sync(0xffffffffu, v, src);
}

// per-lane interpolated soft value of sync symbol `lane` (< 24) for the pointer sequence starting at
// offset - 23*sps (c4fm.py:416-459). Returns validity.
__device__ __forceinline__ bool c4_sync_sample(const SyncCtx& c, double offset, int lane, double& soft) {
    // the reference accumulates ptr += sps: every lane runs the 23-term chain (unrolled, no branch) and keeps its own term
    double q = offset - (23.0 * c.sps);
    double p = q;
#pragma unroll
    for (int t = 1; t < 24; ++t) {
        q += c.sps;
        p = (lane >= t) ? q : p;
    }
    const int bi = (int)p;
    const int io = bi - 3;
    if (lane >= 24 || io < 0 || io > C4_RING - 8) return false;
    int row = (int)((1.0 - (p - (double)bi)) * 128.0 + 0.5);
    row = min(max(row, 0), 128);
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = c4_get(c, io + j);
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += (double)__fmul_rn(x[j], c.taps[row * C4_ROW + j]);
    soft = (acc + c.pll) * c.gain;
    return true;
}

// _timing_score_jit at NS offsets at once: per offset the sum over the 24 sync symbols in order i = 0..23. The NS
// evaluations are independent, so their chains interleave, and lane q adds up offset q's terms (results broadcast).
template <int NS>
__device__ __forceinline__ void c4_scores(const SyncCtx& c, const double (&offs)[NS], int lane, double* sterm, double (&out)[NS]) {
    unsigned vm[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) {
        double soft = 0.0;
        const bool ok = c4_sync_sample(c, offs[q], lane, soft);
        vm[q] = __ballot_sync(0xffffffffu, ok);
        if (lane < 24) sterm[q * 24 + lane] = ok ? soft * (double)c.my_sync : 0.0;
    }
    __syncwarp();
    double score = 0.0;
    if (lane < NS) {
        unsigned v = vm[0];
#pragma unroll
        for (int q = 1; q < NS; ++q) v = (lane == q) ? vm[q] : v;
        const double* t = sterm + lane * 24;
        if ((v & 0xffffffu) == 0xffffffu) {  // the usual case: all 24 samples inside the buffer — same order, loads up front
#pragma unroll
            for (int i = 0; i < 24; ++i) score += t[i];
        } else {
            for (int i = 0; i < 24; ++i)
                if (v & (1u << i)) score += t[i];
        }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < NS; ++q) out[q] = shfl_d(score, q);
}

__device__ __forceinline__ double c4_score(const SyncCtx& c, double offset, int lane, double* sterm) {
    const double o[1] = {offset};
    double r[1];
    c4_scores<1>(c, o, lane, sterm, r);
    return r[0];
}

// _timing_correction_jit (c4fm.py:462-540)
__device__ void c4_correction(const SyncCtx& c, double offset, int lane, double* sterm, double& pll_adj, double& gain_acc) {
    double soft = 0.0;
    const bool ok = c4_sync_sample(c, offset, lane, soft);
    const unsigned vm = __ballot_sync(0xffffffffu, ok);
    if (lane < 24) sterm[lane] = soft;
    __syncwarp();
    double pa = 0.0, ga = 0.0;
    if (lane == 0) {
        double bp = 0.0, bm = 0.0;
        int pc = 0, mc = 0;
        for (int i = 0; i < 24; ++i) {
            if (!(vm & (1u << i))) continue;
            const double s = sterm[i];
            const double ideal = (double)c_sync[i];
            if (ideal > 0.0) {
                bp += s - ideal;
                ++pc;
            } else {
                bm += s - ideal;
                ++mc;
            }
            ga += fabs(ideal) - fabs(s);
        }
        if (pc > 0) bp /= (double)(-pc);
        if (mc > 0) bm /= (double)(-mc);
        pa = (bp + bm) / 2.0;
        pa = fmin(fmax(pa, -C4_HALF_PI), C4_HALF_PI);
        ga = ga / (24.0 * 2.356194490192345);
    }
    __syncwarp();
    pll_adj = shfl_d(pa, 0);
    gain_acc = shfl_d(ga, 0);
}

// DISC = true: the sync loop of C4FMDemodulator.demodulate_discriminator (c4fm.py:2896-2966) instead of demodulate's
// (:2596-2807): same detectors and lagging path, but an accepted sync only re-times the sample point — no PLL / gain
// correction, no re-slicing, no sync counter — and the optimiser is handed the sample point as its buffer offset.
// Two warps per channel: warp 1 extracts the symbols (a float64 recurrence on one lane + 32 interpolations per block) and
// publishes how far it is; warp 0 runs the sync loop a window behind it. The two halves were 22 % / 78 % of a serial kernel.
constexpr int C4_THREADS = 64;
constexpr int C4_ADJ = 256;   // sample-point corrections warp 0 may hold back until warp 1 has delivered the final sample point
template <bool DISC>
__global__ void __launch_bounds__(C4_THREADS) c4fm_sync_kernel(const SyncArgs a) {
    __shared__ volatile int s_prod, s_done, s_nsym;
    __shared__ volatile double s_sp;
    __shared__ double s_adj[C4_ADJ];
    __shared__ double sterm[3 * 24];
    __shared__ float lagbuf[24 + 32];
    __shared__ float s_soft[23 + C4_SOFTWIN];   // soft[kw - 23 .. kw + 255] of the sync loop's current window
    __shared__ int s_idx[C4_SOFTWIN];
    __shared__ float s_win[C4_WIN];
    __shared__ int pos_n[2][32];
    __shared__ double pos_mu[2][32];
    __shared__ float s_taps[129 * C4_ROW];
    for (int i = threadIdx.x; i < 129 * 8; i += C4_THREADS) s_taps[(i >> 3) * C4_ROW + (i & 7)] = c_interp[i >> 3][i & 7];
    if (threadIdx.x == 0) {
        s_prod = 0;
        s_done = 0;
    }
    __syncthreads();
    const int ch = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    C4State& S = a.st[ch];
    const float* ph = a.ph + (long long)ch * a.n;
    unsigned char* dib = a.dibits + (long long)ch * a.max_sym;
    volatile float* soft = a.soft + (long long)ch * a.max_sym;
    int* idx = a.idx + (long long)ch * a.max_sym;
    const double sps = a.k.sps;

    // ---- chunk geometry: shifts the reference performs while writing n samples (c4fm.py:709-725)
    const int ptr0 = S.ptr;
    const long long endp = (long long)ptr0 + a.n;
    const int n_shift = (endp < C4_RING - 1) ? 0 : 1 + (int)((endp - (C4_RING - 1)) / C4_HALF);
    const long long idx_base = (long long)ptr0 + 1 - (long long)n_shift * C4_HALF;

    // ---- symbol extraction at the start-of-call pll/gain (c4fm.py:649-783)
    const double pll0 = S.pll, gain0 = S.gain;
    const float prev_phase = S.prev_phase;
    if (warp == 1) {
        int nsym = 0;
        // Positions are a sequential float64 recurrence (lane 0); the 32 interpolations of a block are parallel. The two
        // are software-pipelined: while the block's phase loads are in flight, lane 0 works out the next block's positions.
        double sp = S.sample_point;
        int nn = -1;  // chunk-relative index of the last consumed sample
        bool done = (a.n == 0);
        auto positions = [&](int* pn, double* pm) -> int {   // lane 0: up to 32 symbol positions; sets `done` at the end
            int cnt = 0;
            while (cnt < 32) {
                const int left = a.n - 1 - nn;
                // eight symbols in straight-line code while the end of the chunk is out of reach: each takes
                // floor(sp) <= sp < sps + 1 samples. floor() by a round-down addition of 2^52 (sp < 2^31 here), the
                // subtractions are exact, (int)f == the reference's int(floor(sp)).
                if (cnt <= 24 && sp >= 1.0 && sp < 1.0e9 && sps >= 1.0 && sps < 1.0e6 && (double)left > sp + 8.0 * (sps + 1.0)) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const double f = __dadd_rd(sp, 4503599627370496.0) - 4503599627370496.0;
                        sp = sp - f;
                        nn += (int)f;
                        pn[cnt + q] = nn;
                        pm[cnt + q] = 1.0 - sp;
                        sp += sps;
                    }
                    cnt += 8;
                    continue;
                }
                if (left <= 0) {
                    done = true;
                    break;
                }
                // number of `sp -= 1.0` steps until sp < 1.0 (each step is exact for sp >= 1)
                int j = 1;
                if (sp >= 1.0) {
                    const double f = floor(sp);
                    j = (f > 2.0e9) ? 2000000000 : (int)f;
                }
                if (j > left) {
                    sp = sp - (double)left;
                    nn += left;
                    done = true;
                    break;
                }
                sp = sp - (double)j;
                nn += j;
                pn[cnt] = nn;
                pm[cnt] = 1.0 - sp;
                ++cnt;
                sp += sps;
            }
            return cnt;
        };
        int cur = 0, cnt = 0;
        if (!done && lane == 0) cnt = positions(pos_n[0], pos_mu[0]);
        cnt = __shfl_sync(0xffffffffu, cnt, 0);
        done = __shfl_sync(0xffffffffu, (int)done, 0) != 0;
        __syncwarp();
        while (cnt > 0) {
            const bool mine = lane < cnt && nsym + lane < a.max_sym;
            int m = 0;
            double mu = 0.0;
            float x1 = 0.f, x2 = 0.f;
            if (mine) {
                m = pos_n[cur][lane];
                mu = pos_mu[cur][lane];
                x2 = ph[m];
                x1 = (m > 0) ? ph[m - 1] : prev_phase;
            }
            int ncnt = 0;
            if (!done && lane == 0) ncnt = positions(pos_n[cur ^ 1], pos_mu[cur ^ 1]);
            if (mine) {
                double v;
                if (mu < 0.0) v = (double)x1;
                else if (mu > 1.0) v = (double)x2;
                else v = (double)x1 + (double)__fsub_rn(x2, x1) * mu;
                const double sr = (v + pll0) * gain0;
                dib[nsym + lane] = (unsigned char)c4_slice(sr);
                soft[nsym + lane] = (float)(sr * C4_NORM);
                const long long vi = idx_base + m;
                idx[nsym + lane] = (vi >= 0) ? (int)vi : -1;
            }
            nsym = min(nsym + cnt, a.max_sym);
            cnt = __shfl_sync(0xffffffffu, ncnt, 0);
            done = __shfl_sync(0xffffffffu, (int)done, 0) != 0;
            cur ^= 1;
            __threadfence();         // this block's dibits / soft values / indices (read back through L2) before the count
            __syncwarp();            // that announces them
            if (lane == 0) s_prod = nsym;
        }
        if (lane == 0) {
            s_sp = sp;
            s_nsym = nsym;
            __threadfence();
            s_done = 1;
        }
        return;
    }

    // ---- warp 0: how far the extraction is
    int nsym = 0x7fffffff;   // the call's symbol count, known when warp 1 is done
    bool have_n = false;
    auto wait_for = [&](int need) {   // until `need` symbols exist or the extraction is complete
        while (!have_n) {
            int d = 0, pr = 0;
            if (lane == 0) {
                d = s_done;
                pr = s_prod;
            }
            d = __shfl_sync(0xffffffffu, d, 0);
            pr = __shfl_sync(0xffffffffu, pr, 0);
            if (d) {
                nsym = s_nsym;
                have_n = true;
            } else if (pr >= need) {
                break;
            } else {
                __nanosleep(100);
            }
        }
        __threadfence_block();
    };

    SyncCtx c;
    c.ring = a.ring + (long long)ch * C4_RING;
    c.ptr = (int)(endp - (long long)n_shift * C4_HALF);
    c.shift_mod = (S.shift_mod + n_shift * C4_HALF) & (C4_RING - 1);
    c.sps = sps;
    c.pll = pll0;
    c.gain = gain0;
    c.taps = s_taps;
    c.win = s_win;
    c.win_lo = 0;
    c.win_n = 0;
    c.my_sync = (lane < 24) ? c_sync[lane] : 0.f;
    int fine = S.fine, since = S.since_sync, eq_init = S.eq_init, sync_count = S.sync_count, n_events = 0;
    // The sample point after extraction is warp 1's last value. demodulate() only ever ADDS corrections to it, in event
    // order: they are held back (s_adj) and added in that order once it is known. demodulate_discriminator() reads it at
    // every event: that flavour waits for the extraction first.
    double sample_point = 0.0;
    bool sp_known = false;
    int n_adj = 0;
    auto resolve_sp = [&]() {
        if (sp_known) return;
        wait_for(0x7fffffff);
        __syncwarp();
        sample_point = s_sp;
        for (int i = 0; i < n_adj; ++i) sample_point += s_adj[i];
        sp_known = true;
    };
    if (DISC) resolve_sp();

    if (lane < 24) lagbuf[lane] = S.lag[lane];
    __syncwarp();

    // ---- sync loop (c4fm.py:2596-2807), 32 symbols per step until the first threshold crossing
    int k0 = 0;
    int kw = 0;             // first symbol of the staged soft / index window
    bool win_ok = false;    // false: the window has to be (re)read — at the start and after a message was re-sliced
    while (k0 < nsym) {
        int L = min(32, nsym - k0);
        if (fine) L = min(L, max(1, 3601 - since));
        const int k = k0 + lane;
        const bool active = lane < L;
        // the soft values (and buffer indices) the correlations of the next 256 symbols read, staged once: one pipelined
        // batch of loads per eight steps instead of a dependent global round trip per step
        if (!win_ok || k0 + 32 > kw + C4_SOFTWIN) {
            kw = k0;
            // everything the steps of this window touch: its symbols, and the <= 340 an event re-slices behind them
            wait_for(kw + C4_SOFTWIN + C4_MSG_DIBITS + 32);
            if (k0 >= nsym) break;
            {
                // all loads first, then the stores: through `soft` (volatile) each load would wait for the one before it —
                // nine L2 round trips per reload, a reload per sync event (23 % of the kernel in the ncu source view).
                // ld.cg reads L2, where warp 1's values and this warp's own re-sliced ones are after their fences.
                const float* softp = const_cast<const float*>(soft);
                constexpr int NQ = (23 + C4_SOFTWIN + 31) / 32;
                float tv[NQ];
                int ti[C4_SOFTWIN / 32];
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const int i = lane + 32 * q;
                    const int t = kw - 23 + i;
                    tv[q] = 0.f;
                    if (i < 23 + C4_SOFTWIN) tv[q] = (t < 0) ? S.det[24 + t] : ((t < nsym) ? __ldcg(softp + t) : 0.f);
                }
#pragma unroll
                for (int q = 0; q < C4_SOFTWIN / 32; ++q) {
                    const int i = lane + 32 * q;
                    ti[q] = (kw + i < nsym) ? __ldcg(idx + kw + i) : -1;
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    if (lane + 32 * q < 23 + C4_SOFTWIN) s_soft[lane + 32 * q] = tv[q];
#pragma unroll
                for (int q = 0; q < C4_SOFTWIN / 32; ++q) s_idx[lane + 32 * q] = ti[q];
            }
            win_ok = true;
            __syncwarp();
        }
        const int wo = k0 - kw;
        double sp_score = 0.0;
        if (active) {
#pragma unroll
            for (int i = 0; i < 24; ++i) sp_score += (double)__fmul_rn(c_sync[i], s_soft[wo + lane + i]);
        }
        // lagging detector (:2626-2659), only while acquiring
        bool fed = false;
        float fedval = 0.f;
        const int my_idx = active ? s_idx[wo + lane] : -1;
        if (active && !fine && my_idx >= 0) {
            const int lag_pos = my_idx - a.k.lag_int;
            if (lag_pos >= 4 && lag_pos < C4_RING) {
                const int lo = lag_pos - 4;
                float v;
                if (lo + 1 < C4_RING) {
                    const float x1 = c4_buf(c, lo), x2 = c4_buf(c, lo + 1);
                    if (a.k.lag_mu < 0.0) v = x1;
                    else if (a.k.lag_mu > 1.0) v = x2;
                    else v = __fadd_rn(x1, __fmul_rn(__fsub_rn(x2, x1), (float)a.k.lag_mu));
                } else {
                    v = c4_buf(c, min(max(lo, 0), C4_RING - 1));
                }
                v = __fmul_rn(__fadd_rn(v, (float)c.pll), (float)c.gain);  // _Equalizer.get_equalized_symbol, f32
                fedval = __fmul_rn(v, (float)C4_NORM);
                fed = true;
            }
        }
        const unsigned fedmask = __ballot_sync(0xffffffffu, fed);
        const int my_cnt = __popc(fedmask & ((1u << lane) - 1u));
        if (fed) lagbuf[24 + my_cnt] = fedval;
        __syncwarp();
        double sl = 0.0;
        if (fed) {
            for (int i = 0; i < 24; ++i) sl += (double)__fmul_rn(c_sync[i], lagbuf[my_cnt + 1 + i]);
        }
        const bool use_lag = fed && (sl > sp_score) && (sl >= C4_THRESH);
        const double score = use_lag ? sl : sp_score;
        const unsigned cm = __ballot_sync(0xffffffffu, active && score >= C4_THRESH);
        const int e = cm ? (__ffs(cm) - 1) : L;
        const int nproc = cm ? e + 1 : L;
        // commit the lagging ring: values fed by symbols k0 .. k0+nproc-1
        const int nf = __popc(fedmask & ((nproc >= 32) ? 0xffffffffu : ((1u << nproc) - 1u)));
        float keep = 0.f;
        if (lane < 24) keep = lagbuf[nf + lane];
        __syncwarp();
        if (lane < 24) lagbuf[lane] = keep;
        __syncwarp();
        // symbols before the event: only the since-sync bookkeeping (:2798-2801)
        for (int t = 0; t < e; ++t) {
            ++since;
            if (since > 3600) {
                fine = 0;
                since = 0;
            }
        }
        if (!cm) {
            k0 += L;
            continue;
        }
        // ---- threshold crossing at symbol ke
        const int ke = k0 + e;
        ++since;
        const int idxe = __shfl_sync(0xffffffffu, my_idx, e);
        const bool ul = __shfl_sync(0xffffffffu, (int)use_lag, e) != 0;
        k0 = ke + 1;
        if (idxe < 0) continue;  // shifted out of the buffer: the reference `continue`s (:2663-2664)
        const double extra = ul ? -a.k.lag_offset : 0.0;
        if (DISC) {
            since = 0;
            if (fine) continue;  // since == 0: the 3600-symbol check below cannot fire either
        }
        // demodulate: buffer offset of the sync symbol; demodulate_discriminator passes self._sample_point (:2947-2949)
        const double off = DISC ? sample_point : ((double)idxe + 0.5) + extra;
        // every interpolation of this event reads buffer values within 23 symbols below and (max_adj + step <= 1.125 sps)
        // above / below `off`: staged once
        {
            const int need = (int)(27.0 * sps) + 24;
            if (sps < 1.0e3 && need <= C4_WIN && fabs(off) < 1.0e9) {
                c.win_lo = (int)floor(off - 25.0 * sps) - 8;
                c.win_n = need;
                __syncwarp();
                for (int u = lane; u < need; u += 32) s_win[u] = c4_buf(c, c.win_lo + u);
                __syncwarp();
            } else {
                c.win_n = 0;
            }
        }
        // _timing_optimize_jit (:543-644)
        double step = fine ? sps / 16.0 : sps / 8.0;
        const double step_min = sps / 200.0;
        const double max_adj = fine ? sps : sps / 2.0;
        double adj = 0.0;
        double sc, sL, sR;
        {
            const double o3[3] = {off, off - step, off + step};
            double r3[3];
            c4_scores<3>(c, o3, lane, sterm, r3);
            sc = r3[0];
            sL = r3[1];
            sR = r3[2];
        }
        while (step > step_min && fabs(adj) <= max_adj) {
            if (sL > sR && sL > sc) {
                adj -= step;
                sR = sc;
                sc = sL;
                sL = c4_score(c, (off + adj) - step, lane, sterm);
            } else if (sR > sL && sR > sc) {
                adj += step;
                sL = sc;
                sc = sR;
                sR = c4_score(c, (off + adj) + step, lane, sterm);
            } else {
                step *= 0.5;
                if (step > step_min) {
                    const double o2[2] = {(off + adj) - step, (off + adj) + step};
                    double r2[2];
                    c4_scores<2>(c, o2, lane, sterm, r2);
                    sL = r2[0];
                    sR = r2[1];
                }
            }
        }
        if (DISC) {
            const double total = adj + extra;  // :2950-2958
            if (fabs(total) >= 0.1) {
                sample_point += total;
                if (sample_point >= sps) sample_point -= sps;
                else if (sample_point < 0.0) sample_point += sps;
                fine = 1;
                c.gain = 1.0;
                ++n_events;
            }
            continue;
        }
        double pa, ga;
        c4_correction(c, off + adj, lane, sterm, pa, ga);
        if (sc >= C4_THRESH) {
            if (fine) adj = fmin(fmax(adj, -a.k.max_fine_adj), a.k.max_fine_adj);
            {
                const double val = adj + extra;
                if (!sp_known && n_adj == C4_ADJ) resolve_sp();
                if (sp_known) {
                    sample_point += val;
                } else {
                    if (lane == 0) s_adj[n_adj] = val;
                    ++n_adj;
                }
            }
            // _Equalizer.apply_correction (:260-272)
            if (eq_init) {
                c.pll += pa * C4_LOOP_GAIN;
                c.gain += ga * C4_LOOP_GAIN;
            } else {
                c.pll += pa;
                c.gain += ga;
                eq_init = 1;
            }
            c.pll = fmin(fmax(c.pll, -C4_MAX_PLL), C4_MAX_PLL);
            c.gain = fmin(fmax(c.gain, 1.0), C4_MAX_GAIN);
            ++sync_count;
            ++n_events;
            fine = 1;
            since = 0;
            // _resample_message_jit (:795-869) over the following <= 340 symbols of this call
            const double start0 = (((double)idxe - 23.0 * sps) + adj) + extra;
            const double start = start0 + 24.0 * sps;
            const int nres = min(C4_MSG_DIBITS, nsym - (ke + 1));
            for (int i0 = 0; i0 < nres; i0 += 128) {   // four symbols per lane per trip: the eight loads go out together
                float x1[4], x2[4];
                double mu4[4];
                bool two[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + 32 * u + lane;
                    const double pos = start + (double)i * sps;
                    const int bi = (int)pos;
                    mu4[u] = pos - (double)bi;
                    two[u] = bi >= 0 && bi + 1 < C4_RING;
                    x1[u] = 0.f;
                    x2[u] = 0.f;
                    if (i < nres) {
                        x1[u] = c4_buf(c, two[u] ? bi : min(max(bi, 0), C4_RING - 1));
                        if (two[u]) x2[u] = c4_buf(c, bi + 1);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + 32 * u + lane;
                    if (i < nres) {
                        double v;
                        if (two[u]) {
                            if (mu4[u] < 0.0) v = (double)x1[u];
                            else if (mu4[u] > 1.0) v = (double)x2[u];
                            else v = (double)x1[u] + (double)__fsub_rn(x2[u], x1[u]) * mu4[u];
                        } else {
                            v = (double)x1[u];
                        }
                        const double sr = (v + c.pll) * c.gain;
                        dib[ke + 1 + i] = (unsigned char)c4_slice(sr);
                        soft[ke + 1 + i] = (float)(sr * C4_NORM);
                    }
                }
            }
            win_ok = false;   // the staged soft window is stale from ke + 1 on
            __threadfence();
            __syncwarp();
        }
        if (since > 3600) {
            fine = 0;
            since = 0;
        }
    }

    // ---- state epilogue
    __syncwarp();
    resolve_sp();
    float newdet = 0.f;
    if (lane < 24) {
        const int t = nsym - 24 + lane;
        newdet = (t >= 0) ? soft[t] : S.det[24 + t];
    }
    const float last_phase = (a.n > 0) ? ph[a.n - 1] : prev_phase;
    __syncwarp();
    if (lane < 24) {
        S.det[lane] = newdet;
        S.lag[lane] = lagbuf[lane];
    }
    if (lane == 0) {
        S.sample_point = sample_point;
        S.pll = c.pll;
        S.gain = c.gain;
        S.ptr = c.ptr;
        S.shift_mod = c.shift_mod;
        S.eq_init = eq_init;
        S.fine = fine;
        S.since_sync = since;
        S.sync_count = sync_count;
        S.prev_phase = last_phase;
        S.n_events = n_events;
        a.n_sym[ch] = nsym;
    }
}

__global__ void c4fm_reset_kernel(C4State* st, float* ring, float2* hist0, float2* hist1, int hl, float2* tail0, float2* tail1,
                                  int ov, double sps, int ch_lo, int ch_hi) {
    const int ch = ch_lo + blockIdx.x;
    if (ch >= ch_hi) return;
    for (int i = threadIdx.x; i < C4_RING; i += blockDim.x) ring[(long long)ch * C4_RING + i] = 0.f;
    for (int i = threadIdx.x; i < hl; i += blockDim.x) {
        hist0[(long long)ch * hl + i] = make_float2(0.f, 0.f);
        hist1[(long long)ch * hl + i] = make_float2(0.f, 0.f);
    }
    for (int i = threadIdx.x; i < ov; i += blockDim.x) {
        tail0[(long long)ch * ov + i] = make_float2(0.f, 0.f);
        tail1[(long long)ch * ov + i] = make_float2(0.f, 0.f);
    }
    if (threadIdx.x == 0) {
        C4State s;
        memset(&s, 0, sizeof(s));
        s.sample_point = sps;           // c4fm.py:2488
        s.gain = C4_INITIAL_GAIN;       // _Equalizer.reset, :224-228
        st[ch] = s;
    }
}

// ---- stand-alone stages for the helper classes benchmark_dsp.py times (_Interpolator, _SoftSyncDetector) ----
// _Interpolator.filter (c4fm.py:2204-2253): 8-tap dot product at row clamp(int((1-mu)*128+0.5)); taps outside the
// array contribute nothing (the reference's slow path); float32 products, float64 sum.
__global__ void c4fm_interp_kernel(const float* __restrict__ x, int n, const int* __restrict__ offs, const double* __restrict__ mus,
                                   int count, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int row = (int)((1.0 - mus[i]) * 128 + 0.5);
    row = min(max(row, 0), 128);
    const int o = offs[i];
    double acc = 0.0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int j = o + t;
        if (j >= 0 && j < n) acc += (double)__fmul_rn(x[j], c_interp[row][t]);
    }
    out[i] = acc;
}

// _SoftSyncDetector.process for a block of symbols (c4fm.py:2306-2329): score[k] = sum_i sync[i] * s[k-23+i] with the
// 24 symbols before the block in hist (oldest first); new_hist = last 24 of [hist | s].
__global__ void c4fm_sync_score_kernel(const float* __restrict__ s, int n, const float* __restrict__ hist, double* __restrict__ score,
                                       float* __restrict__ new_hist) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < 24) {
        const int t = n - 24 + k;
        new_hist[k] = (t >= 0) ? s[t] : hist[24 + t];
    }
    if (k >= n) return;
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
        const int t = k - 23 + i;
        const float v = (t >= 0) ? s[t] : hist[24 + t];
        acc += (double)__fmul_rn(c_sync[i], v);
    }
    score[k] = acc;
}

// ---- host-side filter design (used when the caller passes no taps) ----
// scipy.signal.firwin(numtaps, cutoff_hz, fs=fs, window="hamming"): the branch design_baseband_lpf
// (c4fm.py:95-132) ends up in with scipy >= 1.15, where remez(..., Hz=) raises.
static void firwin_hamming_lowpass(int numtaps, double cutoff_norm, std::vector<double>& h) {
    h.resize(numtaps);
    const double alpha = 0.5 * (numtaps - 1);
    double s = 0.0;
    for (int n = 0; n < numtaps; ++n) {
        const double m = n - alpha;
        const double xx = cutoff_norm * m;
        const double sinc = (xx == 0.0) ? 1.0 : sin(M_PI * xx) / (M_PI * xx);
        const double win = 0.54 - 0.46 * cos(2.0 * M_PI * n / (numtaps - 1));
        h[n] = cutoff_norm * sinc * win;
        s += h[n];
    }
    for (int n = 0; n < numtaps; ++n) h[n] /= s;
}

// design_rrc_filter (c4fm.py:135-183): sum-normalised root raised cosine
static void design_rrc(double sps, int num_taps, double alpha, std::vector<double>& h) {
    if (num_taps % 2 == 0) ++num_taps;
    h.resize(num_taps);
    double s = 0.0;
    for (int i = 0; i < num_taps; ++i) {
        const double t = ((double)i - (num_taps - 1) / 2.0) / sps;
        double v;
        if (t == 0.0) v = 1.0 - alpha + 4.0 * alpha / M_PI;
        else if (fabs(t) == 1.0 / (4.0 * alpha))
            v = (alpha / sqrt(2.0)) * ((1.0 + 2.0 / M_PI) * sin(M_PI / (4.0 * alpha)) + (1.0 - 2.0 / M_PI) * cos(M_PI / (4.0 * alpha)));
        else
            v = (sin(M_PI * t * (1.0 - alpha)) + 4.0 * alpha * t * cos(M_PI * t * (1.0 + alpha))) /
                (M_PI * t * (1.0 - (4.0 * alpha * t) * (4.0 * alpha * t)));
        h[i] = v;
        s += v;
    }
    for (auto& v : h) v /= s;
}

}  // namespace wc

using namespace wc;

struct wc_c4fm {
    int C = 0;
    int sample_rate = 0, symbol_rate = 0;
    C4Const k;
    int n_lpf = 0, n_rrc = 0, ntp = 0, hl = 0;
    std::vector<float> lpf, rrc;
    double* d_taps = nullptr;
    C4State* d_state = nullptr;
    float* d_ring = nullptr;
    float2* d_hist[2] = {nullptr, nullptr};
    float2* d_tail[2] = {nullptr, nullptr};
    int cur = 0;
    // per-call scratch
    float2* d_filt = nullptr;  size_t filt_cap = 0;
    float* d_ph = nullptr;     size_t ph_cap = 0;
    int* d_idx = nullptr;      size_t idx_cap = 0;
    // host API staging
    void* d_in = nullptr;      size_t in_cap = 0;
    unsigned char* d_dib = nullptr; size_t dib_cap = 0;
    float* d_soft = nullptr;   size_t soft_cap = 0;
    int* d_nsym = nullptr;
    // discriminator-audio entry (demodulate_discriminator): its own RRC state, untouched by reset() like the reference's
    double* d_rrc64 = nullptr;
    double* d_dhist[2] = {nullptr, nullptr};
    int* d_dinit = nullptr;
    double* d_dfirst = nullptr;
    int dcur = 0;
    cudaStream_t stream = nullptr;
};

template <typename T>
static int ensure_buf(T** p, size_t* cap, size_t need_elems) {
    if (*cap >= need_elems) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    WC_CUDA(cudaMalloc((void**)p, need_elems * sizeof(T)));
    *cap = need_elems;
    return 0;
}

static int c4fm_reset_range(wc_c4fm* h, int lo, int hi, cudaStream_t s) {
    c4fm_reset_kernel<<<hi - lo, 256, 0, s>>>(h->d_state, h->d_ring, h->d_hist[0], h->d_hist[1], h->hl, h->d_tail[0],
                                             h->d_tail[1], h->k.ov, h->k.sps, lo, hi);
    WC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" {

int wc_c4fm_create(int n_channels, int sample_rate, int symbol_rate, int wide_pulse, const float* lpf_taps, int n_lpf,
                   const float* rrc_taps, int n_rrc, wc_c4fm** out) {
    WC_REQUIRE(out != nullptr, "wc_c4fm_create: out is null");
    WC_REQUIRE(n_channels >= 1 && sample_rate > 0 && symbol_rate > 0, "wc_c4fm_create: bad parameters");
    const double sps = (double)sample_rate / (double)symbol_rate;
    WC_REQUIRE(sps >= 2.0 && sps <= 24.0, "wc_c4fm_create: samples per symbol %.3f outside [2, 24]", sps);
    wc_c4fm* h = new wc_c4fm();
    h->C = n_channels;
    h->sample_rate = sample_rate;
    h->symbol_rate = symbol_rate;
    // filters: as passed (the Python host designs them with scipy exactly like c4fm.py:2436-2462) or designed here
    if (lpf_taps && n_lpf > 0) h->lpf.assign(lpf_taps, lpf_taps + n_lpf);
    else {
        std::vector<double> d;
        firwin_hamming_lowpass(63, (wide_pulse ? 10000.0 : 5200.0) / (sample_rate / 2.0), d);
        h->lpf.resize(d.size());
        for (size_t i = 0; i < d.size(); ++i) h->lpf[i] = (float)d[i];
    }
    if (rrc_taps && n_rrc > 0) h->rrc.assign(rrc_taps, rrc_taps + n_rrc);
    else {
        std::vector<double> d;
        design_rrc(sps, (int)(16 * sps) + 1, wide_pulse ? 0.5 : 0.2, d);
        h->rrc.resize(d.size());
        for (size_t i = 0; i < d.size(); ++i) h->rrc[i] = (float)d[i];
    }
    h->n_lpf = (int)h->lpf.size();
    h->n_rrc = (int)h->rrc.size();
    const int nt = h->n_lpf + h->n_rrc - 1;
    h->ntp = (nt + 7) & ~7;
    if (h->ntp > FIR_MAX_TAPS) {
        set_error("wc_c4fm_create: combined filter length %d exceeds %d", nt, FIR_MAX_TAPS);
        delete h;
        return -1;
    }
    h->hl = h->ntp - 1;
    // combined taps: float32 designs promoted to float64 (lfilter computes in f64), convolved in long double
    std::vector<double> comb(h->ntp, 0.0);
    for (int t = 0; t < nt; ++t) {
        long double s = 0.0L;
        for (int i = 0; i < h->n_lpf; ++i) {
            const int j = t - i;
            if (j >= 0 && j < h->n_rrc) s += (long double)(double)h->lpf[i] * (long double)(double)h->rrc[j];
        }
        comb[t] = (double)s;
    }
    // _FMDemodulator.__init__ (c4fm.py:289-317)
    const double fl = floor(sps);
    const double mu = fmod(sps, 1.0);
    h->k.sps = sps;
    h->k.ov = (int)fl + 4;
    h->k.interp_off = ((int)fl - 4 > 0) ? (int)fl - 4 : 0;
    int row = (int)((1.0 - mu) * 128 + 0.5);
    h->k.row = row < 0 ? 0 : (row > 128 ? 128 : row);
    h->k.lag_offset = sps / 2.0;
    h->k.lag_int = (int)h->k.lag_offset;
    h->k.lag_mu = 1.0 - (h->k.lag_offset - (double)h->k.lag_int);
    h->k.max_fine_adj = sps * 0.2;

    float sync[24];
    const unsigned long long pat = 0x5575F5FF77FFull;
    for (int i = 0; i < 24; ++i) sync[i] = (((pat >> ((23 - i) * 2)) & 3ull) == 1ull) ? 3.0f : -3.0f;

    const size_t C = (size_t)n_channels;
    bool ok = cudaMalloc(&h->d_taps, sizeof(double) * h->ntp) == cudaSuccess &&
              cudaMalloc(&h->d_state, sizeof(C4State) * C) == cudaSuccess &&
              cudaMalloc(&h->d_ring, sizeof(float) * C4_RING * C) == cudaSuccess &&
              cudaMalloc(&h->d_hist[0], sizeof(float2) * h->hl * C) == cudaSuccess &&
              cudaMalloc(&h->d_hist[1], sizeof(float2) * h->hl * C) == cudaSuccess &&
              cudaMalloc(&h->d_tail[0], sizeof(float2) * h->k.ov * C) == cudaSuccess &&
              cudaMalloc(&h->d_tail[1], sizeof(float2) * h->k.ov * C) == cudaSuccess &&
              cudaMalloc(&h->d_nsym, sizeof(int) * C) == cudaSuccess &&
              cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMemcpyToSymbol(c_sync, sync, sizeof(sync)) == cudaSuccess &&
              cudaMemcpy(h->d_taps, comb.data(), sizeof(double) * h->ntp, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        set_error("wc_c4fm_create: CUDA allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete h;
        return -2;
    }
    if (c4fm_reset_range(h, 0, n_channels, h->stream)) return -2;
    WC_CUDA(cudaStreamSynchronize(h->stream));
    *out = h;
    return 0;
}

void wc_c4fm_destroy(wc_c4fm* h) {
    if (!h) return;
    cudaFree(h->d_taps);
    cudaFree(h->d_state);
    cudaFree(h->d_ring);
    for (int i = 0; i < 2; ++i) {
        cudaFree(h->d_hist[i]);
        cudaFree(h->d_tail[i]);
    }
    cudaFree(h->d_nsym);
    if (h->d_rrc64) cudaFree(h->d_rrc64);
    if (h->d_dhist[0]) cudaFree(h->d_dhist[0]);
    if (h->d_dhist[1]) cudaFree(h->d_dhist[1]);
    if (h->d_dinit) cudaFree(h->d_dinit);
    if (h->d_dfirst) cudaFree(h->d_dfirst);
    if (h->d_filt) cudaFree(h->d_filt);
    if (h->d_ph) cudaFree(h->d_ph);
    if (h->d_idx) cudaFree(h->d_idx);
    if (h->d_in) cudaFree(h->d_in);
    if (h->d_dib) cudaFree(h->d_dib);
    if (h->d_soft) cudaFree(h->d_soft);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int wc_c4fm_info(const wc_c4fm* h, int* n_channels, double* samples_per_symbol, int* n_lpf, int* n_rrc) {
    WC_REQUIRE(h != nullptr, "wc_c4fm_info: null handle");
    if (n_channels) *n_channels = h->C;
    if (samples_per_symbol) *samples_per_symbol = h->k.sps;
    if (n_lpf) *n_lpf = h->n_lpf;
    if (n_rrc) *n_rrc = h->n_rrc;
    return 0;
}

int wc_c4fm_get_taps(const wc_c4fm* h, float* lpf, float* rrc) {
    WC_REQUIRE(h != nullptr, "wc_c4fm_get_taps: null handle");
    if (lpf) memcpy(lpf, h->lpf.data(), sizeof(float) * h->lpf.size());
    if (rrc) memcpy(rrc, h->rrc.data(), sizeof(float) * h->rrc.size());
    return 0;
}

int wc_c4fm_max_symbols(const wc_c4fm* h, int n_samples) {
    if (!h || n_samples <= 0) return 0;
    // one symbol per >= 1 sample in the worst case of a huge negative timing adjustment is bounded by the
    // reference to +-sps per event; n/floor(sps)+2 covers the steady state, +26 the adjustments
    return (int)(n_samples / floor(h->k.sps)) + 28;
}

int wc_c4fm_reset(wc_c4fm* h, int channel) {
    WC_REQUIRE(h != nullptr, "wc_c4fm_reset: null handle");
    WC_REQUIRE(channel >= -1 && channel < h->C, "wc_c4fm_reset: channel %d out of range", channel);
    const int lo = channel < 0 ? 0 : channel, hi = channel < 0 ? h->C : channel + 1;
    if (c4fm_reset_range(h, lo, hi, h->stream)) return -2;
    WC_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int wc_c4fm_demod(wc_c4fm* h, const void* iq_dev, long long chan_stride, int n_samples, unsigned char* dibits_dev,
                  float* soft_dev, int* n_sym_dev, int max_sym, void* stream_v) {
    WC_REQUIRE(h && iq_dev && dibits_dev && soft_dev && n_sym_dev, "wc_c4fm_demod: null argument");
    WC_REQUIRE(n_samples >= 0 && chan_stride >= n_samples, "wc_c4fm_demod: bad sizes");
    cudaStream_t s = (cudaStream_t)stream_v;
    const int C = h->C;
    if (n_samples == 0) {
        WC_CUDA(cudaMemsetAsync(n_sym_dev, 0, sizeof(int) * C, s));
        return 0;
    }
    WC_REQUIRE(max_sym >= wc_c4fm_max_symbols(h, n_samples), "wc_c4fm_demod: max_sym %d < %d", max_sym,
               wc_c4fm_max_symbols(h, n_samples));
    if (ensure_buf(&h->d_filt, &h->filt_cap, (size_t)C * n_samples)) return -2;
    if (ensure_buf(&h->d_ph, &h->ph_cap, (size_t)C * n_samples)) return -2;
    if (ensure_buf(&h->d_idx, &h->idx_cap, (size_t)C * max_sym)) return -2;
    const float2* x = reinterpret_cast<const float2*>(iq_dev);
    FirArgs f;
    f.x = x;
    f.stride = chan_stride;
    f.n = n_samples;
    f.hist = h->d_hist[h->cur];
    f.taps = h->d_taps;
    f.ntp = h->ntp;
    f.y = h->d_filt;
    dim3 fg((n_samples + FIR_TILE - 1) / FIR_TILE, C);
    p25_fir_kernel<<<fg, FIR_THREADS, 0, s>>>(f);
    p25_hist_kernel<<<C, 128, 0, s>>>(x, chan_stride, n_samples, h->d_hist[h->cur], h->d_hist[h->cur ^ 1], h->hl);
    PhaseArgs p;
    p.filt = h->d_filt;
    p.tail = h->d_tail[h->cur];
    p.new_tail = h->d_tail[h->cur ^ 1];
    p.n = n_samples;
    p.k = h->k;
    p.st = h->d_state;
    p.ph = h->d_ph;
    p.ring = h->d_ring;
    const int pn = n_samples > h->k.ov ? n_samples : h->k.ov;
    dim3 pg((pn + 255) / 256, C);
    c4fm_phase_kernel<<<pg, 256, 0, s>>>(p);
    SyncArgs y;
    y.n = n_samples;
    y.max_sym = max_sym;
    y.k = h->k;
    y.st = h->d_state;
    y.ph = h->d_ph;
    y.ring = h->d_ring;
    y.dibits = dibits_dev;
    y.soft = soft_dev;
    y.idx = h->d_idx;
    y.n_sym = n_sym_dev;
    c4fm_sync_kernel<false><<<C, C4_THREADS, 0, s>>>(y);
    WC_CUDA(cudaGetLastError());
    h->cur ^= 1;
    return 0;
}

int wc_c4fm_demod_host(wc_c4fm* h, const void* iq_host, int n_samples, unsigned char* dibits_host, float* soft_host,
                       int* n_sym_host, int max_sym) {
    WC_REQUIRE(h && iq_host && dibits_host && soft_host && n_sym_host, "wc_c4fm_demod_host: null argument");
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
    if (ensure_buf(&h->d_dib, &h->dib_cap, (size_t)C * max_sym)) return -2;
    if (ensure_buf(&h->d_soft, &h->soft_cap, (size_t)C * max_sym)) return -2;
    WC_CUDA(cudaMemcpyAsync(h->d_in, iq_host, in_bytes, cudaMemcpyHostToDevice, h->stream));
    int rc = wc_c4fm_demod(h, h->d_in, n_samples, n_samples, h->d_dib, h->d_soft, h->d_nsym, max_sym, h->stream);
    if (rc) return rc;
    WC_CUDA(cudaMemcpyAsync(dibits_host, h->d_dib, (size_t)C * max_sym, cudaMemcpyDeviceToHost, h->stream));
    WC_CUDA(cudaMemcpyAsync(soft_host, h->d_soft, sizeof(float) * (size_t)C * max_sym, cudaMemcpyDeviceToHost, h->stream));
    WC_CUDA(cudaMemcpyAsync(n_sym_host, h->d_nsym, sizeof(int) * C, cudaMemcpyDeviceToHost, h->stream));
    WC_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

/* C4FMDemodulator.demodulate_discriminator (c4fm.py:2817-2992): discriminator audio float32 [C][chan_stride] ->
 * dibits / soft / n_sym like wc_c4fm_demod. first_host: float64 [C], audio[0] of each channel in the caller's own
 * precision — only read on a channel's first call (it scales lfilter_zi). Shares the phase buffer, sample point,
 * equaliser and sync detectors with wc_c4fm_demod exactly like the two methods share one Python object. */
int wc_c4fm_demod_disc(wc_c4fm* h, const float* audio_dev, long long chan_stride, int n_samples, const double* first_host,
                       unsigned char* dibits_dev, float* soft_dev, int* n_sym_dev, int max_sym, void* stream_v) {
    WC_REQUIRE(h && audio_dev && dibits_dev && soft_dev && n_sym_dev, "wc_c4fm_demod_disc: null argument");
    WC_REQUIRE(n_samples >= 0 && chan_stride >= n_samples, "wc_c4fm_demod_disc: bad sizes");
    cudaStream_t s = (cudaStream_t)stream_v;
    const int C = h->C;
    if (n_samples == 0) {
        WC_CUDA(cudaMemsetAsync(n_sym_dev, 0, sizeof(int) * C, s));
        return 0;
    }
    WC_REQUIRE(max_sym >= wc_c4fm_max_symbols(h, n_samples), "wc_c4fm_demod_disc: max_sym %d < %d", max_sym,
               wc_c4fm_max_symbols(h, n_samples));
    const int hl = h->n_rrc - 1;
    if (!h->d_rrc64) {
        std::vector<double> t(h->rrc.begin(), h->rrc.end());
        WC_CUDA(cudaMalloc(&h->d_rrc64, sizeof(double) * h->n_rrc));
        WC_CUDA(cudaMalloc(&h->d_dhist[0], sizeof(double) * (size_t)C * (hl > 0 ? hl : 1)));
        WC_CUDA(cudaMalloc(&h->d_dhist[1], sizeof(double) * (size_t)C * (hl > 0 ? hl : 1)));
        WC_CUDA(cudaMalloc(&h->d_dinit, sizeof(int) * (size_t)C));
        WC_CUDA(cudaMalloc(&h->d_dfirst, sizeof(double) * (size_t)C));
        WC_CUDA(cudaMemcpy(h->d_rrc64, t.data(), sizeof(double) * h->n_rrc, cudaMemcpyHostToDevice));
        WC_CUDA(cudaMemset(h->d_dinit, 0, sizeof(int) * (size_t)C));
        WC_CUDA(cudaMemset(h->d_dfirst, 0, sizeof(double) * (size_t)C));
    }
    if (first_host) WC_CUDA(cudaMemcpyAsync(h->d_dfirst, first_host, sizeof(double) * (size_t)C, cudaMemcpyHostToDevice, s));
    if (ensure_buf(&h->d_ph, &h->ph_cap, (size_t)C * n_samples)) return -2;
    if (ensure_buf(&h->d_idx, &h->idx_cap, (size_t)C * max_sym)) return -2;
    DiscArgs d;
    d.x = audio_dev;
    d.stride = chan_stride;
    d.n = n_samples;
    d.hist = h->d_dhist[h->dcur];
    d.new_hist = h->d_dhist[h->dcur ^ 1];
    d.taps = h->d_rrc64;
    d.nt = h->n_rrc;
    d.init = h->d_dinit;
    d.first = h->d_dfirst;
    d.sps = h->k.sps;
    d.st = h->d_state;
    d.ph = h->d_ph;
    d.ring = h->d_ring;
    dim3 g((n_samples + 255) / 256, C);
    disc_fir_kernel<<<g, 256, 0, s>>>(d);
    disc_hist_kernel<<<C, 128, 0, s>>>(d);
    h->dcur ^= 1;
    SyncArgs y;
    y.n = n_samples;
    y.max_sym = max_sym;
    y.k = h->k;
    y.st = h->d_state;
    y.ph = h->d_ph;
    y.ring = h->d_ring;
    y.dibits = dibits_dev;
    y.soft = soft_dev;
    y.idx = h->d_idx;
    y.n_sym = n_sym_dev;
    c4fm_sync_kernel<true><<<C, C4_THREADS, 0, s>>>(y);
    WC_CUDA(cudaGetLastError());
    return 0;
}

int wc_c4fm_demod_disc_host(wc_c4fm* h, const float* audio_host, int n_samples, const double* first_host,
                            unsigned char* dibits_host, float* soft_host, int* n_sym_host, int max_sym) {
    WC_REQUIRE(h && audio_host && dibits_host && soft_host && n_sym_host, "wc_c4fm_demod_disc_host: null argument");
    const int C = h->C;
    if (n_samples <= 0) {
        for (int c = 0; c < C; ++c) n_sym_host[c] = 0;
        return 0;
    }
    const size_t in_bytes = sizeof(float) * (size_t)C * n_samples;
    if (h->in_cap < in_bytes) {
        if (h->d_in) cudaFree(h->d_in);
        h->d_in = nullptr;
        h->in_cap = 0;
        WC_CUDA(cudaMalloc(&h->d_in, in_bytes));
        h->in_cap = in_bytes;
    }
    if (ensure_buf(&h->d_dib, &h->dib_cap, (size_t)C * max_sym)) return -2;
    if (ensure_buf(&h->d_soft, &h->soft_cap, (size_t)C * max_sym)) return -2;
    WC_CUDA(cudaMemcpyAsync(h->d_in, audio_host, in_bytes, cudaMemcpyHostToDevice, h->stream));
    int rc = wc_c4fm_demod_disc(h, (const float*)h->d_in, n_samples, n_samples, first_host, h->d_dib, h->d_soft, h->d_nsym,
                                max_sym, h->stream);
    if (rc) return rc;
    WC_CUDA(cudaMemcpyAsync(dibits_host, h->d_dib, (size_t)C * max_sym, cudaMemcpyDeviceToHost, h->stream));
    WC_CUDA(cudaMemcpyAsync(soft_host, h->d_soft, sizeof(float) * (size_t)C * max_sym, cudaMemcpyDeviceToHost, h->stream));
    WC_CUDA(cudaMemcpyAsync(n_sym_host, h->d_nsym, sizeof(int) * C, cudaMemcpyDeviceToHost, h->stream));
    WC_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

/* _FMDemodulator.demodulate (c4fm.py:324-395) on its own: filtered (i, q) pairs [C][n] -> phases float32 [C][n], with the
 * overlap carried in the handle like the reference object does. A handle used this way should not also be used for
 * wc_c4fm_demod (both advance the same overlap state). */
int wc_c4fm_diffdemod(wc_c4fm* h, const void* iq_pairs_dev, int n_samples, float* phases_dev, void* stream_v) {
    WC_REQUIRE(h && iq_pairs_dev && phases_dev, "wc_c4fm_diffdemod: null argument");
    if (n_samples <= 0) return 0;
    PhaseArgs p;
    p.filt = reinterpret_cast<const float2*>(iq_pairs_dev);
    p.tail = h->d_tail[h->cur];
    p.new_tail = h->d_tail[h->cur ^ 1];
    p.n = n_samples;
    p.k = h->k;
    p.st = h->d_state;
    p.ph = phases_dev;
    p.ring = h->d_ring;
    const int pn = n_samples > h->k.ov ? n_samples : h->k.ov;
    c4fm_phase_kernel<<<dim3((pn + 255) / 256, h->C), 256, 0, (cudaStream_t)stream_v>>>(p);
    WC_CUDA(cudaGetLastError());
    h->cur ^= 1;
    return 0;
}

/* _Interpolator.filter for `count` (offset, mu) pairs over one float32 sample array (c4fm.py:2204-2253) -> float64 */
int wc_c4fm_interp(const float* samples_dev, int n, const int* offsets_dev, const double* mus_dev, int count, double* out_dev,
                   void* stream_v) {
    WC_REQUIRE(samples_dev && offsets_dev && mus_dev && out_dev, "wc_c4fm_interp: null argument");
    if (count <= 0) return 0;
    c4fm_interp_kernel<<<(count + 127) / 128, 128, 0, (cudaStream_t)stream_v>>>(samples_dev, n, offsets_dev, mus_dev, count, out_dev);
    WC_CUDA(cudaGetLastError());
    return 0;
}

/* _SoftSyncDetector.process over a block (c4fm.py:2268-2329): scores float64 [n]; hist24 in/out are device float32[24] */
int wc_c4fm_sync_scores(const float* soft_dev, int n, const float* hist24_dev, double* scores_dev, float* new_hist24_dev,
                        void* stream_v) {
    WC_REQUIRE(soft_dev && hist24_dev && scores_dev && new_hist24_dev, "wc_c4fm_sync_scores: null argument");
    if (n <= 0) return 0;
    float sync[24];
    const unsigned long long pat = 0x5575F5FF77FFull;
    for (int i = 0; i < 24; ++i) sync[i] = (((pat >> ((23 - i) * 2)) & 3ull) == 1ull) ? 3.0f : -3.0f;
    WC_CUDA(cudaMemcpyToSymbolAsync(c_sync, sync, sizeof(sync), 0, cudaMemcpyHostToDevice, (cudaStream_t)stream_v));
    const int m = n > 24 ? n : 24;
    c4fm_sync_score_kernel<<<(m + 127) / 128, 128, 0, (cudaStream_t)stream_v>>>(soft_dev, n, hist24_dev, scores_dev, new_hist24_dev);
    WC_CUDA(cudaGetLastError());
    return 0;
}

/* state[ch] = {pll, gain, sample_point, buffer_pointer, fine_sync, symbols_since_sync, sync_count, events_last_call} */
int wc_c4fm_get_state(wc_c4fm* h, int channel, double* state8) {
    WC_REQUIRE(h && state8, "wc_c4fm_get_state: null argument");
    WC_REQUIRE(channel >= 0 && channel < h->C, "wc_c4fm_get_state: channel %d out of range", channel);
    C4State s;
    WC_CUDA(cudaDeviceSynchronize());
    WC_CUDA(cudaMemcpy(&s, h->d_state + channel, sizeof(s), cudaMemcpyDeviceToHost));
    state8[0] = s.pll;
    state8[1] = s.gain;
    state8[2] = s.sample_point;
    state8[3] = s.ptr;
    state8[4] = s.fine;
    state8[5] = s.since_sync;
    state8[6] = s.sync_count;
    state8[7] = s.n_events;
    return 0;
}

}  // extern "C"
