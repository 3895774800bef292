// Analog demod chain (configs C1/C2): the per-(chunk, channel) work of
// wavecapsdr/capture.py:298-439 (_process_channel_dsp_stateless) as batched GPU stages.
//
//   front   capture.freq_shift (capture.py:166-193, float32-phase NCO restarted per chunk)
//           + RSSI power (capture.py:331-334) + demod front end:
//             FM  : dsp/fm.py:65-97 quadrature_demod
//             AM  : |base|                    (dsp/am.py:99)
//             SSB : Re(base * exp(+j*2*pi*bfo*t)), float64 phase (dsp/am.py:23-42,219-223)
//           One CTA stages a tile of IQ (cf32 or cs16) in shared memory ONCE and loops over all
//           channels, so HBM sees each IQ byte once per chunk no matter how many channels.
//   iir     scipy.signal.lfilter(b, a, x) with zero initial state, float64 direct-form-II-transposed
//           (dsp/fm.py:123,178; dsp/filters.py:124,170,217,260; dsp/agc.py:93,100) as a two-level
//           block scan: 64-sample segments run in parallel from zero state, segment end states are
//           chained with the exact transition matrix A^64, then every segment is re-run from its
//           true start state. Linear recurrences compose exactly, so this equals the sequential
//           recursion up to float64 rounding.
//   rms     dsp/fm.py:42-62 (sum of squares per sequence)
//   resamp  scipy.signal.resample_poly index math (upfirdn with zero extension), float64
//           accumulation, only kept outputs are computed; epilogue fuses the RMS scale, tanh soft
//           clip (dsp/fm.py:26-39), the finite/|x|<=1.2 validity gate (validation.py:41-52) and the
//           audio power reduction (capture.py:436-437).
//   agc     dsp/agc.py:169-242 (two one-pole envelopes via `iir`, then gain + tanh).
#include <math.h>
#include <string.h>
#include <vector>

#include "../../include/wcsdr_b200.h"
#include "common.cuh"

namespace wc {

// ---------------------------------------------------------------------------------------------
// front end
// ---------------------------------------------------------------------------------------------
constexpr int FR_TILE = 4096;
constexpr int FR_THREADS = 256;

struct FrontChan {
    float k32;        // float32(-2*pi*round(offset)/fs); 0 => no shift (capture.py:185-186)
    int shift;        // 0: offset == 0 (samples pass through untouched)
    int mode;         // WC_MODE_*
    double bfo_turns; // SSB: +-bfo/fs in turns per sample
    float disc_scale; // float32(fs/(2*pi*75000))
};

struct FrontArgs {
    const void* iq;        // [n_chunks][chunk_stride] cf32 or cs16 pairs
    long long chunk_stride;
    int n;                 // samples per chunk
    int fmt;               // 0 cf32, 1 cs16 (scaled by 1/32768, cli.py:449-453)
    int n_ch;
    const FrontChan* ch;   // [n_ch]
    float* out;            // [n_ch][n_chunks][n] demod front-end output
    float2* base_out;      // optional [n_ch][n_chunks][n] shifted IQ (freq_shift result), may be null
    double* power;         // [n_ch][n_chunks] sum |base|^2
    double* out_sumsq;     // optional [n_ch][n_chunks] sum out^2 (rms_normalize input when no filter sits in between)
    int n_chunks;
    int* nonfinite;        // [n_chunks] set to 1 if any input sample is not finite
};

// nco_f32 (capture.freq_shift's float32-phase oscillator) lives in common.cuh

// The same oscillator for the front end's sample loop, where it is a quarter of the instruction stream: theta =
// fl32(k32 * fl32(n)) exactly as numpy forms it, reduced to turns in float-float arithmetic (as nco_f32f), then
// cos/sin from a 256-entry table (shared memory, one 8-byte load) times a small-angle rotation for the remaining
// |delta| <= 2 pi / 512: sin d = d (1 - d^2/6) and cos d = 1 - d^2/2 + d^4/24 are exact to 2e-12 there, the table entries to
// float32 rounding — the accuracy class of sincospif at about half its instructions (no quadrant logic, two short
// polynomials instead of two long ones).
constexpr int NCO_TAB = 256;
__device__ __forceinline__ void nco_f32f_tab(float k32, float nf, const float2* __restrict__ tab, float& c, float& s) {
    const float th = __fmul_rn(k32, nf);
    const float C_HI = 0.15915493667125702f, C_LO = 6.4206382432985265e-09f;   // 1/(2 pi) = C_HI + C_LO
    const float hi = __fmul_rn(th, C_HI);
    const float e = __fmaf_rn(th, C_HI, -hi);
    const float lo = __fmaf_rn(th, C_LO, e);
    const float fr = __fadd_rn(__fsub_rn(hi, rintf(hi)), lo);        // turns, |fr| <= 0.5 (+ a few ulp)
    const float kf = rintf(fr * (float)NCO_TAB);
    const float rem = fmaf(kf, -1.0f / (float)NCO_TAB, fr);           // exact: |rem| <= 1 / 512 turn
    const float2 t = tab[(int)kf & (NCO_TAB - 1)];                    // (cos, sin)(2 pi k / 256); k = -128 and 128 share an entry
    const float d = rem * 6.283185307179586f;
    const float d2 = d * d;
    const float sd = d * fmaf(d2, -1.0f / 6.0f, 1.0f);
    const float cd = fmaf(d2, fmaf(d2, 1.0f / 24.0f, -0.5f), 1.0f);
    c = fmaf(t.x, cd, -t.y * sd);
    s = fmaf(t.y, cd, t.x * sd);
}

// One channel over one staged tile, specialised on what the channel needs so that the sample loop carries no mode
// tests (ncu: the generic loop spent 14 of its 200 instructions per sample on branches). KIND: 0 = power only (NONE /
// RAW), 1 = FM discriminator, 2 = AM envelope, 3 = SSB product. Returns this thread's float32 partials of
// (sum |base|^2, sum out^2).
template <int KIND, bool SHIFT, bool BASE>
__device__ __forceinline__ float2 front_tile(const FrontArgs& a, const FrontChan& ch, const float2* tile, const float2* nco_tab,
                                            int t0, int cnt, long long obase, int tid) {
    float psum32 = 0.f, osum32 = 0.f;   // this thread's <= 16 samples of the tile; everything above that level is float64
    const bool exact_idx = (t0 + FR_TILE) <= (1 << 24);   // float32 index by exact increments (numpy's float32 arange)
    // Each warp walks its own contiguous segment of the tile, 32 consecutive samples per step (coalesced stores). The FM
    // discriminator's previous mixed sample is the neighbouring lane's current one (shuffle) and, for lane 0, lane 31's of
    // the previous step — so the oscillator and the complex product run once per sample; only a segment's very first
    // sample evaluates them a second time (ncu: recomputing it in lane 0 of every step cost the whole warp ~35 issue
    // slots per step).
    constexpr int SEG = FR_TILE / (FR_THREADS / 32);
    const int lane = tid & 31;
    const int seg0 = (tid >> 5) * SEG;
    float nf = (float)(t0 + seg0 + lane);
    float2 carry = make_float2(0.f, 0.f);     // lane 31's mixed sample of the previous step
    if (seg0 < cnt) {
        if (KIND == 1) {
            carry = tile[seg0];               // sample t0 + seg0 - 1 (the tile carries one sample of halo in front)
            if (SHIFT) {
                float c0, s0;
                nco_f32(ch.k32, t0 + seg0 - 1, c0, s0);
                carry = make_float2(carry.x * c0 - carry.y * s0, carry.x * s0 + carry.y * c0);
            }
        }
#pragma unroll 2
        for (int base = seg0; base < min(cnt, seg0 + SEG); base += 32, nf += 32.0f) {
            const int i = base + lane;
            const bool live = i < cnt;
            const int n = t0 + i;
            float2 b1 = live ? tile[i + 1] : make_float2(0.f, 0.f);
            if (SHIFT) {
                float c1, s1;
                nco_f32f_tab(ch.k32, exact_idx ? nf : (float)n, nco_tab, c1, s1);
                b1 = make_float2(b1.x * c1 - b1.y * s1, b1.x * s1 + b1.y * c1);
            }
            float o = 0.f;
            if (KIND == 1) {
                float2 b0;
                b0.x = __shfl_up_sync(0xffffffffu, b1.x, 1);
                b0.y = __shfl_up_sync(0xffffffffu, b1.y, 1);
                if (lane == 0) b0 = carry;
                carry.x = __shfl_sync(0xffffffffu, b1.x, 31);
                carry.y = __shfl_sync(0xffffffffu, b1.y, 31);
                // angle(x[n] * conj(x[n-1])) * scale, out[0] = 0
                const float pr = b1.x * b0.x + b1.y * b0.y;
                const float pi = b1.y * b0.x - b1.x * b0.y;
                o = (n == 0) ? 0.0f : fast_atan2f_hi(pi, pr) * ch.disc_scale;
            }
            const float pw = fmaf(b1.x, b1.x, b1.y * b1.y);      // |base|^2 (np.abs(base)**2 to 1 ulp); 0 for dead lanes
            psum32 += pw;
            if (KIND == 2) o = sqrtf(pw);                         // np.abs(base)
            if (KIND == 3) {
                // t = n / fs in float64; shift = complex64(exp(2j*pi*f*t)); real(iq * shift)
                double s, co;
                const double ph = ch.bfo_turns * (double)n;
                sincospi(2.0 * (ph - rint(ph)), &s, &co);
                o = b1.x * (float)co - b1.y * (float)s;
            }
            if (live) {
                // RAW (capture.py:415-420) is served by base_out; NONE only needs the power sum
                if (KIND != 0) {
                    a.out[obase + n] = o;
                    osum32 = fmaf(o, o, osum32);
                }
                if (BASE) a.base_out[obase + n] = b1;
            }
        }
    }
    return make_float2(psum32, osum32);
}

// FM channels of one capture share ONE discriminator. The reference mixes the whole capture down per channel and runs the
// discriminator on the unfiltered result (capture.py:298-369: freq_shift -> quadrature_demod, no channel filter in
// between), and angle(z[n] conj(z[n-1])) with z[n] = x[n] e^{j th[n]} is angle(x[n] conj(x[n-1])) + (th[n] - th[n-1]) mod
// 2 pi, where th[n] = fl32(k32 * fl32(n)) is the float32 phase numpy hands to exp() (its sample-to-sample increment jitters
// by ulp(th) around k32; both values are formed here exactly as numpy forms them and their float32 difference is exact). So
// the tile's discriminator d[n] is evaluated once (`dsh`), and a channel costs two multiplies, three adds and a wrap per sample
// instead of an oscillator, a complex product and an atan2. The two ways differ by the float32 roundings of z (~3e-7 rad);
// where that could decide which side of the branch cut a sample falls on (within 1e-4 rad of +-pi: 3 samples in 100 000),
// or where the product is exactly zero (d = NaN), the sample is redone in the reference's own order of operations.
// Returns this thread's partial of sum out^2.
// slow form: per-sample test for the branch cut / zero product, n = 0 -> 0 (quadrature_demod's out[0])
__device__ __noinline__ float front_tile_fm_shared_checked(const FrontArgs& a, const FrontChan& ch, const float2* tile,
                                                           const float* dsh, int t0, int cnt, long long obase, int tid) {
    const float PI_F = 3.14159265358979324f, TWO_PI_F = 6.28318530717958648f;
    float osum32 = 0.f;
    for (int i = tid; i < cnt; i += FR_THREADS) {
        const int n = t0 + i;
        float sft = __fadd_rn(dsh[i], __fsub_rn(__fmul_rn(ch.k32, (float)n), __fmul_rn(ch.k32, (float)(n - 1))));
        sft = (sft > PI_F) ? sft - TWO_PI_F : ((sft <= -PI_F) ? sft + TWO_PI_F : sft);
        if (!(fabsf(sft) < PI_F - 1e-4f)) {
            float c1, s1, c0, s0;
            nco_f32(ch.k32, n, c1, s1);
            nco_f32(ch.k32, n - 1, c0, s0);
            const float2 x1 = tile[i + 1], x0 = tile[i];
            const float2 b1 = make_float2(x1.x * c1 - x1.y * s1, x1.x * s1 + x1.y * c1);
            const float2 b0 = make_float2(x0.x * c0 - x0.y * s0, x0.x * s0 + x0.y * c0);
            sft = fast_atan2f_hi(b1.y * b0.x - b1.x * b0.y, b1.x * b0.x + b1.y * b0.y);
        }
        const float o = (n == 0) ? 0.0f : sft * ch.disc_scale;
        a.out[obase + n] = o;
        osum32 = fmaf(o, o, osum32);
    }
    return osum32;
}

// fast form: branch-free sample loop that only NOTES whether one of this thread's samples came within 1e-4 rad of the cut
// (or hit a zero product); such a thread (5 in 10 000), and the one that owns sample 0, redoes its 16 samples the slow way.
template <bool EXACT_IDX>
__device__ __forceinline__ float front_tile_fm_shared(const FrontArgs& a, const FrontChan& ch, const float2* tile, const float* dsh,
                                                      int t0, int cnt, long long obase, int tid) {
    const float PI_F = 3.14159265358979324f, TWO_PI_F = 6.28318530717958648f;
    float nf = (float)(t0 + tid);      // EXACT_IDX: the float32 sample index advances by exact float additions (no int -> float
    float osum32 = 0.f;                // conversions, a quarter-rate pipe, in the loop); beyond 2^24 it is converted per sample
    bool redo = (t0 == 0 && tid == 0);
    const float* dp = dsh + tid;
    float* op = a.out + obase + t0 + tid;
#pragma unroll 4
    for (int i = tid; i < cnt; i += FR_THREADS) {
        const float n1 = EXACT_IDX ? nf : (float)(t0 + i);
        const float n0 = EXACT_IDX ? nf - 1.0f : (float)(t0 + i - 1);
        float sft = __fadd_rn(*dp, __fsub_rn(__fmul_rn(ch.k32, n1), __fmul_rn(ch.k32, n0)));
        sft = (sft > PI_F) ? sft - TWO_PI_F : ((sft <= -PI_F) ? sft + TWO_PI_F : sft);
        redo |= !(fabsf(sft) < PI_F - 1e-4f);
        const float o = sft * ch.disc_scale;
        *op = o;
        osum32 = fmaf(o, o, osum32);
        dp += FR_THREADS;
        op += FR_THREADS;
        nf += (float)FR_THREADS;
    }
    if (redo) osum32 = front_tile_fm_shared_checked(a, ch, tile, dsh, t0, cnt, obase, tid);
    return osum32;
}

template <int KIND>
__device__ __forceinline__ float2 front_tile_dispatch(const FrontArgs& a, const FrontChan& ch, const float2* tile,
                                                     const float2* nco_tab, int t0, int cnt, long long obase, int tid) {
    if (ch.shift) {
        return a.base_out ? front_tile<KIND, true, true>(a, ch, tile, nco_tab, t0, cnt, obase, tid)
                          : front_tile<KIND, true, false>(a, ch, tile, nco_tab, t0, cnt, obase, tid);
    }
    return a.base_out ? front_tile<KIND, false, true>(a, ch, tile, nco_tab, t0, cnt, obase, tid)
                      : front_tile<KIND, false, false>(a, ch, tile, nco_tab, t0, cnt, obase, tid);
}

// Channel groups (blockIdx.z) for a front-end launch: enough CTAs for ~5 waves of the 4 resident per SM; when every channel
// takes the shared-discriminator path the per-CTA set-up (staging + one atan2 per sample) is the larger part of a CTA's work,
// so half as many groups (each CTA then serves more channels from one discriminator).
static int front_groups(long long base_ctas, int n_ch, bool all_fm_shared) {
    const long long want = (all_fm_shared ? 2LL : 4LL) * 5 * sm_count();
    long long groups = (want + base_ctas - 1) / base_ctas;
    if (groups > n_ch) groups = n_ch;
    if (groups < 1) groups = 1;
    return (int)groups;
}

struct FrontSmem {                       // 51.3 KB: dynamic shared memory (above the 48 KB static limit), 4 CTAs per SM
    double red[FR_THREADS / 32];
    double red2[FR_THREADS / 32];
    float2 tile[FR_TILE + 1];            // staged input, one halo sample in front
    float2 nco_tab[NCO_TAB];
    float dsh[FR_TILE];                  // discriminator of the un-shifted tile, shared by the shifted FM channels
};
static std::atomic<unsigned long long> g_front_optin{0};

__global__ void __launch_bounds__(FR_THREADS, 4) front_kernel(const FrontArgs a) {
    extern __shared__ __align__(16) unsigned char front_smem_raw[];
    FrontSmem& fsm = *reinterpret_cast<FrontSmem*>(front_smem_raw);
    float2* const tile = fsm.tile;
    float* const dsh = fsm.dsh;
    float2* const nco_tab = fsm.nco_tab;
    double* const red = fsm.red;
    double* const red2 = fsm.red2;
    for (int i = threadIdx.x; i < NCO_TAB; i += FR_THREADS) {
        float sv, cv;
        sincospif((float)i * (2.0f / (float)NCO_TAB), &sv, &cv);
        nco_tab[i] = make_float2(cv, sv);
    }
    const int chunk = blockIdx.y;
    const int t0 = blockIdx.x * FR_TILE;
    const int cnt = min(FR_TILE, a.n - t0);
    const int tid = threadIdx.x;
    // stage tile (+1 halo sample in front) as cf32
    bool bad = false;
    for (int i = tid; i < cnt + 1; i += FR_THREADS) {
        const int n = t0 - 1 + i;
        float2 v = make_float2(0.f, 0.f);
        if (n >= 0) {
            if (a.fmt == 0) {
                v = reinterpret_cast<const float2*>(a.iq)[(long long)chunk * a.chunk_stride + n];
            } else {
                const short2 q = reinterpret_cast<const short2*>(a.iq)[(long long)chunk * a.chunk_stride + n];
                v = make_float2((float)q.x / 32768.0f, (float)q.y / 32768.0f);
            }
            if (a.fmt == 0 && i > 0 && !(isfinite(v.x) && isfinite(v.y))) bad = true;   // int16 samples are always finite
        }
        tile[i] = v;
    }
    if (__syncthreads_or(bad) && tid == 0) atomicExch(a.nonfinite + chunk, 1);

    // shifted FM channels (no base output wanted) share one discriminator and one power sum of the un-shifted tile
    bool fm_shared = false;
    if (!a.base_out) {
        for (int c = blockIdx.z; c < a.n_ch; c += gridDim.z) {
            const int m = a.ch[c].mode;
            fm_shared |= (m == WC_MODE_WBFM || m == WC_MODE_NBFM) && a.ch[c].shift;
        }
    }
    double shared_pw = 0.0;        // sum |x|^2 of the tile (thread 0): the power of every shifted FM channel
    if (fm_shared) {
        float tile_pw = 0.f;
        for (int i = tid; i < cnt; i += FR_THREADS) {
            const float2 x1 = tile[i + 1], x0 = tile[i];
            const float pr = x1.x * x0.x + x1.y * x0.y, pi = x1.y * x0.x - x1.x * x0.y;
            dsh[i] = (pr == 0.0f && pi == 0.0f) ? __int_as_float(0x7fc00000) : fast_atan2f_hi(pi, pr);
            tile_pw += fmaf(x1.x, x1.x, x1.y * x1.y);       // |x e^{j th}|^2 = |x|^2 to one float32 rounding
        }
        const double w = warp_sum((double)tile_pw);
        if ((tid & 31) == 0) red[tid >> 5] = w;
        __syncthreads();
        if (tid == 0)
            for (int q = 0; q < FR_THREADS / 32; ++q) shared_pw += red[q];
        __syncthreads();
    }
    const bool exact_idx = (t0 + FR_TILE) <= (1 << 24);

    // blockIdx.z splits the channel loop when (tiles x chunks) alone would leave the last wave mostly empty
    for (int c = blockIdx.z; c < a.n_ch; c += gridDim.z) {
        const FrontChan ch = a.ch[c];
        const long long obase = ((long long)c * a.n_chunks + chunk) * a.n;
        float2 p32;
        if ((ch.mode == WC_MODE_WBFM || ch.mode == WC_MODE_NBFM) && ch.shift && !a.base_out) {
            const float o32 = exact_idx ? front_tile_fm_shared<true>(a, ch, tile, dsh, t0, cnt, obase, tid)
                                        : front_tile_fm_shared<false>(a, ch, tile, dsh, t0, cnt, obase, tid);
            if (a.out_sumsq) {          // only sum out^2 is this channel's own: <= 512 samples per warp in float32, float64 above
                float w32 = o32;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) w32 += __shfl_xor_sync(0xffffffffu, w32, off);
                if ((tid & 31) == 0) red2[tid >> 5] = (double)w32;
                __syncthreads();
            }
            if (tid == 0) {
                atomicAdd(a.power + (long long)c * a.n_chunks + chunk, shared_pw);
                if (a.out_sumsq) {
                    double s2 = 0.0;
                    for (int q = 0; q < FR_THREADS / 32; ++q) s2 += red2[q];
                    atomicAdd(a.out_sumsq + (long long)c * a.n_chunks + chunk, s2);
                }
            }
            if (a.out_sumsq) __syncthreads();
            continue;
        }
        if (ch.mode == WC_MODE_WBFM || ch.mode == WC_MODE_NBFM) p32 = front_tile_dispatch<1>(a, ch, tile, nco_tab, t0, cnt, obase, tid);
        else if (ch.mode == WC_MODE_AM) p32 = front_tile_dispatch<2>(a, ch, tile, nco_tab, t0, cnt, obase, tid);
        else if (ch.mode == WC_MODE_SSB) p32 = front_tile_dispatch<3>(a, ch, tile, nco_tab, t0, cnt, obase, tid);
        else p32 = front_tile_dispatch<0>(a, ch, tile, nco_tab, t0, cnt, obase, tid);
        const double psum = warp_sum((double)p32.x);
        const double osum = a.out_sumsq ? warp_sum((double)p32.y) : 0.0;
        if ((tid & 31) == 0) {
            red[tid >> 5] = psum;
            red2[tid >> 5] = osum;
        }
        __syncthreads();
        if (tid == 0) {
            double s = 0.0, s2 = 0.0;
            for (int w = 0; w < FR_THREADS / 32; ++w) {
                s += red[w];
                s2 += red2[w];
            }
            atomicAdd(a.power + (long long)c * a.n_chunks + chunk, s);
            if (a.out_sumsq) atomicAdd(a.out_sumsq + (long long)c * a.n_chunks + chunk, s2);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// IIR block scan (float64 DF2T segments, double-double chaining)
// ---------------------------------------------------------------------------------------------
// tf-form Butterworth filters with poles clustered near z=1 (100 Hz high-pass at 48 kS/s, order-10
// band-pass) have transition matrices A^64 with entries ~1e6..1e9 and massive cancellation in
// P*s + z. Chaining segment states in plain float64 amplifies rounding twice and produces garbage;
// with the transition powers and the chained state in double-double the scan reproduces the
// sequential float64 recursion to its own rounding level (validated with exact rationals,
// DESIGN.md "IIR scan").
constexpr int IIR_L = 64;    // samples per segment
constexpr int IIR_T = 128;   // segments per tile (one per thread)
constexpr int IIR_TILE = IIR_L * IIR_T;
constexpr int IIR_KMAX = 10;
constexpr int IIR_LOG_T = 7;

struct dd {
    double hi, lo;
};
__host__ __device__ __forceinline__ dd dd_two_sum(double a, double b) {
    const double s = a + b;
    const double bb = s - a;
    return dd{s, (a - (s - bb)) + (b - bb)};
}
__host__ __device__ __forceinline__ dd dd_quick(double a, double b) {
    const double s = a + b;
    return dd{s, b - (s - a)};
}
__host__ __device__ __forceinline__ dd dd_add(dd a, dd b) {
    dd s = dd_two_sum(a.hi, b.hi);
    const dd t = dd_two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = dd_quick(s.hi, s.lo);
    s.lo += t.lo;
    return dd_quick(s.hi, s.lo);
}
__host__ __device__ __forceinline__ dd dd_mul(dd a, dd b) {
    const double p = a.hi * b.hi;
    double e = fma(a.hi, b.hi, -p);
    e = fma(a.hi, b.lo, e);
    e = fma(a.lo, b.hi, e);
    return dd_quick(p, e);
}

struct IirCoef {
    int K;                                   // order
    double b0;
    double a[IIR_KMAX];                      // a[i+1]
    double b[IIR_KMAX];                      // b[i+1]
    dd Ppow[IIR_LOG_T][IIR_KMAX * IIR_KMAX]; // A^(L*2^k), k = 0..6 (segment transition powers)
    dd PT[IIR_KMAX * IIR_KMAX];              // A^(L*T)    (tile transition)
};

template <int K>
__device__ __forceinline__ double df2t_step(const double b0, const double (&cb)[K], const double (&ca)[K],
                                            double (&z)[K], double x) {
    const double y = fma(b0, x, z[0]);
#pragma unroll
    for (int i = 0; i < K - 1; ++i) z[i] = fma(cb[i], x, fma(-ca[i], y, z[i + 1]));
    z[K - 1] = fma(cb[K - 1], x, -ca[K - 1] * y);
    return y;
}

// o = M * s + o   (row-major K x K double-double). PLAIN: float64 only (the .hi parts) — enough for well-conditioned
// filters, 20 x fewer FP64 operations per term; chosen per handle at create time (iir_scan_plain_ok).
template <int K, bool PLAIN>
__device__ __forceinline__ void dd_matvec_acc(const dd* __restrict__ M, const dd* s, dd* o) {
    for (int r = 0; r < K; ++r) {
        dd acc = o[r];
        if (PLAIN) {
            for (int c = 0; c < K; ++c) acc.hi = fma(M[r * K + c].hi, s[c].hi, acc.hi);
        } else {
            // compensated dot product (Dot2 style): exact product error by FMA, the running high part by two_sum, all low-order
            // terms collected in one float64 and folded back once per row — double-double accuracy for the K-term sum at 12
            // FP64 operations per term instead of 28 for a full dd_mul + dd_add (the scan is ~6/7 of this kernel's FP64 work)
            double hi = acc.hi, lo = acc.lo;
            for (int c = 0; c < K; ++c) {
                const dd m = M[r * K + c], v = s[c];
                const double p = m.hi * v.hi;
                double e = fma(m.hi, v.hi, -p);
                e = fma(m.hi, v.lo, e);
                e = fma(m.lo, v.hi, e);
                const dd t = dd_two_sum(hi, p);
                hi = t.hi;
                lo += t.lo + e;
            }
            acc = dd_quick(hi, lo);
        }
        o[r] = acc;
    }
}

// Dynamic shared memory layout: float sx[T*(L+1)] | dd vb[2][T][K] | dd sP[LOG_T][K*K]
template <int K>
constexpr size_t iir_smem_bytes() {
    return sizeof(float) * IIR_T * (IIR_L + 1) + sizeof(dd) * 2 * IIR_T * K + sizeof(dd) * IIR_LOG_T * K * K + 16;
}

// FINAL = 0 (pass 1): zero-state run of every 64-sample segment -> z_i (global), inclusive scan
//   -> tile aggregate Z_c (double-double).
// FINAL = 1 (pass 3): scan again with the true tile-start state S_c folded into segment 0, then
//   re-run every segment from its true start state and write y.
// sumsq (FINAL only, optional): per-sequence sum of the float32 outputs squared, for rms_normalize right after the
// last stage of a chain (saves the separate pass over the filtered signal).
template <int K, int FINAL, bool PLAIN, typename TIn>
__global__ void __launch_bounds__(IIR_T) iir_kernel(const IirCoef* __restrict__ cf, const TIn* __restrict__ x,
                                                    float* __restrict__ y, int n, long long seq_stride,
                                                    double* __restrict__ zseg, dd* __restrict__ ztile,
                                                    const dd* __restrict__ stile, int tiles, int absin,
                                                    double* __restrict__ sumsq) {
    extern __shared__ __align__(16) unsigned char iir_smem[];
    float* sx = reinterpret_cast<float*>(iir_smem);
    dd* vb = reinterpret_cast<dd*>(iir_smem + ((sizeof(float) * IIR_T * (IIR_L + 1) + 15) & ~size_t(15)));
    // transition powers staged once per CTA: the scan reads K*K of them per step and thread, and from global memory
    // every one of those was a long-scoreboard stall (ncu round 1: 10.3 stalled warps per issue at 21 % of the warp slots)
    dd* sP = vb + 2 * IIR_T * K;
    const int tid = threadIdx.x;
    for (int e = tid; e < IIR_LOG_T * K * K; e += IIR_T) sP[e] = cf->Ppow[e / (K * K)][e % (K * K)];
    const int tile = blockIdx.x, seq = blockIdx.y;
    const int t0 = tile * IIR_TILE;
    const int cnt = min(IIR_TILE, n - t0);
    const TIn* xs = x + (long long)seq * seq_stride + t0;
    for (int e = tid; e < IIR_TILE; e += IIR_T) {
        float v = 0.f;
        if (e < cnt) {
            v = (float)xs[e];
            if (absin) v = fabsf(v);
        }
        sx[(e >> 6) * (IIR_L + 1) + (e & 63)] = v;
    }
    double cb[K], ca[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        cb[i] = cf->b[i];
        ca[i] = cf->a[i];
    }
    const double b0 = cf->b0;
    __syncthreads();
    const int nseg = (cnt + IIR_L - 1) / IIR_L;
    const long long tix = (long long)seq * tiles + tile;
    const long long zoff = tix * IIR_T * K;
    float* row = sx + tid * (IIR_L + 1);
    double z[K];
    dd v[K];
    if (!FINAL) {
#pragma unroll
        for (int i = 0; i < K; ++i) z[i] = 0.0;
        if (tid < nseg) {
#pragma unroll 4
            for (int j = 0; j < IIR_L; ++j) df2t_step<K>(b0, cb, ca, z, (double)row[j]);
        }
#pragma unroll
        for (int i = 0; i < K; ++i) zseg[zoff + (long long)tid * K + i] = z[i];
    } else {
#pragma unroll
        for (int i = 0; i < K; ++i) z[i] = zseg[zoff + (long long)tid * K + i];
    }
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = dd{z[i], 0.0};
    dd sc[K];
    if (FINAL) {
#pragma unroll
        for (int i = 0; i < K; ++i) sc[i] = stile[tix * K + i];
        if (tid == 0) dd_matvec_acc<K, PLAIN>(sP, sc, v);  // end(seg 0) = P*S_c + z_0
    }
    // Kogge-Stone inclusive scan: v_i <- v_i + P^(2^k) * v_{i-2^k}
    int cur = 0;
#pragma unroll 1
    for (int k = 0; k < IIR_LOG_T; ++k) {
        dd* buf = vb + (size_t)cur * IIR_T * K;
#pragma unroll
        for (int i = 0; i < K; ++i) buf[tid * K + i] = v[i];
        __syncthreads();
        const int d = 1 << k;
        if (tid >= d) dd_matvec_acc<K, PLAIN>(sP + k * K * K, buf + (tid - d) * K, v);
        cur ^= 1;
    }
    if (!FINAL) {
        if (tid == IIR_T - 1) {
#pragma unroll
            for (int i = 0; i < K; ++i) ztile[tix * K + i] = v[i];
        }
        return;
    }
    // exclusive: start state of segment i is the (true) end state of segment i-1
    dd* buf = vb + (size_t)cur * IIR_T * K;
#pragma unroll
    for (int i = 0; i < K; ++i) buf[tid * K + i] = v[i];
    __syncthreads();
    if (tid < nseg) {
#pragma unroll
        for (int i = 0; i < K; ++i) z[i] = (tid == 0) ? sc[i].hi : buf[(tid - 1) * K + i].hi;
#pragma unroll 4
        for (int j = 0; j < IIR_L; ++j) row[j] = (float)df2t_step<K>(b0, cb, ca, z, (double)row[j]);
    }
    __syncthreads();
    float* ys = y + (long long)seq * seq_stride + t0;
    double ss = 0.0;
    for (int e = tid; e < cnt; e += IIR_T) {
        const float v = sx[(e >> 6) * (IIR_L + 1) + (e & 63)];
        ys[e] = v;
        ss += (double)(v * v);   // x**2 is float32 in the reference (dsp/fm.py:58)
    }
    if (sumsq) {
        ss = warp_sum(ss);
        if ((tid & 31) == 0) atomicAdd(sumsq + seq, ss);
    }
}

// pass 2: chain tile aggregates in double-double: S_0 = 0, S_{c+1} = PT * S_c + Z_c.
// One warp per sequence, lane r computes row r.
template <int K, bool PLAIN>
__global__ void iir_chain_kernel(const IirCoef* __restrict__ cf, const dd* __restrict__ ztile,
                                 dd* __restrict__ stile, int tiles, int n_seq) {
    __shared__ dd s_sh[4][IIR_KMAX];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seq = blockIdx.x * 4 + w;
    if (seq >= n_seq) return;
    dd s = dd{0.0, 0.0};
    for (int c = 0; c < tiles; ++c) {
        const long long off = ((long long)seq * tiles + c) * K;
        if (lane < K) {
            stile[off + lane] = s;
            s_sh[w][lane] = s;
        }
        __syncwarp();
        if (lane < K) {
            dd acc = ztile[off + lane];
            if (PLAIN) {
                for (int j = 0; j < K; ++j) acc.hi = fma(cf->PT[lane * K + j].hi, s_sh[w][j].hi, acc.hi);
            } else {
                for (int j = 0; j < K; ++j) acc = dd_add(acc, dd_mul(cf->PT[lane * K + j], s_sh[w][j]));
            }
            s = acc;
        }
        __syncwarp();
    }
}

// Exact replay of scipy.signal.lfilter's float64 recursion (scipy/signal/_lfilter.c.in: y = z0 + b0*x;
// z[i] = z[i+1] + x*b[i+1] - y*a[i+1]; products and sums rounded separately — x86-64 wheels are built without FMA
// contraction, and a numpy replay of exactly these operations is bit-equal to lfilter, tests/test_oracle_analog.py).
// For tf-form filters whose transition matrix has large transient growth (the order-10 band-pass of ssb_demod,
// dsp/filters.py:177-217: |A^64| ~ 2e9) a block scan that hands each segment an independently rounded start state
// perturbs the state off the manifold the sequential recursion stays on, and that perturbation is amplified to
// ~5e-4 of the output; replaying the recursion reproduces the reference's own rounding bit for bit instead.
// One lane per sequence, a warp stages 32 sequences x 64 samples through shared memory so global traffic is coalesced.
constexpr int IIRS_SEG = 64;
template <int K>
__global__ void __launch_bounds__(32) iir_seq_kernel(const IirCoef* __restrict__ cf, const float* __restrict__ x,
                                                     float* __restrict__ y, int n, long long seq_stride, int n_seq,
                                                     int absin) {
    __shared__ float tile[2][32][IIRS_SEG + 1];
    const int lane = threadIdx.x;
    const int seq0 = blockIdx.x * 32;
    const int rows = min(32, n_seq - seq0);
    double cb[K], ca[K], z[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        cb[i] = cf->b[i];
        ca[i] = cf->a[i];
        z[i] = 0.0;
    }
    const double b0 = cf->b0;
    // segment t0 of every row -> tile[b], as asynchronous copies: the next segment lands while this one is filtered
    auto stage = [&](int b, int t0) {
        const int cnt = min(IIRS_SEG, n - t0);
        for (int r = 0; r < rows; ++r) {
            const float* xs = x + (long long)(seq0 + r) * seq_stride + t0;
            for (int e = lane; e < cnt; e += 32) cp_async4(&tile[b][r][e], xs + e);
        }
        cp_async_commit();
    };
    if (n > 0) stage(0, 0);
    int b = 0;
    for (int t0 = 0; t0 < n; t0 += IIRS_SEG, b ^= 1) {
        const int cnt = min(IIRS_SEG, n - t0);
        if (t0 + IIRS_SEG < n) {
            stage(b ^ 1, t0 + IIRS_SEG);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        if (lane < rows) {
#pragma unroll 2
            for (int j = 0; j < cnt; ++j) {
                const float xf = tile[b][lane][j];
                const double xv = (double)(absin ? fabsf(xf) : xf);
                const double yv = __dadd_rn(z[0], __dmul_rn(b0, xv));
#pragma unroll
                for (int i = 0; i < K - 1; ++i)
                    z[i] = __dsub_rn(__dadd_rn(z[i + 1], __dmul_rn(xv, cb[i])), __dmul_rn(yv, ca[i]));
                z[K - 1] = __dsub_rn(__dmul_rn(xv, cb[K - 1]), __dmul_rn(yv, ca[K - 1]));
                tile[b][lane][j] = (float)yv;
            }
        }
        __syncwarp();
        if (cnt == IIRS_SEG) {          // full segment: four rows' shared-memory loads in flight per round of stores
            float* yb = y + (long long)seq0 * seq_stride + t0 + lane;
            int r = 0;
            for (; r + 4 <= rows; r += 4) {
                float v[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    v[2 * q] = tile[b][r + q][lane];
                    v[2 * q + 1] = tile[b][r + q][lane + 32];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    yb[(long long)(r + q) * seq_stride] = v[2 * q];
                    yb[(long long)(r + q) * seq_stride + 32] = v[2 * q + 1];
                }
            }
            for (; r < rows; ++r) {
                yb[(long long)r * seq_stride] = tile[b][r][lane];
                yb[(long long)r * seq_stride + 32] = tile[b][r][lane + 32];
            }
        } else {
            for (int r = 0; r < rows; ++r) {
                float* ys = y + (long long)(seq0 + r) * seq_stride + t0;
                for (int e = lane; e < cnt; e += 32) ys[e] = tile[b][r][e];
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// reductions / elementwise
// ---------------------------------------------------------------------------------------------
__global__ void sumsq_kernel(const float* __restrict__ x, int n, long long seq_stride, double* __restrict__ out) {
    __shared__ double red[8];
    const int seq = blockIdx.y;
    const float* xs = x + (long long)seq * seq_stride;
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float v = xs[i];
        acc += (double)(v * v);  // x**2 is float32 in the reference
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        atomicAdd(out + seq, s);
    }
}

__device__ __forceinline__ float soft_clip_fm(float x) {  // dsp/fm.py:26-39
    return tanhf(x * 1.5f) * 1.1047914f * 0.95f;           // float32(1/tanh(1.5)) = 1.1047914
}
__device__ __forceinline__ float soft_clip_agc(float x) {  // dsp/agc.py:58-70
    return tanhf(x * 1.5f) * 1.1047914f;
}

// apply_agc tail (dsp/agc.py:228-242): gain = min(target / max(max(env_a, env_r), 1e-6), max_gain)
__global__ void agc_apply_kernel(const float* __restrict__ x, const float* __restrict__ env_a,
                                 const float* __restrict__ env_r, float* __restrict__ y, long long total,
                                 float target_lin, float max_gain) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float env = fmaxf(env_a[i], env_r[i]);
        const float g = fminf(target_lin / fmaxf(env, 1e-6f), max_gain);
        y[i] = soft_clip_agc(x[i] * g);
    }
}

__global__ void elementwise_kernel(const float* __restrict__ x, float* __restrict__ y, long long total, int op,
                                   float p0) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float v = x[i];
        float o = v;
        if (op == 0) o = soft_clip_fm(v);
        else if (op == 1) o = soft_clip_agc(v);
        else if (op == 2) o = v * p0;
        y[i] = o;
    }
}

// ---------------------------------------------------------------------------------------------
// polyphase resampler (scipy.signal.resample_poly / upfirdn index math)
// ---------------------------------------------------------------------------------------------
struct ResampArgs {
    const float* x;          // [n_seq][seq_stride]
    long long seq_stride;
    int n_in, n_out;
    int up, down;
    long long q0;            // Q(m) = (m + n_pre_remove)*down - n_pre_pad = q0 + m*down
    int ntaps;               // len(h) (unpadded)
    int tpp;                 // taps per phase (ceil(ntaps/up))
    const double* hp;        // [up][tpp] polyphase taps: hp[phi][i] = h[phi + up*i] (0 past the end)
    const double* g;         // residue form: [up][E][down] taps g[mu][j][r0] = h[B(mu,r0) + up*down*(e_lo[r0] + j)] (0 outside)
    const float* gf;         // the same table rounded to float32 (mixed-precision residue kernel)
    const int* e_lo;         // [down] first tap offset of residue r0 (the union over mu of its tap ranges starts here)
    int E;                   // padded taps per (mu, r0)
    float* out;              // [n_seq][n_out]
    // epilogue
    int epi;                 // 0 none, 1 rms-scale + fm soft clip, 2 fm soft clip, 3 rms-scale only, 4 agc soft clip
    const double* sumsq;     // [n_seq] for epi 1/3 (rms = sqrt(sumsq / n_in))
    float target_rms, min_rms;
    double* power;           // optional [n_seq]: sum out^2
    int* invalid;            // optional [n_seq]: set when an output is non-finite or |x| > max_abs
    float max_abs;
};

constexpr int RS_WARPS = 8;
constexpr int RS_R = 8;   // outputs of one polyphase branch per warp task

// One warp computes RS_R outputs that use the SAME polyphase branch (m = m0 + up*k: the branch index repeats every `up`
// outputs once up/down are reduced), lanes striding over the branch's taps: every tap is loaded once per RS_R
// multiply-adds instead of once per multiply-add. With one output per warp the kernel was bound by L1 bandwidth
// (12 B per float64 FMA); this form moves 5 B.
__global__ void __launch_bounds__(RS_WARPS * 32) resample_kernel(const ResampArgs a) {
    const int seq = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* xs = a.x + (long long)seq * a.seq_stride;
    float scale = 1.0f;
    if (a.epi == 1 || a.epi == 3) {
        const float rms = (float)sqrt(a.sumsq[seq] / (double)a.n_in);
        if (rms > a.min_rms) scale = (float)((double)a.target_rms / (double)rms);
    }
    double pw = 0.0;
    bool bad = false;
    // tasks: (branch residue m0 in [0, up), group g of RS_R outputs m0 + up*(RS_R*g + r))
    const int per_branch = (a.n_out + a.up - 1) / a.up;            // upper bound of outputs per residue
    const int groups = (per_branch + RS_R - 1) / RS_R;
    const long long n_tasks = (long long)groups * a.up;
    for (long long task = (long long)blockIdx.x * RS_WARPS + warp; task < n_tasks; task += (long long)gridDim.x * RS_WARPS) {
        const int m0 = (int)(task % a.up);
        const int g = (int)(task / a.up);
        const int mfirst = m0 + a.up * (RS_R * g);
        if (mfirst >= a.n_out) continue;
        const long long Q0 = a.q0 + (long long)mfirst * a.down;
        long long phi = Q0 % a.up;
        if (phi < 0) phi += a.up;
        const double* h = a.hp + phi * a.tpp;
        // output r: Q_r = Q0 + r*up*down -> nmax_r = nmax_0 + r*down (same branch)
        const long long nmax0 = (Q0 - phi) / a.up;
        int nr = 0;          // valid outputs in this task
        while (nr < RS_R && mfirst + a.up * nr < a.n_out) ++nr;
        double acc[RS_R];
#pragma unroll
        for (int r = 0; r < RS_R; ++r) acc[r] = 0.0;
        // interior: every output of the task has all of [i_in_lo, i_in_hi] inside its input range
        // valid i for output r: max(0, nmax_r - n_in + 1) <= i <= min(tpp - 1, nmax_r)
        const long long nmax_last = nmax0 + (long long)(nr - 1) * a.down;
        long long lo = nmax_last - a.n_in + 1;
        if (lo < 0) lo = 0;
        long long hi = (nmax0 < a.tpp - 1) ? nmax0 : a.tpp - 1;
        if (nr == RS_R && lo <= hi) {
            for (int i = (int)lo + lane; i <= (int)hi; i += 32) {
                const double hv = h[i];
                const float* xp = xs + (nmax0 - i);
#pragma unroll
                for (int r = 0; r < RS_R; ++r) acc[r] = fma(hv, (double)xp[(long long)r * a.down], acc[r]);
            }
            // edges of the individual outputs outside the common range
#pragma unroll
            for (int r = 0; r < RS_R; ++r) {
                const long long nm = nmax0 + (long long)r * a.down;
                long long l = nm - a.n_in + 1;
                if (l < 0) l = 0;
                const long long hh = (nm < a.tpp - 1) ? nm : a.tpp - 1;
                for (long long i = l + lane; i < lo; i += 32) acc[r] = fma(h[i], (double)xs[nm - i], acc[r]);
                for (long long i = hi + 1 + lane; i <= hh; i += 32) acc[r] = fma(h[i], (double)xs[nm - i], acc[r]);
            }
        } else {
            for (int r = 0; r < nr; ++r) {
                const long long nm = nmax0 + (long long)r * a.down;
                long long l = nm - a.n_in + 1;
                if (l < 0) l = 0;
                const long long hh = (nm < a.tpp - 1) ? nm : a.tpp - 1;
                for (long long i = l + lane; i <= hh; i += 32) acc[r] = fma(h[i], (double)xs[nm - i], acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < RS_R; ++r) {
            const double sum = warp_sum(acc[r]);
            if (lane == 0 && r < nr) {
                float v = (float)sum;  // resample_poly(...).astype(float32)
                if (a.epi == 1) v = soft_clip_fm((float)(sum * (double)scale));
                else if (a.epi == 2) v = soft_clip_fm(v);
                else if (a.epi == 3) v = (float)(sum * (double)scale);
                else if (a.epi == 4) v = soft_clip_agc(v);
                a.out[(long long)seq * a.n_out + (mfirst + a.up * r)] = v;
                pw += (double)(v * v);
                if (!isfinite(v) || fabsf(v) > a.max_abs) bad = true;
            }
        }
    }
    if (lane == 0) {
        if (a.power && pw != 0.0) atomicAdd(a.power + seq, pw);
        if (a.invalid && bad) atomicExch(a.invalid + seq, 1);
    }
}

// one output per warp: used when the outputs of a branch are far apart in the input (large `down`), where the grouped
// form above gains nothing — that case is bound by the float64 pipe (DFMA + the f32->f64 conversions), not by L1
__global__ void __launch_bounds__(RS_WARPS * 32) resample_single_kernel(const ResampArgs a) {
    const int seq = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* xs = a.x + (long long)seq * a.seq_stride;
    float scale = 1.0f;
    if (a.epi == 1 || a.epi == 3) {
        const float rms = (float)sqrt(a.sumsq[seq] / (double)a.n_in);
        if (rms > a.min_rms) scale = (float)((double)a.target_rms / (double)rms);
    }
    double pw = 0.0;
    bool bad = false;
    for (int m = blockIdx.x * RS_WARPS + warp; m < a.n_out; m += gridDim.x * RS_WARPS) {
        const long long Q = a.q0 + (long long)m * a.down;
        // taps t = Q - n*up in [0, ntaps); phase phi = Q mod up; n = nmax - i with i = (t - phi)/up
        long long phi = Q % a.up;
        if (phi < 0) phi += a.up;
        const long long nmax = (Q - phi) / a.up;
        const double* h = a.hp + phi * a.tpp;
        // valid i: 0 <= i < tpp_phi, 0 <= nmax - i < n_in
        int i_lo = (nmax >= a.n_in) ? (int)(nmax - a.n_in + 1) : 0;
        long long i_hi_ll = nmax;  // inclusive upper bound from n >= 0
        int i_hi = (i_hi_ll >= a.tpp) ? a.tpp - 1 : (int)i_hi_ll;
        double acc = 0.0;
        for (int i = i_lo + lane; i <= i_hi; i += 32) acc = fma(h[i], (double)xs[nmax - i], acc);
        acc = warp_sum(acc);
        if (lane == 0) {
            float v = (float)acc;  // resample_poly(...).astype(float32)
            if (a.epi == 1) v = soft_clip_fm((float)(acc * (double)scale));
            else if (a.epi == 2) v = soft_clip_fm(v);
            else if (a.epi == 3) v = (float)(acc * (double)scale);
            else if (a.epi == 4) v = soft_clip_agc(v);
            a.out[(long long)seq * a.n_out + m] = v;
            pw += (double)(v * v);
            if (!isfinite(v) || fabsf(v) > a.max_abs) bad = true;
        }
    }
    if (lane == 0) {
        if (a.power && pw != 0.0) atomicAdd(a.power + seq, pw);
        if (a.invalid && bad) atomicExch(a.invalid + seq, 1);
    }
}

// Residue form of the same sum, for decimating ratios (down >> up), where it removes the per-tap loads entirely.
// With m = up*a + mu and n = r0 + down*k (r0 in [0, down)), the tap of (m, n) is h[B(mu, r0) + up*down*(a - k)],
// B = q0 + down*mu - up*r0: for a fixed residue r0 and output phase mu the resampler is a SHORT FIR along k — E ~
// ntaps/(up*down) taps (7 for 3/625, 21 for 1/50) — and the output is the sum of those FIRs over the residues:
//   y[up*a + mu] = sum_{r0} sum_{j<E} g[mu][j][r0] * x[r0 + down*(a - e_lo[r0] - j)].
// A thread owns residues r0 = tid, tid + T, ...; per residue it holds the up*E taps and the RS_A + E - 1 inputs of an
// RS_A-output block in registers (lanes run along r0: every load is coalesced) and issues up*RS_A*E float64 FMAs for
// up*E + RS_A + E - 1 loads (3/625: 168 FMAs per 35 loads; the one-output-per-warp form needs two loads per FMA and
// was bound by L1 bandwidth). Float64 accumulation as scipy does; the per-thread partial sums are reduced across the
// CTA once per block of up*RS_A outputs.
constexpr int RS_A = 8;          // outputs per phase per block of outputs
constexpr int RS_RT = 128;       // threads per CTA (residue owners)
constexpr int RS_NB = 1;         // output blocks a thread group walks in sequence (measured on B200: 1, 2, 4 within 4 %)

// Thread groups: with down <= 64 residues a CTA splits into G = 2 or 4 groups of 64 / 32 threads, each walking its own
// output blocks, so small decimation factors (1/50) do not leave most of the CTA idle.
// T = double: float64 products and sums like scipy. T = float (default): a thread's partial sum — its ~down/128
// residues x E taps, 40 terms for 3/625 — is formed in float32 (FFMA at 8x the rate of this part's FP64 pipe, no
// conversions) and everything after that, i.e. the sum over the 128 threads, is float64. Measured against scipy's
// float64 result: 3/625 and 1/50 outputs differ by a few 1e-8 relative RMS either way — the float32 rounding of the
// OUTPUT, which the reference applies too (`.astype(np.float32)`), not the accumulation.
template <int UP, int E, typename T>
__global__ void __launch_bounds__(RS_RT, (sizeof(T) == 4 ? 6 : 3)) resample_residue_kernel(const ResampArgs a, int n_blocks, int G, int NB) {
    constexpr int W = RS_A + E - 1;
    __shared__ double part[RS_RT / 32][UP * RS_A];
    const int seq = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gt = RS_RT / G;                 // threads per group
    const int grp = tid / gt, gtid = tid - grp * gt;
    const int wpg = gt >> 5;                  // warps per group
    const float* xs = a.x + (long long)seq * a.seq_stride;
    float scale = 1.0f;
    if (a.epi == 1 || a.epi == 3) {
        const float rms = (float)sqrt(a.sumsq[seq] / (double)a.n_in);
        if (rms > a.min_rms) scale = (float)((double)a.target_rms / (double)rms);
    }
    double pw = 0.0;
    bool bad = false;
    for (int nb = 0; nb < NB; ++nb) {
        const int blk = (blockIdx.x * G + grp) * NB + nb;       // may run past n_blocks: computes nothing, keeps the barriers
        const int a0 = blk * RS_A;
        T pacc[UP][RS_A];
#pragma unroll
        for (int mu = 0; mu < UP; ++mu)
#pragma unroll
            for (int i = 0; i < RS_A; ++i) pacc[mu][i] = (T)0;
        if (blk < n_blocks) {
            const T* gt_tab = (sizeof(T) == 8) ? reinterpret_cast<const T*>(a.g) : reinterpret_cast<const T*>(a.gf);
            for (int r0 = gtid; r0 < a.down; r0 += gt) {
                const long long kbase = (long long)a0 - __ldg(a.e_lo + r0) - (E - 1);
                T w[W];
                const long long n_first = r0 + (long long)a.down * kbase;
                if (n_first >= 0 && n_first + (long long)a.down * (W - 1) < a.n_in) {
                    // interior block (all but the first and last few): no bounds tests, 32-bit strides
                    const float* xp = xs + n_first;
#pragma unroll
                    for (int q = 0; q < W; ++q) w[q] = (T)xp[q * a.down];
                } else {
#pragma unroll
                    for (int q = 0; q < W; ++q) {
                        const long long n = n_first + (long long)a.down * q;
                        w[q] = (n >= 0 && n < a.n_in) ? (T)xs[n] : (T)0;     // zero extension, as upfirdn
                    }
                }
#pragma unroll
                for (int mu = 0; mu < UP; ++mu) {
                    T g[E];
#pragma unroll
                    for (int j = 0; j < E; ++j) g[j] = __ldg(gt_tab + (mu * E + j) * a.down + r0);   // table < 2^31 entries
#pragma unroll
                    for (int j = 0; j < E; ++j)
#pragma unroll
                        for (int i = 0; i < RS_A; ++i) pacc[mu][i] = fma(g[j], w[E - 1 + i - j], pacc[mu][i]);
                }
            }
        }
        double acc[UP][RS_A];
#pragma unroll
        for (int mu = 0; mu < UP; ++mu)
#pragma unroll
            for (int i = 0; i < RS_A; ++i) acc[mu][i] = (double)pacc[mu][i];
        // warp reduction by recursive halving: three exchange steps leave every lane with UP of the UP*RS_A sums over
        // its group of 8 lanes (sum index = UP * (4*b0 + 2*b1 + b2) + k, b = bits of the lane), two plain steps finish
        // them over the 4 groups: 6.75 shuffles per value instead of 10 x 24
        {
            double* v = &acc[0][0];
            constexpr int N0 = UP * RS_A;
#pragma unroll
            for (int st = 0; st < 3; ++st) {
                const int half = N0 >> (st + 1);
                const bool up_half = (lane >> st) & 1;
#pragma unroll
                for (int k = 0; k < half; ++k) {
                    const double send = up_half ? v[k] : v[k + half];
                    const double keep = up_half ? v[k + half] : v[k];
                    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 1 << st);
                }
            }
#pragma unroll
            for (int k = 0; k < UP; ++k) {
                double t = v[k];
                t += __shfl_xor_sync(0xffffffffu, t, 8);
                t += __shfl_xor_sync(0xffffffffu, t, 16);
                if (lane < 8) part[warp][UP * (4 * (lane & 1) + 2 * ((lane >> 1) & 1) + ((lane >> 2) & 1)) + k] = t;
            }
        }
        __syncthreads();
        if (gtid < UP * RS_A && blk < n_blocks) {
            double sum = 0.0;
            for (int wv = 0; wv < wpg; ++wv) sum += part[grp * wpg + wv][gtid];
            const int mu = gtid / RS_A, i = gtid % RS_A;
            const long long m = (long long)UP * (a0 + i) + mu;
            if (m < a.n_out) {
                float v = (float)sum;  // resample_poly(...).astype(float32)
                if (a.epi == 1) v = soft_clip_fm((float)(sum * (double)scale));
                else if (a.epi == 2) v = soft_clip_fm(v);
                else if (a.epi == 3) v = (float)(sum * (double)scale);
                else if (a.epi == 4) v = soft_clip_agc(v);
                a.out[(long long)seq * a.n_out + m] = v;
                pw += (double)(v * v);
                bad = bad || !isfinite(v) || fabsf(v) > a.max_abs;
            }
        }
        __syncthreads();
    }
    // the writers are the first UP * RS_A <= 32 threads of every group: lane-0 warps of the groups hold pw / bad
    if (gtid < 32) {
        pw = warp_sum(pw);
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) {
            if (a.power && pw != 0.0) atomicAdd(a.power + seq, pw);
            if (a.invalid && bad) atomicExch(a.invalid + seq, 1);
        }
    }
}

// squelch / validity select (capture.py:2918-2921, 147-162): zero or flag audio per sequence
__global__ void finalize_kernel(float* __restrict__ audio, int n_out, int n_chunks, const double* __restrict__ power_iq,
                                int n_in, const float* __restrict__ squelch_db, const int* __restrict__ has_squelch,
                                float* __restrict__ rssi_db, unsigned char* __restrict__ squelched) {
    const int seq = blockIdx.y;
    const int c = seq / n_chunks;  // sequences are channel-major: seq = c*n_chunks + chunk
    const float rssi = (float)(10.0 * log10(power_iq[seq] / (double)n_in + 1e-10));
    const bool sq = has_squelch[c] && (rssi < squelch_db[c]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        rssi_db[seq] = rssi;
        squelched[seq] = sq ? 1 : 0;
    }
    if (sq)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += gridDim.x * blockDim.x)
            audio[(long long)seq * n_out + i] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------
static void dd_matmul(const dd* A, const dd* B, dd* C, int K) {
    std::vector<dd> t((size_t)K * K);
    for (int r = 0; r < K; ++r)
        for (int c = 0; c < K; ++c) {
            dd acc = dd{0.0, 0.0};
            for (int k = 0; k < K; ++k) acc = dd_add(acc, dd_mul(A[r * K + k], B[k * K + c]));
            t[r * K + c] = acc;
        }
    for (int i = 0; i < K * K; ++i) C[i] = t[i];
}

// Normalise (b, a) like scipy.signal.lfilter (divide by a[0], pad to equal length), build the DF2T
// state matrix A (z' = A z + B x) and its powers A^(64*2^k) in double-double by repeated squaring.
static int make_iir_coef(const double* b, int nb, const double* a, int na, IirCoef* out) {
    if (na < 1 || nb < 1 || a[0] == 0.0) return -1;
    int K = (nb > na ? nb : na) - 1;
    if (K > IIR_KMAX) return -1;
    memset(out, 0, sizeof(*out));
    std::vector<double> bn(K + 1, 0.0), an(K + 1, 0.0);
    for (int i = 0; i < nb; ++i) bn[i] = b[i] / a[0];
    for (int i = 0; i < na; ++i) an[i] = a[i] / a[0];
    out->b0 = bn[0];
    out->K = K;
    if (K == 0) return 0;
    for (int i = 0; i < K; ++i) {
        out->a[i] = an[i + 1];
        out->b[i] = bn[i + 1];
    }
    std::vector<dd> P((size_t)K * K, dd{0.0, 0.0});
    for (int i = 0; i < K; ++i) {
        P[i * K + 0] = dd{-an[i + 1], 0.0};
        if (i + 1 < K) P[i * K + i + 1] = dd_add(P[i * K + i + 1], dd{1.0, 0.0});
    }
    for (int s = 0; s < 6; ++s) dd_matmul(P.data(), P.data(), P.data(), K);  // A^64
    for (int k = 0; k < IIR_LOG_T; ++k) {
        for (int i = 0; i < K * K; ++i) out->Ppow[k][i] = P[i];
        dd_matmul(P.data(), P.data(), P.data(), K);
    }
    for (int i = 0; i < K * K; ++i) out->PT[i] = P[i];  // A^(64*128)
    return 0;
}

// Does the block scan stay inside the parity budget for this filter? Measured directly at create time: 4096 samples
// of white noise through (a) lfilter's sequential float64 recursion — the reference — and (b) what the scan computes:
// every 64-sample segment restarted from the exact state (long double here, double-double on the device) rounded to
// float64. For a well-conditioned filter the two agree to ~1e-9 (MPX low-pass 15 kHz @ 2.4 MS/s: 1.0e-9); tf-form
// filters with large transient growth do not: order-10 band-pass 300-3000 Hz @ 48 kS/s 3.8e-6 here and 5e-4 on the
// device over 500 000 samples, 100 Hz high-pass @ 48 kS/s 2.1e-7, 3 kHz low-pass @ 10 MS/s 4.5e-4 (the regime SURVEY
// App. A.3(iii) describes). Above 1e-7 the handle replays the recursion sequentially (bit-equal to scipy).
static bool iir_needs_sequential(const IirCoef& c) {
    const int K = c.K;
    if (K < 2) return false;
    typedef long double ld;
    const int N = 4096;
    uint64_t lcg = 0x9E3779B97F4A7C15ull;
    std::vector<double> x(N);
    for (int n = 0; n < N; ++n) {
        lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
        x[n] = (double)(lcg >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    }
    std::vector<double> zs(K, 0.0), zb(K, 0.0);
    std::vector<ld> ze(K, 0.0L);
    double num = 0.0, den = 0.0;
    for (int n = 0; n < N; ++n) {
        if ((n & (IIR_L - 1)) == 0)
            for (int i = 0; i < K; ++i) zb[i] = (double)ze[i];
        const double xv = x[n];
        // (a) lfilter: products and sums rounded separately
        const double ya = zs[0] + c.b0 * xv;
        for (int i = 0; i < K - 1; ++i) zs[i] = (zs[i + 1] + xv * c.b[i]) - ya * c.a[i];
        zs[K - 1] = xv * c.b[K - 1] - ya * c.a[K - 1];
        // (b) the scan's segment body (df2t_step)
        const double yb = fma(c.b0, xv, zb[0]);
        for (int i = 0; i < K - 1; ++i) zb[i] = fma(c.b[i], xv, fma(-c.a[i], yb, zb[i + 1]));
        zb[K - 1] = fma(c.b[K - 1], xv, -c.a[K - 1] * yb);
        // exact state
        const ld xe = (ld)xv, ye = ze[0] + (ld)c.b0 * xe;
        for (int i = 0; i < K - 1; ++i) ze[i] = ze[i + 1] + xe * (ld)c.b[i] - ye * (ld)c.a[i];
        ze[K - 1] = xe * (ld)c.b[K - 1] - ye * (ld)c.a[K - 1];
        if (n >= N / 2) {
            num += (yb - ya) * (yb - ya);
            den += ya * ya;
        }
    }
    if (!(den > 0.0) || !(num == num)) return true;   // unstable / non-finite: replay, so the garbage matches too
    return sqrt(num / den) > 1e-7;
}

// May the scan chain its segment states in plain float64 (transition powers and states rounded to double) instead of
// double-double? Measured like iir_needs_sequential: 4096 samples of noise, state chained segment by segment as
// s' = fl(P s + z) with P = A^64 rounded to float64 and z the segment's zero-state response, outputs re-run from those
// states, against lfilter's sequential recursion. Well-conditioned filters (one-pole de-emphasis / AGC envelopes, the MPX
// low-pass) agree to ~1e-12 and take the plain scan: ~20 x fewer FP64 operations per scan term (ncu: the double-double scan
// was ~3/4 of the kernel's FP64 work for K = 5).
static bool iir_scan_plain_ok(const IirCoef& c) {
    const int K = c.K;
    if (K < 1) return true;
    const int N = 4096;
    uint64_t lcg = 0x9E3779B97F4A7C15ull;
    std::vector<double> zs(K, 0.0), z0(K, 0.0), zr(K, 0.0), sp(K, 0.0), nx(K);
    double num = 0.0, den = 0.0;
    for (int n = 0; n < N; ++n) {
        if ((n & (IIR_L - 1)) == 0) {
            if (n > 0) {   // chain: state at the end of the previous segment = P * (its start state) + (its zero-state response)
                for (int r = 0; r < K; ++r) {
                    double acc = z0[r];
                    for (int q = 0; q < K; ++q) acc = fma(c.Ppow[0][r * K + q].hi, sp[q], acc);
                    nx[r] = acc;
                }
                sp = nx;
            }
            zr = sp;
            std::fill(z0.begin(), z0.end(), 0.0);
        }
        lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
        const double xv = (double)(lcg >> 11) * (2.0 / 9007199254740992.0) - 1.0;
        const double ya = zs[0] + c.b0 * xv;
        for (int i = 0; i < K - 1; ++i) zs[i] = (zs[i + 1] + xv * c.b[i]) - ya * c.a[i];
        zs[K - 1] = xv * c.b[K - 1] - ya * c.a[K - 1];
        const double y0 = fma(c.b0, xv, z0[0]);
        for (int i = 0; i < K - 1; ++i) z0[i] = fma(c.b[i], xv, fma(-c.a[i], y0, z0[i + 1]));
        z0[K - 1] = fma(c.b[K - 1], xv, -c.a[K - 1] * y0);
        const double yc = fma(c.b0, xv, zr[0]);
        for (int i = 0; i < K - 1; ++i) zr[i] = fma(c.b[i], xv, fma(-c.a[i], yc, zr[i + 1]));
        zr[K - 1] = fma(c.b[K - 1], xv, -c.a[K - 1] * yc);
        if (n >= N / 2) {
            num += (yc - ya) * (yc - ya);
            den += ya * ya;
        }
    }
    if (!(den > 0.0) || !(num == num)) return false;
    return sqrt(num / den) < 1e-8;
}

struct Workspace {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t need) {
        if (cap >= need) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        WC_CUDA(cudaMalloc(&p, need));
        cap = need;
        return 0;
    }
    ~Workspace() {
        if (p) cudaFree(p);
    }
};

template <int K, bool PLAIN, typename TIn>
static int launch_iir_k(const IirCoef* d_cf, const TIn* x, float* y, int n, long long seq_stride, int n_seq,
                        int absin, double* zseg, dd* ztile, dd* stile, int tiles, double* sumsq, cudaStream_t st) {
    const size_t smem = iir_smem_bytes<K>();
    static std::atomic<unsigned long long> done0{0}, done1{0};
    WC_CUDA(smem_optin(iir_kernel<K, 0, PLAIN, TIn>, (int)smem, done0));
    WC_CUDA(smem_optin(iir_kernel<K, 1, PLAIN, TIn>, (int)smem, done1));
    dim3 grid(tiles, n_seq);
    iir_kernel<K, 0, PLAIN, TIn><<<grid, IIR_T, smem, st>>>(d_cf, x, y, n, seq_stride, zseg, ztile, stile, tiles, absin, nullptr);
    iir_chain_kernel<K, PLAIN><<<(n_seq + 3) / 4, 128, 0, st>>>(d_cf, ztile, stile, tiles, n_seq);
    iir_kernel<K, 1, PLAIN, TIn><<<grid, IIR_T, smem, st>>>(d_cf, x, y, n, seq_stride, zseg, ztile, stile, tiles, absin, sumsq);
    WC_CUDA(cudaGetLastError());
    return 0;
}

template <typename TIn>
static int launch_iir_typed(const IirCoef* d_cf, int K, bool plain, const TIn* x, float* y, int n, long long seq_stride,
                            int n_seq, int absin, double* zseg, dd* ztile, dd* stile, int tiles, double* sumsq, cudaStream_t st) {
#define WC_IIR_CASE(KK)                                                                                                   \
    case KK:                                                                                                              \
        return plain ? launch_iir_k<KK, true, TIn>(d_cf, x, y, n, seq_stride, n_seq, absin, zseg, ztile, stile, tiles, sumsq, st) \
                     : launch_iir_k<KK, false, TIn>(d_cf, x, y, n, seq_stride, n_seq, absin, zseg, ztile, stile, tiles, sumsq, st);
    switch (K) {
        WC_IIR_CASE(1) WC_IIR_CASE(2) WC_IIR_CASE(3) WC_IIR_CASE(4) WC_IIR_CASE(5) WC_IIR_CASE(6) WC_IIR_CASE(7)
        WC_IIR_CASE(8) WC_IIR_CASE(9) WC_IIR_CASE(10)
        default:
            set_error("iir: unsupported order %d", K);
            return -1;
    }
#undef WC_IIR_CASE
}


// ---------------------------------------------------------------------------------------------
// live signal metrics (Channel.update_signal_metrics, capture.py:749-798): |freq_shift(iq)| per channel with the sum
// of squares (RSSI), and the two order statistics np.partition picks for the SNR estimate — exact, by 4-pass radix
// select on the IEEE bit patterns of the (non-negative) magnitudes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) metrics_mag_kernel(const void* __restrict__ iq, int fmt, int n, const FrontChan* __restrict__ chs,
                                                          float* __restrict__ mag, double* __restrict__ power) {
    __shared__ double red[8];
    const int c = blockIdx.y;
    const FrontChan ch = chs[c];
    double psum = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float2 v;
        if (fmt == 0) v = reinterpret_cast<const float2*>(iq)[i];
        else {
            const short2 q = reinterpret_cast<const short2*>(iq)[i];
            v = make_float2((float)q.x / 32768.0f, (float)q.y / 32768.0f);
        }
        if (ch.shift) {
            float cs, sn;
            nco_f32(ch.k32, i, cs, sn);
            v = make_float2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
        }
        const float m = sqrtf(v.x * v.x + v.y * v.y);
        mag[(long long)c * n + i] = m;
        psum += (double)(m * m);
    }
    psum = warp_sum(psum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = psum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(power + c, t);
    }
}

// out[c][q] = the rank[q]-th smallest of mag[c][0..n) (0-based), q = 0, 1
__global__ void __launch_bounds__(1024) metrics_select_kernel(const float* __restrict__ mag, int n, int rank0, int rank1,
                                                              float* __restrict__ out) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_rank;
    const float* xs = mag + (long long)blockIdx.x * n;
    const int ranks[2] = {rank0, rank1};
    for (int q = 0; q < 2; ++q) {
        unsigned prefix = 0u, mask = 0u;
        unsigned rank = (unsigned)ranks[q];
        for (int pass = 3; pass >= 0; --pass) {
            for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const unsigned u = __float_as_uint(xs[i]);
                if ((u & mask) == prefix) atomicAdd(&hist[(u >> (8 * pass)) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned cum = 0u;
                int b = 0;
                for (; b < 255; ++b) {
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
        if (threadIdx.x == 0) out[blockIdx.x * 2 + q] = __uint_as_float(prefix);
    }
}

}  // namespace wc

// Pinned staging ring for the small per-call parameter arrays of the stateless entry points: the async copy reads the
// slot when the stream gets to it, so a slot is only reused after its event has completed (no stream synchronise on
// the call path; the reference calls these entry points from several threads, hence the mutex).
#include <mutex>
namespace {
constexpr int RING_SLOTS = 32;
constexpr size_t RING_BYTES = 64 * 1024;
struct ParamRing {
    std::mutex mu;
    void* host[RING_SLOTS] = {};
    cudaEvent_t ev[RING_SLOTS] = {};
    bool used[RING_SLOTS] = {};
    int next = 0;
    bool ok = false, tried = false;
};
ParamRing g_ring;

// returns a pinned slot (or nullptr: caller falls back to pageable memory + synchronise)
void* ring_acquire(size_t bytes, int* slot) {
    if (bytes > RING_BYTES) return nullptr;
    std::lock_guard<std::mutex> lk(g_ring.mu);
    if (!g_ring.tried) {
        g_ring.tried = true;
        g_ring.ok = true;
        for (int i = 0; i < RING_SLOTS && g_ring.ok; ++i) {
            g_ring.ok = cudaHostAlloc(&g_ring.host[i], RING_BYTES, cudaHostAllocDefault) == cudaSuccess &&
                        cudaEventCreateWithFlags(&g_ring.ev[i], cudaEventDisableTiming) == cudaSuccess;
        }
        if (!g_ring.ok) cudaGetLastError();
    }
    if (!g_ring.ok) return nullptr;
    const int s = g_ring.next;
    g_ring.next = (g_ring.next + 1) % RING_SLOTS;
    if (g_ring.used[s]) cudaEventSynchronize(g_ring.ev[s]);
    g_ring.used[s] = true;
    *slot = s;
    return g_ring.host[s];
}
void ring_release(int slot, cudaStream_t st) { cudaEventRecord(g_ring.ev[slot], st); }
}  // namespace

using namespace wc;

// ---------------------------------------------------------------------------------------------
// C ABI: stage-level operators (device pointers)
// ---------------------------------------------------------------------------------------------
struct wc_iir {
    IirCoef h_cf;
    IirCoef* d_cf = nullptr;
    bool sequential = false;   // exact-replay kernel instead of the block scan (iir_needs_sequential)
    bool plain = false;        // block scan with float64 chaining instead of double-double (iir_scan_plain_ok)
};

template <int K>
static int launch_iir_seq_k(const IirCoef* d_cf, const float* x, float* y, int n, long long seq_stride, int n_seq,
                            int absin, cudaStream_t st) {
    iir_seq_kernel<K><<<(n_seq + 31) / 32, 32, 0, st>>>(d_cf, x, y, n, seq_stride, n_seq, absin);
    WC_CUDA(cudaGetLastError());
    return 0;
}
static int launch_iir_seq(const IirCoef* d_cf, int K, const float* x, float* y, int n, long long seq_stride, int n_seq,
                          int absin, cudaStream_t st) {
#define WC_IIRS_CASE(KK) \
    case KK:             \
        return launch_iir_seq_k<KK>(d_cf, x, y, n, seq_stride, n_seq, absin, st);
    switch (K) {
        WC_IIRS_CASE(1) WC_IIRS_CASE(2) WC_IIRS_CASE(3) WC_IIRS_CASE(4) WC_IIRS_CASE(5) WC_IIRS_CASE(6) WC_IIRS_CASE(7)
        WC_IIRS_CASE(8) WC_IIRS_CASE(9) WC_IIRS_CASE(10)
        default:
            set_error("iir: unsupported order %d", K);
            return -1;
    }
#undef WC_IIRS_CASE
}

// scratch bytes of the block scan for n_seq sequences of n samples (zseg | ztile | stile)
static size_t iir_scan_scratch_bytes(int K, int n, int n_seq) {
    const size_t nz = (size_t)n_seq * ((n + IIR_TILE - 1) / IIR_TILE);
    return sizeof(double) * nz * IIR_T * K + sizeof(dd) * 2 * nz * K;
}

// the filter proper on caller-provided scratch (iir_scan_scratch_bytes; unused by the sequential replay)
// sumsq_dev (optional, zeroed by the caller): per-sequence sum of squares of the float32 output, for the scan flavours;
// returns 1 in *sumsq_done when it was produced (the sequential replay does not fuse it)
static int iir_run(const wc_iir* h, const float* x_dev, float* y_dev, int n, long long seq_stride, int n_seq,
                   int abs_input, void* scratch, cudaStream_t st, double* sumsq_dev = nullptr, bool* sumsq_done = nullptr) {
    const int K = h->h_cf.K;
    if (sumsq_done) *sumsq_done = false;
    if (h->sequential) return launch_iir_seq(h->d_cf, K, x_dev, y_dev, n, seq_stride, n_seq, abs_input, st);
    if (sumsq_done) *sumsq_done = sumsq_dev != nullptr;
    const int tiles = (n + IIR_TILE - 1) / IIR_TILE;
    const size_t nz = (size_t)n_seq * tiles;
    double* zseg = reinterpret_cast<double*>(scratch);
    dd* ztile = reinterpret_cast<dd*>(zseg + nz * IIR_T * K);
    dd* stile = ztile + nz * K;
    return launch_iir_typed<float>(h->d_cf, K, h->plain, x_dev, y_dev, n, seq_stride, n_seq, abs_input, zseg, ztile, stile,
                                   tiles, sumsq_dev, st);
}

extern "C" {

int wc_iir_create(const double* b, int nb, const double* a, int na, wc_iir** out) {
    WC_REQUIRE(b && a && out, "wc_iir_create: null argument");
    wc_iir* h = new wc_iir();
    if (make_iir_coef(b, nb, a, na, &h->h_cf)) {
        delete h;
        set_error("wc_iir_create: unsupported filter (order must be 0..%d, a[0] != 0)", IIR_KMAX);
        return -1;
    }
    if (cudaMalloc(&h->d_cf, sizeof(IirCoef)) != cudaSuccess) {
        delete h;
        set_error("wc_iir_create: cudaMalloc failed");
        return -2;
    }
    cudaMemcpy(h->d_cf, &h->h_cf, sizeof(IirCoef), cudaMemcpyHostToDevice);
    h->sequential = iir_needs_sequential(h->h_cf);
    h->plain = !h->sequential && iir_scan_plain_ok(h->h_cf);
    *out = h;
    return 0;
}

// 1 when the handle runs the sequential exact replay (ill-conditioned tf-form filter), 0 for the block scan
int wc_iir_is_sequential(const wc_iir* h) { return (h && h->sequential) ? 1 : 0; }
// 0: block scan chained in double-double, 1: block scan chained in float64, 2: sequential exact replay
int wc_iir_kind(const wc_iir* h) { return !h ? -1 : h->sequential ? 2 : h->plain ? 1 : 0; }

void wc_iir_destroy(wc_iir* h) {
    if (!h) return;
    if (h->d_cf) cudaFree(h->d_cf);
    delete h;
}

// y = lfilter(b, a, x) per sequence, zero initial state; x float32 [n_seq][seq_stride], y may alias x.
// abs_input != 0 filters |x| (AGC envelope, dsp/agc.py:83).
int wc_iir_lfilter(wc_iir* h, const float* x_dev, float* y_dev, int n, long long seq_stride, int n_seq,
                   int abs_input, void* stream) {
    WC_REQUIRE(h && x_dev && y_dev, "wc_iir_lfilter: null argument");
    if (n <= 0 || n_seq <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int K = h->h_cf.K;
    if (K == 0) {
        const long long total = (long long)n;
        for (int s = 0; s < n_seq; ++s)
            elementwise_kernel<<<(unsigned)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256), 256, 0, st>>>(
                x_dev + s * seq_stride, y_dev + s * seq_stride, total, 2, (float)h->h_cf.b0);
        WC_CUDA(cudaGetLastError());
        return 0;
    }
    if (h->sequential) return iir_run(h, x_dev, y_dev, n, seq_stride, n_seq, abs_input, nullptr, st);
    // Handles are shared between threads (dsp/_stages.py caches them by coefficients and the reference calls
    // _process_channel_dsp_stateless from a 3-worker pool, capture.py:1906-1925), so a call owns its scratch:
    // stream-ordered allocation from the device's default pool (wc_init keeps the pool from trimming).
    void* scratch = nullptr;
    WC_CUDA(cudaMallocAsync(&scratch, iir_scan_scratch_bytes(K, n, n_seq), st));
    const int rc = iir_run(h, x_dev, y_dev, n, seq_stride, n_seq, abs_input, scratch, st);
    cudaFreeAsync(scratch, st);
    return rc;
}

// sum of squares per sequence (float64 accumulators); out_dev[n_seq] is overwritten.
int wc_sumsq(const float* x_dev, int n, long long seq_stride, int n_seq, double* out_dev, void* stream) {
    WC_REQUIRE(x_dev && out_dev, "wc_sumsq: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    WC_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double) * n_seq, st));
    if (n <= 0) return 0;
    int bx = (n + 255) / 256;
    if (bx > 64) bx = 64;
    sumsq_kernel<<<dim3(bx, n_seq), 256, 0, st>>>(x_dev, n, seq_stride, out_dev);
    WC_CUDA(cudaGetLastError());
    return 0;
}

int wc_elementwise(const float* x_dev, float* y_dev, long long total, int op, float p0, void* stream) {
    WC_REQUIRE(x_dev && y_dev, "wc_elementwise: null argument");
    if (total <= 0) return 0;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    elementwise_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x_dev, y_dev, total, op, p0);
    WC_CUDA(cudaGetLastError());
    return 0;
}

int wc_agc_apply(const float* x_dev, const float* env_attack_dev, const float* env_release_dev, float* y_dev,
                 long long total, float target_linear, float max_gain_linear, void* stream) {
    WC_REQUIRE(x_dev && env_attack_dev && env_release_dev && y_dev, "wc_agc_apply: null argument");
    if (total <= 0) return 0;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    agc_apply_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x_dev, env_attack_dev, env_release_dev, y_dev,
                                                                        total, target_linear, max_gain_linear);
    WC_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"

// ---- resampler handle -------------------------------------------------------------------------
struct wc_resampler {
    int up, down, ntaps, tpp, half_len;
    long long n_pre_pad, n_pre_remove;
    double* d_hp = nullptr;
    double* d_g = nullptr;     // residue-form taps [up][E][down] when the ratio qualifies (E > 0)
    float* d_gf = nullptr;     // the same, float32
    int* d_elo = nullptr;      // [down]
    int E = 0, E_need = 0;
};

// template instances of resample_residue_kernel: (up, padded E)
static int residue_slot(int up, int e_need) {
    if (up == 1 && e_need <= 21) return 21;
    if (up == 2 && e_need <= 12) return 12;
    if (up == 3 && e_need <= 8) return 8;
    return 0;
}

extern "C" {

// taps = the FIR scipy.signal.resample_poly would use: firwin(2*half_len+1, 1/max(up,down),
// window=("kaiser", 5.0)) * up with half_len = 10*max(up,down). Passing the taps keeps filter design
// on the host (as in the reference); the index math (pre-pad / pre-remove / zero extension) is here.
int wc_resampler_create(int up, int down, const double* taps, int ntaps, wc_resampler** out) {
    WC_REQUIRE(out && taps && up >= 1 && down >= 1 && ntaps >= 1 && (ntaps & 1), "wc_resampler_create: bad arguments");
    wc_resampler* h = new wc_resampler();
    h->up = up;
    h->down = down;
    h->ntaps = ntaps;
    h->half_len = (ntaps - 1) / 2;
    h->n_pre_pad = down - h->half_len % down;                   // scipy: n_pre_pad = down - half_len % down
    h->n_pre_remove = (h->half_len + h->n_pre_pad) / down;      // scipy: (half_len + n_pre_pad) // down
    h->tpp = (ntaps + up - 1) / up;
    std::vector<double> hp((size_t)up * h->tpp, 0.0);
    for (int t = 0; t < ntaps; ++t) hp[(size_t)(t % up) * h->tpp + t / up] = taps[t];
    if (cudaMalloc(&h->d_hp, hp.size() * sizeof(double)) != cudaSuccess) {
        delete h;
        set_error("wc_resampler_create: cudaMalloc failed");
        return -2;
    }
    cudaMemcpy(h->d_hp, hp.data(), hp.size() * sizeof(double), cudaMemcpyHostToDevice);
    // residue form (decimating ratios): tap of (m = up*a + mu, n = r0 + down*k) is h[B + up*down*e], e = a - k
    if (down >= 32 && up <= 3) {
        const long long q0 = h->n_pre_remove * down - h->n_pre_pad, UD = (long long)up * down;
        auto fdiv = [](long long x, long long y) { return (x >= 0) ? x / y : -((-x + y - 1) / y); };   // floor
        std::vector<int> elo(down, 0);
        int e_need = 0;
        for (int r0 = 0; r0 < down; ++r0) {
            long long lo_min = 1LL << 40, hi_max = -(1LL << 40);
            for (int mu = 0; mu < up; ++mu) {
                const long long B = q0 + (long long)down * mu - (long long)up * r0;
                const long long lo = -fdiv(B, UD);                 // ceil(-B / UD)
                const long long hi = fdiv((long long)ntaps - 1 - B, UD);
                if (lo <= hi) {
                    if (lo < lo_min) lo_min = lo;
                    if (hi > hi_max) hi_max = hi;
                }
            }
            if (hi_max >= lo_min) {
                elo[r0] = (int)lo_min;
                if ((int)(hi_max - lo_min + 1) > e_need) e_need = (int)(hi_max - lo_min + 1);
            }
        }
        const int E = e_need ? residue_slot(up, e_need) : 0;
        if (E) {
            std::vector<double> g((size_t)up * E * down, 0.0);
            for (int mu = 0; mu < up; ++mu)
                for (int r0 = 0; r0 < down; ++r0) {
                    const long long B = q0 + (long long)down * mu - (long long)up * r0;
                    for (int j = 0; j < E; ++j) {
                        const long long t = B + UD * (elo[r0] + j);
                        if (t >= 0 && t < ntaps) g[((size_t)mu * E + j) * down + r0] = taps[t];
                    }
                }
            std::vector<float> gf(g.begin(), g.end());
            if (cudaMalloc(&h->d_g, g.size() * sizeof(double)) == cudaSuccess &&
                cudaMalloc(&h->d_gf, gf.size() * sizeof(float)) == cudaSuccess &&
                cudaMalloc(&h->d_elo, elo.size() * sizeof(int)) == cudaSuccess) {
                cudaMemcpy(h->d_g, g.data(), g.size() * sizeof(double), cudaMemcpyHostToDevice);
                cudaMemcpy(h->d_gf, gf.data(), gf.size() * sizeof(float), cudaMemcpyHostToDevice);
                cudaMemcpy(h->d_elo, elo.data(), elo.size() * sizeof(int), cudaMemcpyHostToDevice);
                h->E = E;
                h->E_need = e_need;
            }
        }
    }
    *out = h;
    return 0;
}

void wc_resampler_destroy(wc_resampler* h) {
    if (!h) return;
    if (h->d_hp) cudaFree(h->d_hp);
    if (h->d_g) cudaFree(h->d_g);
    if (h->d_gf) cudaFree(h->d_gf);
    if (h->d_elo) cudaFree(h->d_elo);
    delete h;
}

long long wc_resampler_out_len(const wc_resampler* h, long long n_in) {
    if (!h || n_in <= 0) return 0;
    return (n_in * h->up + h->down - 1) / h->down;  // ceil(n_in*up/down)
}

// epilogue: WC_EPI_NONE / WC_EPI_RMS_CLIP / WC_EPI_CLIP / WC_EPI_RMS (see header)
int wc_resampler_run(wc_resampler* h, const float* x_dev, int n_in, long long seq_stride, int n_seq, float* out_dev,
                     int epilogue, const double* sumsq_dev, float target_rms, float min_rms, double* power_dev,
                     int* invalid_dev, float max_abs, void* stream) {
    WC_REQUIRE(h && x_dev && out_dev, "wc_resampler_run: null argument");
    WC_REQUIRE(!((epilogue == 1 || epilogue == 3) && !sumsq_dev), "wc_resampler_run: epilogue needs sumsq");
    const long long n_out = wc_resampler_out_len(h, n_in);
    if (n_out <= 0 || n_seq <= 0) return 0;
    ResampArgs a;
    a.x = x_dev;
    a.seq_stride = seq_stride;
    a.n_in = n_in;
    a.n_out = (int)n_out;
    a.up = h->up;
    a.down = h->down;
    a.q0 = h->n_pre_remove * h->down - h->n_pre_pad;
    a.ntaps = h->ntaps;
    a.tpp = h->tpp;
    a.hp = h->d_hp;
    a.g = h->d_g;
    a.gf = h->d_gf;
    a.e_lo = h->d_elo;
    a.E = h->E;
    a.out = out_dev;
    a.epi = epilogue;
    a.sumsq = sumsq_dev;
    a.target_rms = target_rms;
    a.min_rms = min_rms;
    a.power = power_dev;
    a.invalid = invalid_dev;
    a.max_abs = max_abs;
    const long long per_branch = (n_out + h->up - 1) / h->up;
    const long long n_tasks = ((per_branch + RS_R - 1) / RS_R) * h->up;
    int bx = (int)((n_tasks + RS_WARPS - 1) / RS_WARPS);
    if (bx < 1) bx = 1;
    if (bx > 148 * 8) bx = 148 * 8;
    const bool no_residue = env_int("WC_RESAMPLE_RESIDUE", 1) == 0;   // A/B timing
    if (h->E && !no_residue) {
        const long long per_phase = (n_out + h->up - 1) / h->up;
        const int n_blocks = (int)((per_phase + RS_A - 1) / RS_A);
        const int G = (h->down <= 32) ? 4 : (h->down <= 64) ? 2 : 1;
        const int NB = max(1, env_int("WC_RS_NB", RS_NB));
        const dim3 grid((unsigned)((n_blocks + G * NB - 1) / (G * NB)), (unsigned)n_seq);
        cudaStream_t st = (cudaStream_t)stream;
        const bool f64 = env_int("WC_RESAMPLE_F64", 0) == 1;   // A/B: all-float64 kernel
        if (f64) {
            if (h->up == 1) resample_residue_kernel<1, 21, double><<<grid, RS_RT, 0, st>>>(a, n_blocks, G, NB);
            else if (h->up == 2) resample_residue_kernel<2, 12, double><<<grid, RS_RT, 0, st>>>(a, n_blocks, G, NB);
            else resample_residue_kernel<3, 8, double><<<grid, RS_RT, 0, st>>>(a, n_blocks, G, NB);
        } else {
            if (h->up == 1) resample_residue_kernel<1, 21, float><<<grid, RS_RT, 0, st>>>(a, n_blocks, G, NB);
            else if (h->up == 2) resample_residue_kernel<2, 12, float><<<grid, RS_RT, 0, st>>>(a, n_blocks, G, NB);
            else resample_residue_kernel<3, 8, float><<<grid, RS_RT, 0, st>>>(a, n_blocks, G, NB);
        }
        WC_CUDA(cudaGetLastError());
        return 0;
    }
    // grouped kernel only while the RS_R outputs of a task read one L1-friendly window (measured on B200: 1/50 resampler
    // 11 % faster grouped, 3/625 resampler 14 % slower)
    if ((long long)(RS_R - 1) * h->down * (long long)sizeof(float) <= 8192) {
        resample_kernel<<<dim3(bx, n_seq), RS_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    } else {
        int b1 = (int)((n_out + RS_WARPS - 1) / RS_WARPS);
        if (b1 > 148 * 8) b1 = 148 * 8;
        resample_single_kernel<<<dim3(b1, n_seq), RS_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    }
    WC_CUDA(cudaGetLastError());
    return 0;
}

// ---- front end ----------------------------------------------------------------------------------
// One launch for all channels of all chunks; see FrontArgs. chan arrays are host pointers [n_ch].
int wc_front_run(const void* iq_dev, int fmt, int n, int n_chunks, long long chunk_stride, int n_ch,
                 const int* modes, const double* offsets_hz, const double* bfo_hz, int sample_rate,
                 float* out_dev, void* base_out_dev, double* power_dev, int* nonfinite_dev, void* chan_scratch_dev,
                 void* stream) {
    return wc_front_run_ex(iq_dev, fmt, n, n_chunks, chunk_stride, n_ch, modes, offsets_hz, bfo_hz, sample_rate, out_dev,
                           base_out_dev, power_dev, nullptr, nonfinite_dev, chan_scratch_dev, stream);
}

int wc_front_run_ex(const void* iq_dev, int fmt, int n, int n_chunks, long long chunk_stride, int n_ch,
                    const int* modes, const double* offsets_hz, const double* bfo_hz, int sample_rate,
                    float* out_dev, void* base_out_dev, double* power_dev, double* out_sumsq_dev, int* nonfinite_dev,
                    void* chan_scratch_dev, void* stream) {
    WC_REQUIRE(iq_dev && modes && offsets_hz && power_dev && nonfinite_dev && chan_scratch_dev,
               "wc_front_run: null argument");
    WC_REQUIRE(n_ch >= 1 && n_ch <= 4096, "wc_front_run: n_ch out of range");
    cudaStream_t st = (cudaStream_t)stream;
    int slot = -1;
    FrontChan* pinned = reinterpret_cast<FrontChan*>(ring_acquire(sizeof(FrontChan) * n_ch, &slot));
    std::vector<FrontChan> pageable;
    if (!pinned) pageable.resize(n_ch);
    FrontChan* ch = pinned ? pinned : pageable.data();
    for (int c = 0; c < n_ch; ++c) {
        FrontChan& f = ch[c];
        f.mode = modes[c];
        f.shift = (offsets_hz[c] != 0.0) ? 1 : 0;  // capture.py:185 / :328
        // numpy: complex64(-1j*2*pi*(round(off)/fs)) * float32(n)  ->  k32 = float32(2*pi*round(off)/fs)
        const double off = nearbyint(offsets_hz[c]);  // Python round() = banker's rounding
        f.k32 = (float)(-(2.0 * M_PI * (off / (double)sample_rate)));
        f.bfo_turns = bfo_hz ? bfo_hz[c] / (double)sample_rate : 0.0;
        f.disc_scale = (float)((double)sample_rate / (2.0 * M_PI * 75000.0));
    }
    WC_CUDA(cudaMemcpyAsync(chan_scratch_dev, ch, sizeof(FrontChan) * n_ch, cudaMemcpyHostToDevice, st));
    if (pinned) ring_release(slot, st);
    WC_CUDA(cudaMemsetAsync(power_dev, 0, sizeof(double) * (size_t)n_chunks * n_ch, st));
    WC_CUDA(cudaMemsetAsync(nonfinite_dev, 0, sizeof(int) * (size_t)n_chunks, st));
    if (out_sumsq_dev) WC_CUDA(cudaMemsetAsync(out_sumsq_dev, 0, sizeof(double) * (size_t)n_chunks * n_ch, st));
    if (n <= 0) {
        if (!pinned) WC_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
    FrontArgs a;
    a.iq = iq_dev;
    a.chunk_stride = chunk_stride;
    a.n = n;
    a.fmt = fmt;
    a.n_ch = n_ch;
    a.n_chunks = n_chunks;
    a.ch = reinterpret_cast<const FrontChan*>(chan_scratch_dev);
    a.out = out_dev;
    a.base_out = reinterpret_cast<float2*>(base_out_dev);
    a.power = power_dev;
    a.out_sumsq = out_sumsq_dev;
    a.nonfinite = nonfinite_dev;
    // ncu, C2 (123 tiles x 8 chunks = 984 CTAs at 5 per SM): 1.33 waves, i.e. a third of the run on a quarter-full machine.
    // Channel groups bring the grid to >= 4 waves; the tile is then staged once per group (L2 hits, 1/16 of a channel's work).
    const int tiles = (n + FR_TILE - 1) / FR_TILE;
    const long long base_ctas = (long long)tiles * n_chunks;
    bool all_fm_shared = base_out_dev == nullptr;
    for (int c = 0; c < n_ch; ++c)
        all_fm_shared = all_fm_shared && (modes[c] == WC_MODE_WBFM || modes[c] == WC_MODE_NBFM) && offsets_hz[c] != 0.0;
    const int groups = front_groups(base_ctas, n_ch, all_fm_shared);
    WC_CUDA(smem_optin(front_kernel, (int)sizeof(FrontSmem), g_front_optin));
    front_kernel<<<dim3(tiles, n_chunks, (unsigned)groups), FR_THREADS, sizeof(FrontSmem), st>>>(a);
    WC_CUDA(cudaGetLastError());
    if (!pinned) WC_CUDA(cudaStreamSynchronize(st));   // a pageable host vector must outlive the async copy
    return 0;
}

int wc_front_chan_scratch_bytes(int n_ch) { return (int)(sizeof(FrontChan) * (size_t)n_ch); }

int wc_finalize(float* audio_dev, int n_out, int n_chunks, int n_seq, const double* power_iq_dev, int n_in,
                const float* squelch_db_dev, const int* has_squelch_dev, float* rssi_db_dev,
                unsigned char* squelched_dev, void* stream) {
    WC_REQUIRE(audio_dev && power_iq_dev && squelch_db_dev && has_squelch_dev && rssi_db_dev && squelched_dev,
               "wc_finalize: null argument");
    if (n_seq <= 0) return 0;
    int bx = n_out > 0 ? (n_out + 255) / 256 : 1;
    if (bx > 16) bx = 16;
    finalize_kernel<<<dim3(bx, n_seq), 256, 0, (cudaStream_t)stream>>>(audio_dev, n_out, n_chunks, power_iq_dev, n_in,
                                                                       squelch_db_dev, has_squelch_dev, rssi_db_dev,
                                                                       squelched_dev);
    WC_CUDA(cudaGetLastError());
    return 0;
}

/* Channel.update_signal_metrics (capture.py:749-798) for n_ch channels of one chunk: power_dev[c] = sum |freq_shift(iq)|^2,
 * pct_dev[c] = {magnitudes partitioned at n//10, at n - n//10 - 1} (only when want_snr and those ranks are usable,
 * :781-785). mag_scratch_dev: float32 [n_ch][n]. */
int wc_signal_metrics(const void* iq_dev, int fmt, int n, int sample_rate, const double* offsets_hz, int n_ch, int want_snr,
                      float* mag_scratch_dev, double* power_dev, float* pct_dev, void* chan_scratch_dev, void* stream) {
    WC_REQUIRE(iq_dev && offsets_hz && mag_scratch_dev && power_dev && pct_dev && chan_scratch_dev,
               "wc_signal_metrics: null argument");
    WC_REQUIRE(n_ch >= 1 && n_ch <= 4096 && n >= 0, "wc_signal_metrics: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<FrontChan> ch(n_ch);
    for (int c = 0; c < n_ch; ++c) {
        FrontChan& f = ch[c];
        f.mode = WC_MODE_NONE;
        f.shift = (offsets_hz[c] != 0.0) ? 1 : 0;
        const double off = nearbyint(offsets_hz[c]);
        f.k32 = (float)(-(2.0 * M_PI * (off / (double)sample_rate)));
        f.bfo_turns = 0.0;
        f.disc_scale = 0.f;
    }
    WC_CUDA(cudaMemcpyAsync(chan_scratch_dev, ch.data(), sizeof(FrontChan) * n_ch, cudaMemcpyHostToDevice, st));
    WC_CUDA(cudaMemsetAsync(power_dev, 0, sizeof(double) * (size_t)n_ch, st));
    WC_CUDA(cudaMemsetAsync(pct_dev, 0, sizeof(float) * 2 * (size_t)n_ch, st));
    if (n > 0) {
        int bx = (n + 255) / 256;
        if (bx > 592) bx = 592;
        metrics_mag_kernel<<<dim3(bx, n_ch), 256, 0, st>>>(iq_dev, fmt, n, reinterpret_cast<const FrontChan*>(chan_scratch_dev),
                                                           mag_scratch_dev, power_dev);
        const int k_noise = n / 10, k_signal = n - n / 10 - 1;
        if (want_snr && k_noise > 0 && k_signal > k_noise)
            metrics_select_kernel<<<n_ch, 1024, 0, st>>>(mag_scratch_dev, n, k_noise, k_signal, pct_dev);
        WC_CUDA(cudaGetLastError());
    }
    WC_CUDA(cudaStreamSynchronize(st));  // the pageable host vector must outlive the async copy
    return 0;
}

}  // extern "C"

// =================================================================================================
// Synchronous AM: carrier-recovery PLL (dsp/sam.py:26-129 CarrierRecoveryPLL.process, :132-270 sam_demod)
// =================================================================================================
// The reference walks the samples in a Python loop: lo = exp(-1j * phase) (complex128), mixed = iq[i] * lo, coherent
// I / Q stored as float32, phase_error = arctan2(imag, |real| + 1e-10), second-order loop filter, phase wrap — all in
// float64, state (phase, frequency, integrator) carried by the PLL object. Each sample depends on the previous one, so a
// sequence is one thread; sequences (channels x chunks) are independent: 32 per warp, rows staged 64 samples at a time
// through shared memory so that global traffic is coalesced (the layout of iir_seq_kernel). numpy's complex product
// rounds every product and sum separately: __dmul_rn / __dadd_rn keep the compiler from contracting them.
// EXACT = true replays that arithmetic in float64 (coherent I / Q bit-equal to the reference on the goldens; ~440 ns per
// sample, the latency of a float64 sincos + atan2 chain). EXACT = false keeps the loop state (phase, integrator) and its
// updates in float64 but evaluates the oscillator, the mixer and the phase detector in float32 (MUFU sine / cosine, abs
// error 4e-7 on [-pi, pi]; degree-13 atan2, 6.5e-7 relative): the detector error enters the phase scaled by alpha ~ 1e-3 and
// the loop tracks it out, the oscillator error is 4e-7 of the coherent output — three orders inside the 1e-4 gate.
namespace {
constexpr int SAM_SEG = 32;
template <bool EXACT>
__global__ void __launch_bounds__(32) sam_pll_kernel(const float2* __restrict__ iq, long long seq_stride, int n, int n_seq,
                                                     double alpha, double beta, int sideband,   // 0 dsb, 1 usb, 2 lsb
                                                     double* __restrict__ state,                // [n_seq][3] phase, frequency, integrator
                                                     float* __restrict__ audio, float* __restrict__ coh_i,
                                                     float* __restrict__ coh_q) {
    __shared__ float2 tin[2][32][SAM_SEG + 1];
    __shared__ float to[3][32][SAM_SEG + 1];
    const int lane = threadIdx.x;
    const int seq0 = blockIdx.x * 32;
    const int rows = min(32, n_seq - seq0);
    double phase = 0.0, freq = 0.0, integ = 0.0;
    if (lane < rows) {
        phase = state[(seq0 + lane) * 3 + 0];
        freq = state[(seq0 + lane) * 3 + 1];
        integ = state[(seq0 + lane) * 3 + 2];
    }
    const double PI = 3.141592653589793, TWO_PI = 6.283185307179586;
    // segment t0 of every row -> tin[b] as asynchronous copies: the next segment lands while this one runs through the loop
    auto stage = [&](int b, int t0) {
        const int cnt = min(SAM_SEG, n - t0);
        for (int r = 0; r < rows; ++r) {
            const float2* xs = iq + (long long)(seq0 + r) * seq_stride + t0;
            for (int e = lane; e < cnt; e += 32) cp_async8(&tin[b][r][e], xs + e);
        }
        cp_async_commit();
    };
    if (n > 0) stage(0, 0);
    int b = 0;
    for (int t0 = 0; t0 < n; t0 += SAM_SEG, b ^= 1) {
        const int cnt = min(SAM_SEG, n - t0);
        if (t0 + SAM_SEG < n) {
            stage(b ^ 1, t0 + SAM_SEG);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        if (lane < rows) {
            for (int j = 0; j < cnt; ++j) {
                const float2 x = tin[b][lane][j];
                float ci, cq;
                double pe;
                if (EXACT) {
                    double sn, cs;
                    sincos(phase, &sn, &cs);                  // lo = exp(-1j * phase) = (cos phase, -sin phase)
                    const double xr = (double)x.x, xi = (double)x.y, lr = cs, li = -sn;
                    const double re = __dsub_rn(__dmul_rn(xr, lr), __dmul_rn(xi, li));
                    const double im = __dadd_rn(__dmul_rn(xr, li), __dmul_rn(xi, lr));
                    ci = (float)re;
                    cq = (float)im;
                    pe = atan2(im, fabs(re) + 1e-10);
                } else {
                    float sn, cs;
                    __sincosf((float)phase, &sn, &cs);
                    ci = fmaf(x.x, cs, x.y * sn);
                    cq = fmaf(x.y, cs, -x.x * sn);
                    pe = (double)fast_atan2f(cq, fabsf(ci) + 1e-10f);
                }
                integ = __dadd_rn(integ, __dmul_rn(beta, pe));
                const double fc = __dadd_rn(__dmul_rn(alpha, pe), integ);
                freq = fc;
                phase = __dadd_rn(phase, fc);
                if (phase > PI) phase = __dsub_rn(phase, TWO_PI);
                else if (phase < -PI) phase = __dadd_rn(phase, TWO_PI);
                to[0][lane][j] = sideband == 1 ? __fadd_rn(ci, cq) : sideband == 2 ? __fsub_rn(ci, cq) : ci;
                to[1][lane][j] = ci;
                to[2][lane][j] = cq;
            }
        }
        __syncwarp();
        for (int r = 0; r < rows; ++r) {
            const long long o = (long long)(seq0 + r) * n + t0;
            for (int e = lane; e < cnt; e += 32) {
                if (audio) audio[o + e] = to[0][r][e];
                if (coh_i) coh_i[o + e] = to[1][r][e];
                if (coh_q) coh_q[o + e] = to[2][r][e];
            }
        }
        __syncwarp();
    }
    if (lane < rows) {
        state[(seq0 + lane) * 3 + 0] = phase;
        state[(seq0 + lane) * 3 + 1] = freq;
        state[(seq0 + lane) * 3 + 2] = integ;
    }
}
}  // namespace

extern "C" {
/* CarrierRecoveryPLL.process for n_seq independent sequences of n complex64 samples (row r at iq_dev + r * seq_stride).
 * alpha / beta: the loop coefficients of dsp/sam.py:63-65; state_dev float64 [n_seq][3] = (phase, frequency, integrator),
 * read and updated (zeros = a fresh PLL, which is what sam_demod builds when no pll_state is passed). Outputs, any of them
 * optional, float32 [n_seq][n]: audio_dev = the sideband selection of sam_demod (:214-221: 0 dsb = I, 1 usb = I + Q,
 * 2 lsb = I - Q), coh_i_dev / coh_q_dev = the coherent components process() returns. exact != 0: float64 replay of the
 * reference's arithmetic; 0: float32 oscillator / mixer / phase detector around the float64 loop state (5x faster). */
int wc_sam_pll(const void* iq_dev, long long seq_stride, int n, int n_seq, double alpha, double beta, int sideband, int exact,
               double* state_dev, float* audio_dev, float* coh_i_dev, float* coh_q_dev, void* stream) {
    WC_REQUIRE(iq_dev && state_dev, "wc_sam_pll: null argument");
    WC_REQUIRE(sideband >= 0 && sideband <= 2, "wc_sam_pll: sideband must be 0 (dsb), 1 (usb) or 2 (lsb)");
    if (n <= 0 || n_seq <= 0) return 0;
    const float2* iq = reinterpret_cast<const float2*>(iq_dev);
    const dim3 grid((n_seq + 31) / 32);
    if (exact)
        sam_pll_kernel<true><<<grid, 32, 0, (cudaStream_t)stream>>>(iq, seq_stride, n, n_seq, alpha, beta, sideband, state_dev, audio_dev,
                                                                    coh_i_dev, coh_q_dev);
    else
        sam_pll_kernel<false><<<grid, 32, 0, (cudaStream_t)stream>>>(iq, seq_stride, n, n_seq, alpha, beta, sideband, state_dev, audio_dev,
                                                                     coh_i_dev, coh_q_dev);
    WC_CUDA(cudaGetLastError());
    return 0;
}
}  // extern "C"

// =================================================================================================
// Analog plan: the whole per-chunk analog chain of a capture as ONE call (SURVEY §8b wc_analog_plan_create / wc_analog_run)
// =================================================================================================
// The caller this replaces makes one _process_channel_dsp_stateless call per (chunk, channel) from its worker pool
// (capture.py:2489-2597). The stage-level entry points above reproduce that chain operator by operator, but chaining them
// from Python costs ~40-50 % of a C1/C2 step in host orchestration (tensor allocation, ctypes, handle lookup). A plan
// owns everything the chain needs — per-channel oscillator parameters on the device, IIR / resampler handles, scratch,
// statistics — and wc_analog_run enqueues the whole batch of chunks: front end -> [IIR stages] -> [AGC] -> RMS ->
// resampler with fused scale / clip -> dB metrics, validity gate and squelch, with no allocation, no host
// synchronisation and no per-stage Python. Calls that repeat the same (input, output, n_chunks) triple are replayed
// from a captured CUDA graph (one cudaGraphLaunch instead of ~10-20 launches).
namespace {

constexpr int PLAN_MAX_IIR = 8;
constexpr int PLAN_GRAPHS = 4;

struct PlanRun {
    int first = 0, count = 0;
    int kind = 0;                 // 0: metrics only (raw / digital / unknown), 1: FM chain, 2: AM / SSB chain
    std::vector<wc_iir*> iir;     // stages between the front end and the gain stage, in order
    bool agc = false;
    wc_iir *agc_attack = nullptr, *agc_release = nullptr;
    float agc_target = 0.f, agc_max_gain = 0.f;
    int up = 1, down = 1;
    wc_resampler* rs = nullptr;
    int n_audio = 0;              // output samples per chunk and channel
    long long audio_off = 0;      // float offset of the run's first channel inside the audio buffer, per chunk count 1
};

struct PlanGraph {
    const void* iq = nullptr;
    float* audio = nullptr;
    float* metrics = nullptr;
    int n_chunks = 0;
    int seen = 0;                 // 1 after an eager run with this key; the next one is captured
    cudaGraphExec_t exec = nullptr;
    unsigned long long stamp = 0;
};

// rssi_db / signal_power_db / valid for every (channel, chunk), squelch decision (capture.py:331-334, 436-437, 323-325,
// 433-435, 2918-2921). metrics: float32 [3][n_ch][n_chunks].
__global__ void plan_metrics_kernel(const double* __restrict__ power, const double* __restrict__ apower,
                                    const int* __restrict__ invalid, const int* __restrict__ nonfinite,
                                    const int* __restrict__ n_audio, const int* __restrict__ kind,
                                    const float* __restrict__ squelch_db, int n, int n_ch, int n_chunks,
                                    float* __restrict__ metrics, unsigned char* __restrict__ squelched) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ch * n_chunks) return;
    const int c = i / n_chunks, b = i % n_chunks;
    const float rssi = 10.0f * log10f((float)(power[i] / (double)n) + 1e-10f);
    float sig = rssi;             // digital modes report the same power twice (capture.py:426-428)
    float valid = nonfinite[b] ? 0.f : 1.f;
    if (kind[c] != 0) {
        sig = 10.0f * log10f((float)(apower[i] / (double)max(n_audio[c], 1)) + 1e-10f);
        if (invalid[i]) valid = valid != 0.f ? 0.5f : 0.f;   // 0.5: RSSI is reported, the audio is dropped (:433-435)
    } else if (kind[c] == 0 && n_audio[c] < 0) {
        valid = valid != 0.f ? 0.5f : 0.f;                   // unknown mode: no audio path
    }
    metrics[i] = rssi;
    metrics[(size_t)n_ch * n_chunks + i] = sig;
    metrics[2 * (size_t)n_ch * n_chunks + i] = valid;
    const float sq = squelch_db[c];
    squelched[i] = (sq == sq && rssi < sq) ? 1 : 0;          // NaN = no squelch
}

// zero the audio of squelched sequences (capture.py:2918-2921); rows: [count * n_chunks][n_audio] starting at audio
__global__ void plan_squelch_kernel(float* __restrict__ audio, int n_audio, const unsigned char* __restrict__ squelched) {
    if (!squelched[blockIdx.y]) return;
    float* a = audio + (long long)blockIdx.y * n_audio;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_audio; i += gridDim.x * blockDim.x) a[i] = 0.f;
}

// the resampler's epilogue for chains whose audio rate equals the capture rate (resample_poly returns its input,
// dsp/fm.py:202-203): scale / clip, audio power and validity per sequence. epi as WC_EPI_*.
__global__ void plan_tail_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int epi,
                                 const double* __restrict__ sumsq, float target_rms, float min_rms,
                                 double* __restrict__ power, int* __restrict__ invalid, float max_abs) {
    __shared__ double red[8];
    __shared__ int bad_s;
    const int seq = blockIdx.y;
    if (threadIdx.x == 0) bad_s = 0;
    __syncthreads();
    float scale = 1.f;
    if (epi == WC_EPI_RMS_CLIP || epi == WC_EPI_RMS) {
        const float rms = sqrtf((float)(sumsq[seq] / (double)n));
        if (rms > min_rms) scale = (float)((double)target_rms / (double)rms);
    }
    double acc = 0.0;
    int bad = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float v = x[(long long)seq * n + i] * scale;
        if (epi == WC_EPI_RMS_CLIP || epi == WC_EPI_CLIP) v = soft_clip_fm(v);
        else if (epi == WC_EPI_CLIP_AGC) v = soft_clip_agc(v);
        y[(long long)seq * n + i] = v;
        acc += (double)v * (double)v;
        if (!(fabsf(v) <= max_abs)) bad = 1;
    }
    acc = warp_sum(acc);
    if (bad) bad_s = 1;
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        atomicAdd(power + seq, s);
        if (bad_s) invalid[seq] = 1;
    }
}

}  // namespace

struct wc_analog_plan {
    int fs = 0, n = 0, fmt = 0, n_ch = 0;
    std::vector<FrontChan> h_chan;
    std::vector<float> h_squelch;
    std::vector<int> h_kind, h_naudio;
    std::vector<PlanRun> runs;
    bool finished = false;
    bool all_fm_shared = false;   // every channel is a shifted FM channel: the front end shares one discriminator (front_groups)
    // device state
    FrontChan* d_chan = nullptr;
    float* d_squelch = nullptr;
    int *d_kind = nullptr, *d_naudio = nullptr;
    int cap_chunks = 0;           // buffers below are sized for this many chunks
    float* d_front = nullptr;     // [n_ch][B][n]
    float *d_env_a = nullptr, *d_env_r = nullptr;   // AGC envelopes, sized for the widest AGC run
    double *d_power = nullptr, *d_ss = nullptr, *d_apower = nullptr;
    int *d_invalid = nullptr, *d_nonfinite = nullptr;
    unsigned char* d_squelched = nullptr;
    void* d_iir_scratch = nullptr;
    long long audio_per_chunk = 0;  // floats of audio per chunk over all channels
    PlanGraph graphs[PLAN_GRAPHS];
    unsigned long long clock = 0;
    int use_graph = 1;
};

static void plan_free_buffers(wc_analog_plan* p) {
    for (void* q : {(void*)p->d_front, (void*)p->d_env_a, (void*)p->d_env_r, (void*)p->d_power, (void*)p->d_ss, (void*)p->d_apower,
                    (void*)p->d_invalid, (void*)p->d_nonfinite, (void*)p->d_squelched, p->d_iir_scratch})
        if (q) cudaFree(q);
    p->d_front = p->d_env_a = p->d_env_r = nullptr;
    p->d_power = p->d_ss = p->d_apower = nullptr;
    p->d_invalid = p->d_nonfinite = nullptr;
    p->d_squelched = nullptr;
    p->d_iir_scratch = nullptr;
    for (PlanGraph& g : p->graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g = PlanGraph();
    }
    p->cap_chunks = 0;
}

static int plan_reserve(wc_analog_plan* p, int B) {
    if (B <= p->cap_chunks) return 0;
    plan_free_buffers(p);
    const size_t seqs = (size_t)p->n_ch * B;
    size_t agc_rows = 0, scratch = 16;
    for (const PlanRun& r : p->runs) {
        if (r.agc) agc_rows = std::max(agc_rows, (size_t)r.count * B);
        for (const wc_iir* f : r.iir)
            if (!f->sequential) scratch = std::max(scratch, iir_scan_scratch_bytes(f->h_cf.K, p->n, r.count * B));
        if (r.agc) scratch = std::max(scratch, iir_scan_scratch_bytes(1, p->n, r.count * B));
    }
    WC_CUDA(cudaMalloc(&p->d_front, sizeof(float) * seqs * (size_t)p->n));
    if (agc_rows) {
        WC_CUDA(cudaMalloc(&p->d_env_a, sizeof(float) * agc_rows * (size_t)p->n));
        WC_CUDA(cudaMalloc(&p->d_env_r, sizeof(float) * agc_rows * (size_t)p->n));
    }
    WC_CUDA(cudaMalloc(&p->d_power, sizeof(double) * seqs));
    WC_CUDA(cudaMalloc(&p->d_ss, sizeof(double) * seqs));
    WC_CUDA(cudaMalloc(&p->d_apower, sizeof(double) * seqs));
    WC_CUDA(cudaMalloc(&p->d_invalid, sizeof(int) * seqs));
    WC_CUDA(cudaMalloc(&p->d_nonfinite, sizeof(int) * (size_t)B));
    WC_CUDA(cudaMalloc(&p->d_squelched, seqs));
    WC_CUDA(cudaMalloc(&p->d_iir_scratch, scratch));
    p->cap_chunks = B;
    return 0;
}

// everything wc_analog_run enqueues, on `st` (capturable: no allocation, no synchronisation, no host-side parameters)
static int plan_enqueue(wc_analog_plan* p, const void* iq_dev, int B, float* audio_dev, float* metrics_dev, cudaStream_t st) {
    const int n = p->n, C = p->n_ch;
    const size_t seqs = (size_t)C * B;
    WC_CUDA(cudaMemsetAsync(p->d_power, 0, sizeof(double) * seqs, st));
    WC_CUDA(cudaMemsetAsync(p->d_ss, 0, sizeof(double) * seqs, st));
    WC_CUDA(cudaMemsetAsync(p->d_apower, 0, sizeof(double) * seqs, st));
    WC_CUDA(cudaMemsetAsync(p->d_invalid, 0, sizeof(int) * seqs, st));
    WC_CUDA(cudaMemsetAsync(p->d_nonfinite, 0, sizeof(int) * (size_t)B, st));
    {
        FrontArgs a;
        a.iq = iq_dev;
        a.chunk_stride = n;
        a.n = n;
        a.fmt = p->fmt;
        a.n_ch = C;
        a.n_chunks = B;
        a.ch = p->d_chan;
        a.out = p->d_front;
        a.base_out = nullptr;
        a.power = p->d_power;
        a.out_sumsq = p->d_ss;
        a.nonfinite = p->d_nonfinite;
        const int tiles = (n + FR_TILE - 1) / FR_TILE;
        const long long base_ctas = (long long)tiles * B;
        const int groups = front_groups(base_ctas, C, p->all_fm_shared);
        WC_CUDA(smem_optin(front_kernel, (int)sizeof(FrontSmem), g_front_optin));
        front_kernel<<<dim3(tiles, B, (unsigned)groups), FR_THREADS, sizeof(FrontSmem), st>>>(a);
        WC_CUDA(cudaGetLastError());
    }
    for (const PlanRun& r : p->runs) {
        if (r.kind == 0) continue;
        const int rows = r.count * B;
        float* x = p->d_front + (size_t)r.first * B * n;
        double* ss = p->d_ss + (size_t)r.first * B;
        double* ap = p->d_apower + (size_t)r.first * B;
        int* inv = p->d_invalid + (size_t)r.first * B;
        bool ss_done = r.iir.empty();   // the front end's sum(out**2) holds only while nothing has changed the signal
        for (size_t k = 0; k < r.iir.size(); ++k) {
            const wc_iir* f = r.iir[k];
            const bool last = (k + 1 == r.iir.size()) && r.kind == 1;
            if (f->h_cf.K == 0) {
                elementwise_kernel<<<148 * 8, 256, 0, st>>>(x, x, (long long)rows * n, 2, (float)f->h_cf.b0);
            } else {
                if (last) WC_CUDA(cudaMemsetAsync(ss, 0, sizeof(double) * rows, st));
                if (int rc = iir_run(f, x, x, n, n, rows, 0, p->d_iir_scratch, st, last ? ss : nullptr, last ? &ss_done : nullptr))
                    return rc;
            }
        }
        int epi;
        if (r.kind == 1) {
            epi = WC_EPI_RMS_CLIP;
            if (!ss_done) {   // the last stage could not fuse it (sequential replay / pure gain)
                WC_CUDA(cudaMemsetAsync(ss, 0, sizeof(double) * rows, st));
                int bx = (n + 255) / 256;
                if (bx > 64) bx = 64;
                sumsq_kernel<<<dim3(bx, rows), 256, 0, st>>>(x, n, n, ss);
            }
        } else {
            if (r.agc) {
                if (int rc = iir_run(r.agc_attack, x, p->d_env_a, n, n, rows, 1, p->d_iir_scratch, st)) return rc;
                if (int rc = iir_run(r.agc_release, p->d_env_a, p->d_env_r, n, n, rows, 0, p->d_iir_scratch, st)) return rc;
                agc_apply_kernel<<<148 * 8, 256, 0, st>>>(x, p->d_env_a, p->d_env_r, x, (long long)rows * n, r.agc_target, r.agc_max_gain);
                epi = WC_EPI_NONE;
            } else {
                epi = WC_EPI_CLIP_AGC;
            }
        }
        float* au = audio_dev + r.audio_off * B;
        if (r.rs) {
            if (int rc = wc_resampler_run(r.rs, x, n, n, rows, au, epi, ss, 0.18f, 1e-4f, ap, inv, 1.2f, st)) return rc;
        } else {
            int bx = (n + 1023) / 1024;
            if (bx > 32) bx = 32;
            plan_tail_kernel<<<dim3(bx, rows), 256, 0, st>>>(x, au, n, epi, ss, 0.18f, 1e-4f, ap, inv, 1.2f);
        }
        WC_CUDA(cudaGetLastError());
    }
    plan_metrics_kernel<<<(unsigned)((seqs + 127) / 128), 128, 0, st>>>(p->d_power, p->d_apower, p->d_invalid, p->d_nonfinite, p->d_naudio,
                                                                         p->d_kind, p->d_squelch, n, C, B, metrics_dev, p->d_squelched);
    for (const PlanRun& r : p->runs) {
        if (r.kind == 0 || r.n_audio <= 0) continue;
        bool any = false;
        for (int c = r.first; c < r.first + r.count; ++c) any = any || (p->h_squelch[c] == p->h_squelch[c]);
        if (!any) continue;
        int bx = (r.n_audio + 255) / 256;
        if (bx > 8) bx = 8;
        plan_squelch_kernel<<<dim3(bx, r.count * B), 256, 0, st>>>(audio_dev + r.audio_off * B, r.n_audio, p->d_squelched + (size_t)r.first * B);
    }
    WC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" {

int wc_analog_plan_create(int sample_rate, int chunk_len, int in_fmt, int n_channels, const int* modes, const double* offsets_hz,
                          const double* bfo_hz, const float* squelch_db, wc_analog_plan** out) {
    WC_REQUIRE(out && modes && offsets_hz, "wc_analog_plan_create: null argument");
    WC_REQUIRE(sample_rate > 0 && chunk_len >= 1 && n_channels >= 1 && n_channels <= 4096, "wc_analog_plan_create: bad sizes");
    WC_REQUIRE(in_fmt == 0 || in_fmt == 1, "wc_analog_plan_create: in_fmt must be 0 (cf32) or 1 (cs16)");
    wc_analog_plan* p = new wc_analog_plan();
    p->fs = sample_rate;
    p->n = chunk_len;
    p->fmt = in_fmt;
    p->n_ch = n_channels;
    p->h_chan.resize(n_channels);
    p->h_squelch.assign(n_channels, nanf(""));
    p->h_kind.assign(n_channels, 0);
    p->h_naudio.assign(n_channels, 0);
    for (int c = 0; c < n_channels; ++c) {
        FrontChan& f = p->h_chan[c];
        f.mode = modes[c];
        f.shift = (offsets_hz[c] != 0.0) ? 1 : 0;
        const double off = nearbyint(offsets_hz[c]);
        f.k32 = (float)(-(2.0 * M_PI * (off / (double)sample_rate)));
        f.bfo_turns = bfo_hz ? bfo_hz[c] / (double)sample_rate : 0.0;
        f.disc_scale = (float)((double)sample_rate / (2.0 * M_PI * 75000.0));
        if (squelch_db) p->h_squelch[c] = squelch_db[c];
    }
    p->all_fm_shared = true;
    for (int c = 0; c < n_channels; ++c)
        p->all_fm_shared = p->all_fm_shared && (modes[c] == WC_MODE_WBFM || modes[c] == WC_MODE_NBFM) && offsets_hz[c] != 0.0;
    *out = p;
    return 0;
}

// A run = `count` adjacent channels starting at `first` that share one chain. kind 0: metrics only; 1: FM (rms_normalize +
// fm soft clip); 2: AM / SSB. up/down/taps: scipy resample_poly's design (up == down == 1: no resampler). Returns the run index.
int wc_analog_plan_add_run(wc_analog_plan* p, int first, int count, int kind, int up, int down, const double* taps, int n_taps) {
    WC_REQUIRE(p && !p->finished, "wc_analog_plan_add_run: plan missing or already finished");
    WC_REQUIRE(first >= 0 && count >= 1 && first + count <= p->n_ch && kind >= 0 && kind <= 2, "wc_analog_plan_add_run: bad run");
    PlanRun r;
    r.first = first;
    r.count = count;
    r.kind = kind;
    r.up = up;
    r.down = down;
    if (kind != 0) {
        if (up != 1 || down != 1) {
            WC_REQUIRE(taps && n_taps >= 1, "wc_analog_plan_add_run: resampler taps missing");
            if (int rc = wc_resampler_create(up, down, taps, n_taps, &r.rs)) return rc < 0 ? rc : -rc;
            r.n_audio = (int)wc_resampler_out_len(r.rs, p->n);
        } else {
            r.n_audio = p->n;
        }
    }
    p->runs.push_back(r);
    return (int)p->runs.size() - 1;
}

int wc_analog_plan_add_iir(wc_analog_plan* p, int run, const double* b, int nb, const double* a, int na) {
    WC_REQUIRE(p && !p->finished && run >= 0 && run < (int)p->runs.size(), "wc_analog_plan_add_iir: bad plan / run");
    WC_REQUIRE((int)p->runs[run].iir.size() < PLAN_MAX_IIR, "wc_analog_plan_add_iir: too many stages");
    wc_iir* f = nullptr;
    if (int rc = wc_iir_create(b, nb, a, na, &f)) return rc;
    p->runs[run].iir.push_back(f);
    return 0;
}

// apply_agc (dsp/agc.py:169-242): the two float32 one-pole envelope filters as (b, a) pairs, target and maximum gain (linear)
int wc_analog_plan_set_agc(wc_analog_plan* p, int run, const double* attack_b, const double* attack_a, const double* release_b,
                           const double* release_a, float target_linear, float max_gain_linear) {
    WC_REQUIRE(p && !p->finished && run >= 0 && run < (int)p->runs.size() && p->runs[run].kind == 2, "wc_analog_plan_set_agc: bad plan / run");
    PlanRun& r = p->runs[run];
    if (int rc = wc_iir_create(attack_b, 1, attack_a, 2, &r.agc_attack)) return rc;
    if (int rc = wc_iir_create(release_b, 1, release_a, 2, &r.agc_release)) return rc;
    r.agc = true;
    r.agc_target = target_linear;
    r.agc_max_gain = max_gain_linear;
    return 0;
}

int wc_analog_plan_finish(wc_analog_plan* p) {
    WC_REQUIRE(p && !p->finished, "wc_analog_plan_finish: plan missing or already finished");
    std::vector<int> covered(p->n_ch, 0);
    long long off = 0;
    for (PlanRun& r : p->runs) {
        r.audio_off = off;
        off += (long long)r.count * r.n_audio;
        for (int c = r.first; c < r.first + r.count; ++c) {
            WC_REQUIRE(!covered[c], "wc_analog_plan_finish: channel %d is in two runs", c);
            covered[c] = 1;
            p->h_kind[c] = r.kind;
            p->h_naudio[c] = r.n_audio;
        }
    }
    for (int c = 0; c < p->n_ch; ++c) WC_REQUIRE(covered[c], "wc_analog_plan_finish: channel %d is in no run", c);
    p->audio_per_chunk = off;
    WC_CUDA(cudaMalloc(&p->d_chan, sizeof(FrontChan) * p->n_ch));
    WC_CUDA(cudaMalloc(&p->d_squelch, sizeof(float) * p->n_ch));
    WC_CUDA(cudaMalloc(&p->d_kind, sizeof(int) * p->n_ch));
    WC_CUDA(cudaMalloc(&p->d_naudio, sizeof(int) * p->n_ch));
    WC_CUDA(cudaMemcpy(p->d_chan, p->h_chan.data(), sizeof(FrontChan) * p->n_ch, cudaMemcpyHostToDevice));
    WC_CUDA(cudaMemcpy(p->d_squelch, p->h_squelch.data(), sizeof(float) * p->n_ch, cudaMemcpyHostToDevice));
    WC_CUDA(cudaMemcpy(p->d_kind, p->h_kind.data(), sizeof(int) * p->n_ch, cudaMemcpyHostToDevice));
    WC_CUDA(cudaMemcpy(p->d_naudio, p->h_naudio.data(), sizeof(int) * p->n_ch, cudaMemcpyHostToDevice));
    p->finished = true;
    return 0;
}

void wc_analog_plan_destroy(wc_analog_plan* p) {
    if (!p) return;
    plan_free_buffers(p);
    for (PlanRun& r : p->runs) {
        for (wc_iir* f : r.iir) wc_iir_destroy(f);
        wc_iir_destroy(r.agc_attack);
        wc_iir_destroy(r.agc_release);
        wc_resampler_destroy(r.rs);
    }
    for (void* q : {(void*)p->d_chan, (void*)p->d_squelch, (void*)p->d_kind, (void*)p->d_naudio})
        if (q) cudaFree(q);
    delete p;
}

/* floats of audio per chunk over all channels; channel c's audio of chunk b of a B-chunk call starts at
 * audio + B * wc_analog_plan_audio_offset(c) + b * wc_analog_plan_audio_len(c). */
long long wc_analog_plan_audio_floats(const wc_analog_plan* p) { return p ? p->audio_per_chunk : 0; }
int wc_analog_plan_audio_len(const wc_analog_plan* p, int channel) {
    return (p && channel >= 0 && channel < p->n_ch) ? p->h_naudio[channel] : 0;
}
long long wc_analog_plan_audio_offset(const wc_analog_plan* p, int channel) {
    if (!p || channel < 0 || channel >= p->n_ch) return -1;
    for (const PlanRun& r : p->runs)
        if (channel >= r.first && channel < r.first + r.count) return r.audio_off + (long long)(channel - r.first) * r.n_audio;
    return -1;
}
int wc_analog_plan_use_graph(wc_analog_plan* p, int on) {
    WC_REQUIRE(p != nullptr, "wc_analog_plan_use_graph: null plan");
    p->use_graph = on ? 1 : 0;
    return 0;
}

/* n_chunks consecutive chunks of chunk_len samples at iq_dev (cf32 or cs16 as configured).
 * audio_dev: float32 [n_chunks * wc_analog_plan_audio_floats()] (layout above; squelched sequences are zeroed);
 * metrics_dev: float32 [3][n_channels][n_chunks] = rssi_db | signal_power_db | valid (1 = audio valid, 0.5 = RSSI only:
 * the audio failed the validity gate or the mode has no audio path, 0 = chunk dropped for non-finite IQ). */
int wc_analog_run(wc_analog_plan* p, const void* iq_dev, int n_chunks, float* audio_dev, float* metrics_dev, void* stream) {
    WC_REQUIRE(p && p->finished && iq_dev && metrics_dev, "wc_analog_run: null argument or unfinished plan");
    WC_REQUIRE(audio_dev || p->audio_per_chunk == 0, "wc_analog_run: audio_dev is null");
    WC_REQUIRE(n_chunks >= 1, "wc_analog_run: n_chunks must be >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    if (plan_reserve(p, n_chunks)) return -2;
    if (!p->use_graph || st == nullptr || st == cudaStreamLegacy) return plan_enqueue(p, iq_dev, n_chunks, audio_dev, metrics_dev, st);
    PlanGraph* slot = nullptr;
    for (PlanGraph& g : p->graphs)
        if (g.seen && g.iq == iq_dev && g.audio == audio_dev && g.metrics == metrics_dev && g.n_chunks == n_chunks) slot = &g;
    ++p->clock;
    if (slot && slot->exec) {
        slot->stamp = p->clock;
        WC_CUDA(cudaGraphLaunch(slot->exec, st));
        return 0;
    }
    if (!slot) {   // first sight of this key: run eagerly (also sets the per-device function attributes), remember it
        PlanGraph* lru = &p->graphs[0];
        for (PlanGraph& g : p->graphs)
            if (g.stamp < lru->stamp) lru = &g;
        if (lru->exec) cudaGraphExecDestroy(lru->exec);
        *lru = PlanGraph();
        lru->iq = iq_dev;
        lru->audio = audio_dev;
        lru->metrics = metrics_dev;
        lru->n_chunks = n_chunks;
        lru->seen = 1;
        lru->stamp = p->clock;
        return plan_enqueue(p, iq_dev, n_chunks, audio_dev, metrics_dev, st);
    }
    // second call with the same key: capture it, then launch the graph
    slot->stamp = p->clock;
    cudaGraph_t graph = nullptr;
    WC_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = plan_enqueue(p, iq_dev, n_chunks, audio_dev, metrics_dev, st);
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        p->use_graph = 0;   // capture is not available here: stay on the eager path
        return rc ? rc : plan_enqueue(p, iq_dev, n_chunks, audio_dev, metrics_dev, st);
    }
    const cudaError_t ie = cudaGraphInstantiate(&slot->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
        slot->exec = nullptr;
        cudaGetLastError();
        p->use_graph = 0;
        return plan_enqueue(p, iq_dev, n_chunks, audio_dev, metrics_dev, st);
    }
    WC_CUDA(cudaGraphLaunch(slot->exec, st));
    return 0;
}

}  // extern "C"
