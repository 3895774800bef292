// Control-channel scanner measurement, all candidate frequencies of one wideband block per call:
//   trunking/cc_scanner.py:166-277  _measure_channel: capture.freq_shift, firwin(65, 0.8/D, kaiser 6) lfilter from zero
//                                   state, [::D], mean and max of |y|^2 (signal and the two band-edge noise probes)
//   trunking/cc_scanner.py:279-351  _detect_sync_pattern: angle(y[1:] conj(y[:-1])) at [5::10], normalised correlation
//                                   with the +-0.2356 sync waveform at every symbol offset
//
// One thread per KEPT output (the filter is only evaluated where [::D] keeps a sample): 65 NCO evaluations and complex
// MACs in float64 on the complex64 product, like lfilter does on freq_shift's complex64 result. The wideband block is
// read once per candidate straight from L2 (a 100 ms block at 6 MS/s is 4.8 MB).
#include <math.h>
#include <vector>

#include "../../include/wcsdr_b200.h"
#include "common.cuh"

namespace wc {

constexpr int CS_TAPS = 65;

struct ScanArgs {
    const float2* iq;
    int n, decim, m;          // m = ceil(n / decim) kept samples
    const double* taps;       // [65]
    const float* k32;         // [K] per candidate, 0-offset candidates carry shift 0
    const int* shift;         // [K]
    double2* y;               // [K][m]
    double* power_sum;        // [K]
    unsigned long long* power_max_bits;  // [K] max |y|^2 as ordered bits of a non-negative double
};

__global__ void __launch_bounds__(128) ccscan_ddc_kernel(const ScanArgs a) {
    __shared__ double sh[CS_TAPS];
    __shared__ double red_s[4];
    __shared__ double red_m[4];
    if (threadIdx.x < CS_TAPS) sh[threadIdx.x] = a.taps[threadIdx.x];
    __syncthreads();
    const int c = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const float k32 = a.k32[c];
    const int shift = a.shift[c];
    double p = 0.0;
    if (j < a.m) {
        double yr = 0.0, yi = 0.0;
        if (a.decim > 1) {
            const int n0 = j * a.decim;
#pragma unroll 5
            for (int t = 0; t < CS_TAPS; ++t) {
                const int n = n0 - t;
                if (n < 0) break;
                float2 v = a.iq[n];
                if (shift) {
                    float cs, sn;
                    nco_f32(k32, n, cs, sn);
                    v = make_float2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
                }
                yr = fma(sh[t], (double)v.x, yr);
                yi = fma(sh[t], (double)v.y, yi);
            }
        } else {
            float2 v = a.iq[j];
            if (shift) {
                float cs, sn;
                nco_f32(k32, j, cs, sn);
                v = make_float2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
            }
            yr = v.x;
            yi = v.y;
        }
        a.y[(long long)c * a.m + j] = make_double2(yr, yi);
        const double mag = hypot(yr, yi);   // np.abs(complex128)
        p = mag * mag;
    }
    double s = warp_sum(p);
    double mx = p;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) {
        red_s[threadIdx.x >> 5] = s;
        red_m[threadIdx.x >> 5] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        s = red_s[0] + red_s[1] + red_s[2] + red_s[3];
        mx = fmax(fmax(red_m[0], red_m[1]), fmax(red_m[2], red_m[3]));
        atomicAdd(a.power_sum + c, s);
        atomicMax(a.power_max_bits + c, (unsigned long long)__double_as_longlong(mx));
    }
}

// best normalised correlation per candidate; one CTA per candidate, one thread per symbol offset
__global__ void __launch_bounds__(256) ccscan_sync_kernel(const double2* __restrict__ y, int m, double* __restrict__ corr) {
    __shared__ double best_abs[256];
    __shared__ double best_val[256];
    __shared__ int best_idx[256];
    const int c = blockIdx.x;
    const double2* yc = y + (long long)c * m;
    const int L = m - 1;                       // fm_demod length
    const int count = L / 10;
    const int n_sym = min((L - 5 + 9) / 10, count);   // len(fm[5::10][:count])
    int search = min(n_sym - 24, count - 24);
    if (m < 10 * 24 + 10 || count < 24 || search <= 0) {
        if (threadIdx.x == 0) corr[c] = 0.0;
        return;
    }
    const unsigned long long pat = 0x5575F5FF77FFull;   // dibit 1 -> +dev, dibit 3 -> -dev (cc_scanner.py:104-106)
    double ba = 0.0, bv = 0.0;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < search; i += blockDim.x) {
        double dot = 0.0, e = 0.0;
        for (int k = 0; k < 24; ++k) {
            const int q = 5 + 10 * (i + k);
            const double2 p1 = yc[q + 1], p0 = yc[q];
            // angle(p1 * conj(p0))
            const double re = p1.x * p0.x + p1.y * p0.y, im = p1.y * p0.x - p1.x * p0.y;
            const double f = atan2(im, re);
            const double w = (((pat >> ((23 - k) * 2)) & 3ull) == 1ull) ? 0.2356 : -0.2356;
            dot += f * w;
            e += f * f;
        }
        const double v = dot / (sqrt(e + 1e-10) * sqrt(24.0 * 0.2356 * 0.2356));
        if (fabs(v) > ba) {   // strict: an earlier offset with the same |corr| stays (this thread scans upwards)
            ba = fabs(v);
            bv = v;
            bi = i;
        }
    }
    best_abs[threadIdx.x] = ba;
    best_val[threadIdx.x] = bv;
    best_idx[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const int u = threadIdx.x + o;
            if (best_abs[u] > best_abs[threadIdx.x] || (best_abs[u] == best_abs[threadIdx.x] && best_idx[u] < best_idx[threadIdx.x])) {
                best_abs[threadIdx.x] = best_abs[u];
                best_val[threadIdx.x] = best_val[u];
                best_idx[threadIdx.x] = best_idx[u];
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) corr[c] = best_val[0];
}

}  // namespace wc

using namespace wc;

extern "C" {

int wc_ccscan_out_len(int n_samples, int sample_rate) {
    const int d = sample_rate / 48000 > 1 ? sample_rate / 48000 : 1;
    return n_samples <= 0 ? 0 : (n_samples + d - 1) / d;
}

/* One call = _measure_channel's arithmetic for n_ch frequency offsets of one block (cc_scanner.py:166-351): iq complex64
 * [n]; offsets_hz / taps65 on the host (taps = firwin(65, 0.8/D, kaiser 6.0), D = max(1, fs // 48000));
 * y_dev complex128 [n_ch][m] (m = wc_ccscan_out_len), power_sum_dev / power_max_dev float64 [n_ch] (sum and max of |y|^2),
 * corr_dev float64 [n_ch] best normalised sync correlation (0 when the block is too short). scratch_dev: 8*n_ch bytes. */
int wc_ccscan_measure(const void* iq_dev, int n_samples, int sample_rate, const double* offsets_hz, int n_ch,
                      const double* taps65, void* y_dev, double* power_sum_dev, double* power_max_dev, double* corr_dev,
                      void* scratch_dev, void* stream_v) {
    WC_REQUIRE(iq_dev && offsets_hz && taps65 && y_dev && power_sum_dev && power_max_dev && corr_dev && scratch_dev,
               "wc_ccscan_measure: null argument");
    WC_REQUIRE(n_ch >= 1 && n_ch <= 4096 && n_samples >= 0 && sample_rate > 0, "wc_ccscan_measure: bad sizes");
    cudaStream_t st = (cudaStream_t)stream_v;
    const int decim = sample_rate / 48000 > 1 ? sample_rate / 48000 : 1;
    const int m = wc_ccscan_out_len(n_samples, sample_rate);
    WC_CUDA(cudaMemsetAsync(power_sum_dev, 0, sizeof(double) * n_ch, st));
    WC_CUDA(cudaMemsetAsync(power_max_dev, 0, sizeof(double) * n_ch, st));
    WC_CUDA(cudaMemsetAsync(corr_dev, 0, sizeof(double) * n_ch, st));
    if (m == 0) return 0;
    std::vector<float> k32(n_ch);
    std::vector<int> shift(n_ch);
    for (int c = 0; c < n_ch; ++c) {
        shift[c] = offsets_hz[c] != 0.0 ? 1 : 0;   // capture.freq_shift returns its input for a zero offset
        k32[c] = (float)(-(2.0 * M_PI * (nearbyint(offsets_hz[c]) / (double)sample_rate)));
    }
    double* d_taps = nullptr;
    WC_CUDA(cudaMallocAsync((void**)&d_taps, sizeof(double) * CS_TAPS, st));
    float* d_k = reinterpret_cast<float*>(scratch_dev);
    int* d_s = reinterpret_cast<int*>(d_k + n_ch);
    WC_CUDA(cudaMemcpyAsync(d_taps, taps65, sizeof(double) * CS_TAPS, cudaMemcpyHostToDevice, st));
    WC_CUDA(cudaMemcpyAsync(d_k, k32.data(), sizeof(float) * n_ch, cudaMemcpyHostToDevice, st));
    WC_CUDA(cudaMemcpyAsync(d_s, shift.data(), sizeof(int) * n_ch, cudaMemcpyHostToDevice, st));
    ScanArgs a;
    a.iq = reinterpret_cast<const float2*>(iq_dev);
    a.n = n_samples;
    a.decim = decim;
    a.m = m;
    a.taps = d_taps;
    a.k32 = d_k;
    a.shift = d_s;
    a.y = reinterpret_cast<double2*>(y_dev);
    a.power_sum = power_sum_dev;
    a.power_max_bits = reinterpret_cast<unsigned long long*>(power_max_dev);
    ccscan_ddc_kernel<<<dim3((m + 127) / 128, n_ch), 128, 0, st>>>(a);
    ccscan_sync_kernel<<<n_ch, 256, 0, st>>>(a.y, m, corr_dev);
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaFreeAsync(d_taps, st));
    WC_CUDA(cudaStreamSynchronize(st));   // the pageable host vectors must outlive the async copies
    return 0;
}

}  // extern "C"
