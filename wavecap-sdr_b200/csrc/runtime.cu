// Library-level entry points: init, error string, device info.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/wcsdr_b200.h"
#include "common.cuh"

namespace wc {

static thread_local char g_err[1024] = "";
static int g_sm_count = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

#ifdef WC_DEV
int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}
#endif

int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

}  // namespace wc

extern "C" {

const char* wc_last_error(void) { return wc::get_error(); }
const char* wc_version(void) { return "wcsdr_b200 0.1 (sm_100a)"; }

int wc_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        wc::set_error("wc_init: no CUDA device (%s) — this library has no CPU fallback", cudaGetErrorString(e));
        return -3;
    }
    WC_REQUIRE(device >= 0 && device < n, "wc_init: device %d out of range (have %d)", device, n);
    WC_CUDA(cudaSetDevice(device));
    int major = 0, minor = 0;
    WC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    WC_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    WC_REQUIRE(major == 10, "wc_init: device %d is sm_%d%d; this build carries sm_100a code only", device, major, minor);
    wc::g_sm_count = 0;
    wc::sm_count();
    {
        // stream-ordered scratch (cudaMallocAsync in the stateless entry points) comes from the device's default pool;
        // keep freed blocks in the pool instead of returning them to the driver at every synchronisation
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    return 0;
}

int wc_device_info(int* sm_count, int* cc_major, int* cc_minor, long long* total_mem_bytes) {
    int dev = 0;
    WC_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    WC_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem_bytes) *total_mem_bytes = (long long)p.totalGlobalMem;
    return 0;
}

}  // extern "C"
