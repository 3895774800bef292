// Peer memory for "ONE capture, many GPUs" (SURVEY §8e modes ii/iii): the ingest rank's IQ block is mapped into every
// other rank's address space (CUDA IPC, NVLink P2P) and the consuming kernels — e.g. the channelizer's bulk async
// copies — read their time slab straight out of it, so the transfer overlaps the math tile by tile and no rank receives
// samples it does not process. Flags in the same region order producer and consumers without a host round trip or a
// collective: the producer publishes "block k is in buffer b" with a system-scope release store, consumers spin on it
// with acquire loads (bounded, so a dead peer cannot hang the GPU) and publish "done reading" the same way.
#include <string.h>

#include "../../include/wcsdr_b200.h"
#include "common.cuh"

namespace wc {

__global__ void flag_set_kernel(unsigned* flag, unsigned value) {
    // everything earlier in the stream has completed; make it visible system-wide before the flag
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

__global__ void flag_wait_kernel(const unsigned* __restrict__ flags, int n_flags, long long stride_words, unsigned value,
                                 unsigned long long timeout_ns, int* timed_out) {
    // thread i waits for flags[i * stride_words] >= value (sequence numbers only grow)
    const int i = threadIdx.x;
    if (i >= n_flags) return;
    const unsigned* f = flags + (long long)i * stride_words;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if ((int)(v - value) >= 0) break;
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) {
            if (timed_out) atomicExch(timed_out, 1);
            break;
        }
        __nanosleep(100);
    }
}

}  // namespace wc

extern "C" {

int wc_peer_alloc(long long bytes, void** dev_out, void* handle_out) {
    WC_REQUIRE(bytes > 0 && dev_out && handle_out, "wc_peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == WC_PEER_HANDLE_BYTES, "handle size");
    void* p = nullptr;
    WC_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        wc::set_error("wc_peer_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
        return -2;
    }
    memcpy(handle_out, &h, sizeof(h));
    *dev_out = p;
    return 0;
}

int wc_peer_free(void* dev) {
    if (dev) WC_CUDA(cudaFree(dev));
    return 0;
}

int wc_peer_open(const void* handle, void** dev_out) {
    WC_REQUIRE(handle && dev_out, "wc_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    WC_CUDA(cudaIpcOpenMemHandle(dev_out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int wc_peer_close(void* dev) {
    if (dev) WC_CUDA(cudaIpcCloseMemHandle(dev));
    return 0;
}

int wc_peer_copy(void* dst_dev, const void* src_dev, long long bytes, void* stream) {
    WC_REQUIRE(dst_dev && src_dev && bytes >= 0, "wc_peer_copy: bad arguments");
    WC_CUDA(cudaMemcpyAsync(dst_dev, src_dev, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return 0;
}

int wc_flag_set(unsigned* flag_dev, unsigned value, void* stream) {
    WC_REQUIRE(flag_dev, "wc_flag_set: null flag");
    wc::flag_set_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag_dev, value);
    WC_CUDA(cudaGetLastError());
    return 0;
}

int wc_flag_wait(const unsigned* flags_dev, int n_flags, long long stride_words, unsigned value, int timeout_ms,
                 int* timed_out_dev, void* stream) {
    WC_REQUIRE(flags_dev && n_flags >= 1 && n_flags <= 1024, "wc_flag_wait: bad arguments");
    WC_REQUIRE(timeout_ms > 0 && timeout_ms <= 60000, "wc_flag_wait: timeout_ms must be in (0, 60000]");
    wc::flag_wait_kernel<<<1, ((n_flags + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(
        flags_dev, n_flags, stride_words, value, (unsigned long long)timeout_ms * 1000000ull, timed_out_dev);
    WC_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
