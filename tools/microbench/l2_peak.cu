// Measured L2 read bandwidth of this GPU: the ceiling K3b (k3_count_*_kernel, label-matrix row sweeps) is reported against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/l2_peak tools/microbench/l2_peak.cu && tools/microbench/l2_peak
// Every block streams a buffer that fits the L2 (default 48 MB, well inside the 126 MB) over and over with coalesced loads —
// 8 B per lane (256 B per warp instruction: K3b's access shape) and 16 B per lane — with enough loads in flight per thread
// to cover the latency.  Loads are ld.global.cg (cached in L2 only): an L1 hit would not be L2 bandwidth.  Prints one JSON
// line; bytes = what the SMs requested (all of it is served by the L2 after pass 1).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int VEC>   // 1: 8-byte loads, 2: 16-byte loads
__global__ void __launch_bounds__(256) l2_read_kernel(const unsigned long long* __restrict__ buf, size_t n_words, int passes, unsigned long long* sink) {
    unsigned long long acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x * VEC;
    for (int p = 0; p < passes; p++) {
        size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
#pragma unroll 8
        for (; i + VEC <= n_words; i += stride) {
            if (VEC == 1) {
                unsigned long long v;
                asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(buf + i));
                acc ^= v;
            } else {
                unsigned long long a, b;
                asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(buf + i));
                acc ^= a ^ b;
            }
        }
    }
    if (acc == 0x123456789abcdefull) *sink = acc;
}

template <int VEC>
static double run(const unsigned long long* d, size_t n_words, int blocks, int passes, unsigned long long* sink) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    l2_read_kernel<VEC><<<blocks, 256>>>(d, n_words, 2, sink);     // warm the L2
    cudaDeviceSynchronize();
    double best = 0;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        l2_read_kernel<VEC><<<blocks, 256>>>(d, n_words, passes, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double gbs = (double)n_words * 8 * passes / (ms * 1e-3) / 1e9;
        if (gbs > best) best = gbs;
    }
    return best;
}

int main(int argc, char** argv) {
    const size_t mb = argc > 1 ? (size_t)atoll(argv[1]) : 48;
    const size_t n_words = mb * (1 << 20) / 8;
    unsigned long long *d = nullptr, *sink = nullptr;
    cudaMalloc(&d, n_words * 8); cudaMalloc(&sink, 8);
    cudaMemset(d, 1, n_words * 8);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    double best8 = 0, best16 = 0; int b8 = 0, b16 = 0;
    for (int per_sm : {4, 6, 8}) {
        const int blocks = prop.multiProcessorCount * per_sm;
        const double a = run<1>(d, n_words, blocks, 40, sink), b = run<2>(d, n_words, blocks, 40, sink);
        if (a > best8) { best8 = a; b8 = per_sm; }
        if (b > best16) { best16 = b; b16 = per_sm; }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("{\"gpu\": \"%s\", \"buffer_mb\": %zu, \"l2_read_gbs_8B_per_lane\": %.1f, \"blocks_per_sm_8B\": %d, \"l2_read_gbs_16B_per_lane\": %.1f, "
           "\"blocks_per_sm_16B\": %d, \"status\": \"%s\"}\n", prop.name, mb, best8, b8, best16, b16, cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}
