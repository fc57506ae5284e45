// What the box allows for concurrent pinned host->device uploads: the ceiling of the end-to-end (upload-bound) scaling curve.
//   nvcc -O3 -o tools/microbench/h2d_peak tools/microbench/h2d_peak.cu -lpthread && tools/microbench/h2d_peak [MB per GPU = 1024]
// For N = 1, 2, 4, 8 GPUs at once (one host thread per GPU), per-GPU and aggregate GB/s of cudaMemcpyAsync from page-locked
// memory allocated (a) wherever the calling thread's default policy puts it and (b) on the GPU's own NUMA node
// (set_mempolicy(MPOL_PREFERRED) around cudaHostAlloc — the same thing pb_host_alloc_near does).  One JSON line.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <atomic>
#include <cuda_runtime.h>
#include <sys/syscall.h>
#include <unistd.h>

static int numa_node_of_device(int dev) {
    char bus[64] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, dev) != cudaSuccess) return -1;
    for (char* p = bus; *p; p++) *p = (char)tolower(*p);
    std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}
static long set_preferred_node(int node) {   // node < 0: back to the default policy
    unsigned long mask[16] = {0};
    if (node >= 0) mask[node / 64] |= 1ul << (node % 64);
    return syscall(SYS_set_mempolicy, node >= 0 ? 1 /* MPOL_PREFERRED */ : 0 /* MPOL_DEFAULT */, node >= 0 ? mask : nullptr, node >= 0 ? 1024ul : 0ul);
}

struct Res { double gbs = 0; int node = -1; long policy_rc = 0; };

int main(int argc, char** argv) {
    const size_t mb = argc > 1 ? (size_t)atoll(argv[1]) : 1024;
    const size_t bytes = mb << 20;
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    printf("{\"mb_per_gpu\": %zu, \"gpus\": %d, \"runs\": [", mb, ndev);
    bool first = true;
    for (int near = 0; near <= 1; near++)
        for (int n : {1, 2, 4, 8}) {
            if (n > ndev) continue;
            std::vector<Res> res((size_t)n);
            std::atomic<int> ready{0};
            std::vector<std::thread> th;
            for (int g = 0; g < n; g++)
                th.emplace_back([&, g] {
                    cudaSetDevice(g);
                    Res& r = res[(size_t)g];
                    r.node = numa_node_of_device(g);
                    if (near) r.policy_rc = set_preferred_node(r.node);
                    void *h = nullptr, *d = nullptr;
                    cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
                    memset(h, 1, bytes);
                    if (near) set_preferred_node(-1);
                    cudaMalloc(&d, bytes);
                    cudaStream_t st; cudaStreamCreate(&st);
                    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                    cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st);
                    cudaStreamSynchronize(st);
                    ready++;
                    while (ready.load() < n) std::this_thread::yield();
                    cudaEventRecord(e0, st);
                    for (int i = 0; i < 4; i++) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st);
                    cudaEventRecord(e1, st);
                    cudaEventSynchronize(e1);
                    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
                    r.gbs = 4.0 * bytes / (ms * 1e-3) / 1e9;
                    cudaFree(d); cudaFreeHost(h);
                });
            for (auto& t : th) t.join();
            double sum = 0, mn = 1e30;
            for (auto& r : res) { sum += r.gbs; if (r.gbs < mn) mn = r.gbs; }
            printf("%s{\"n\": %d, \"pinned_on_gpu_numa_node\": %s, \"per_gpu_min_gbs\": %.1f, \"aggregate_gbs\": %.1f, \"nodes\": [", first ? "" : ", ", n,
                   near ? "true" : "false", mn, sum);
            for (int g = 0; g < n; g++) printf("%s%d", g ? "," : "", res[(size_t)g].node);
            printf("], \"set_mempolicy_rc\": %ld}", res[0].policy_rc);
            first = false;
        }
    printf("]}\n");
    return 0;
}
