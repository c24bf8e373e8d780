// In which order does cluster launch control (clusterlaunchcontrol.try_cancel) hand out the CTAs of a grid?
// The forward kernel numbers its work items longest-first and relies on them being handed out in that order.
// Each CTA that starts records (first id = blockIdx.x), then keeps cancelling pending CTAs and records the id it
// got and the time, spending ~WORK_US microseconds per item. Output: ids in time order, per-SM sequences.
//   nvcc -arch=sm_100a -o clc_order clc_order.cu && ./clc_order [grid] [work_us] [smem_kb]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cstdint>

__device__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct Rec { unsigned long long t; int id; int sm; int first; int lat_cycles; };

__global__ void clc_kernel(Rec* recs, int* cursor, int work_us) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(16) unsigned int resp[4];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
    const unsigned resp_a = (unsigned)__cvta_generic_to_shared(resp);
    unsigned smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    int id = blockIdx.x, first = 1;
    unsigned phase = 0;
    while (true) {
        long long c0 = clock64();
        // "work"
        unsigned long long t0 = gtimer();
        while (gtimer() - t0 < (unsigned long long)work_us * 1000ull) {}
        // next
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" ::"r"(bar_a) : "memory");
        long long c1 = clock64();
        asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];" ::"r"(resp_a), "r"(bar_a) : "memory");
        unsigned ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar_a), "r"(phase) : "memory");
        }
        long long c2 = clock64();
        phase ^= 1;
        unsigned valid, x;
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b128 r;\n\tld.shared.b128 r, [%2];\n\t"
            "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p, r;\n\tselp.u32 %0, 1, 0, p;\n\tmov.u32 %1, 0;\n\t"
            "@p clusterlaunchcontrol.query_cancel.get_first_ctaid::x.b32.b128 %1, r;\n\t}"
            : "=r"(valid), "=r"(x) : "r"(resp_a) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        int slot = atomicAdd(cursor, 1);
        recs[slot] = Rec{gtimer(), id, (int)smid, first, (int)(c2 - c1)};
        (void)c0;
        if (!valid) break;
        id = (int)x;
        first = 0;
    }
}

int main(int argc, char** argv) {
    int grid = argc > 1 ? atoi(argv[1]) : 1024;
    int work_us = argc > 2 ? atoi(argv[2]) : 20;
    int smem_kb = argc > 3 ? atoi(argv[3]) : 200;
    Rec* d; int* cur;
    cudaMalloc(&d, sizeof(Rec) * (grid + 16));
    cudaMalloc(&cur, 4);
    cudaMemset(cur, 0, 4);
    cudaFuncSetAttribute(clc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
    clc_kernel<<<grid, 128, smem_kb * 1024>>>(d, cur, work_us);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    int n = 0;
    cudaMemcpy(&n, cur, 4, cudaMemcpyDeviceToHost);
    std::vector<Rec> h(n);
    cudaMemcpy(h.data(), d, sizeof(Rec) * n, cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end(), [](const Rec& a, const Rec& b) { return a.t < b.t; });
    int started = 0; long long lat = 0;
    for (auto& r : h) { started += r.first; lat += r.lat_cycles; }
    printf("grid %d, records %d, CTAs that really started %d, mean try_cancel latency %lld cycles\n", grid, n, started, n ? lat / n : 0);
    // ids in completion order: monotone?
    int inversions = 0, maxback = 0;
    std::vector<int> stolen;
    for (auto& r : h) if (!r.first) stolen.push_back(r.id);
    for (size_t i = 1; i < stolen.size(); ++i) if (stolen[i] < stolen[i - 1]) { ++inversions; maxback = std::max(maxback, stolen[i - 1] - stolen[i]); }
    printf("stolen ids: %zu, inversions in time order %d (max step back %d)\n", stolen.size(), inversions, maxback);
    printf("first 40 stolen ids in time order:");
    for (size_t i = 0; i < stolen.size() && i < 40; ++i) printf(" %d", stolen[i]);
    printf("\nlast 20:");
    for (size_t i = stolen.size() > 20 ? stolen.size() - 20 : 0; i < stolen.size(); ++i) printf(" %d", stolen[i]);
    printf("\nfirst-wave ids (min..max):");
    int mn = 1 << 30, mx = -1;
    for (auto& r : h) if (r.first) { mn = std::min(mn, r.id); mx = std::max(mx, r.id); }
    printf(" %d..%d\n", mn, mx);
    return 0;
}
