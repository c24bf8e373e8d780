// Builds oracle/_ref/libref_cpu_attention.so: the REFERENCE's own CPU attention
// (`cpu_attention`, /root/reference/utils/sass/mma_swizzle/forward_kernel.cu:346-370) compiled from the
// reference source file where it lies.  TEST INFRASTRUCTURE ONLY (checker + CPU baseline).
//
// The reference file is a standalone sm_70 harness with its own main(); it is included unmodified
// with main renamed, so the only code that runs here is the host function cpu_attention -- its
// __global__ kernel is compiled (for sm_70, the only arch it accepts) but never launched.
// Nothing of the reference is copied into this repository: REF_SRC points into /root/reference.
#include <thread>
#include <vector>

#define main ref_harness_main
#include REF_SRC
#undef main

extern "C" {

// One head: q [M,D], k,v [N,D], out [M,D], fp32 row-major (the harness's layout).
void ref_cpu_attention(const float* q, const float* k, const float* v, float* out, int M, int N, int D,
                       float scale, int causal) {
    std::vector<float> Q(q, q + (size_t)M * D), K(k, k + (size_t)N * D), V(v, v + (size_t)N * D), O((size_t)M * D);
    cpu_attention(Q, K, V, O, M, N, D, scale, causal != 0);
    std::copy(O.begin(), O.end(), out);
}

// `heads` independent heads laid out back to back, spread over `threads` host threads.
void ref_cpu_attention_batched(const float* q, const float* k, const float* v, float* out, int heads, int M,
                               int N, int D, float scale, int causal, int threads) {
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([=] {
            for (int h = t; h < heads; h += threads)
                ref_cpu_attention(q + (size_t)h * M * D, k + (size_t)h * N * D, v + (size_t)h * N * D,
                                  out + (size_t)h * M * D, M, N, D, scale, causal);
        });
    }
    for (auto& th : pool) th.join();
}

}  // extern "C"
