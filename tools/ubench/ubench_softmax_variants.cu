// Softmax formulations under kernel-like conditions (bring-up tool): a 512-thread CTA with the forward's register split
// (setmaxnreg: softmax warps 192, the rest 48), warps 0-7 run the softmax tile loop back to back (two per scheduler, as
// the two stages do), warp 12 keeps the tensor pipe ~99 % busy with the forward's own GEMMs on the same TMEM
// (S = Q K^T SS-form, O += P V TS-form reading the P columns the softmax warps write). Reported: clocks per two tiles
// (one per stage) for each variant -- 2048 would be tensor-bound at head_dim 128.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench_softmax_variants ubench_softmax_variants.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../flash-attention-v100_b200/csrc/ptx_sm100.cuh"
#include "../../flash-attention-v100_b200/csrc/tmem_ldst_gen.cuh"
#include "../../flash-attention-v100_b200/csrc/umma_issue_gen.cuh"
using namespace fa;

struct V {
    int emu_period, emu_count;  // emu_count of every emu_period pairs on the FMA pipe
    int ld;                     // 0: 4 x32 + one wait; 1: two halves, max of the first overlaps the second's load
    int st;                     // 0: four x16 stores; 1: two x32 stores; 2: one x64 store
    int chains;                 // max chains: 4 or 8
    int wait34;                 // 1: wait::st after 3/4 and at the end (kernel); 0: only at the end
};

template <int EP, int EC, int LD, int ST, int CH, int W34, int PK = 0>
__device__ __forceinline__ void tile(uint32_t tS, uint32_t tP, float sl2, float& m_ref, float& row_sum) {
    float v[128];
    float mx[CH];
    if (LD == 0) {
        tmem_ld_4x32_wait(tS, reinterpret_cast<uint32_t*>(v));
    } else {
        tmem_ld_2x32_wait(tS, reinterpret_cast<uint32_t*>(v));
        tmem_ld_2x32_nowait(tS + 64, reinterpret_cast<uint32_t*>(v + 64));
    }
#pragma unroll
    for (int a = 0; a < CH; ++a) mx[a] = fmaxf(v[2 * a], v[2 * a + 1]);
#pragma unroll
    for (int c = 2 * CH; c < 64; c += 2 * CH) {
#pragma unroll
        for (int a = 0; a < CH; ++a) mx[a] = fmax3(mx[a], v[c + 2 * a], v[c + 2 * a + 1]);
    }
    if (LD == 1) tmem_wait_ld_x64(reinterpret_cast<uint32_t*>(v + 64));
#pragma unroll
    for (int c = 64; c < 128; c += 2 * CH) {
#pragma unroll
        for (int a = 0; a < CH; ++a) mx[a] = fmax3(mx[a], v[c + 2 * a], v[c + 2 * a + 1]);
    }
    float m = mx[0];
#pragma unroll
    for (int a = 1; a < CH; ++a) m = fmaxf(m, mx[a]);
    const float m_new = fmaxf(m_ref, m);
    float acc_scale = 1.0f;
    if ((m_ref - m_new) * sl2 < -8.0f) {
        acc_scale = ex2_approx((m_ref - m_new) * sl2);
        m_ref = m_new;
    }
    const float neg_m = -m_ref * sl2;
    float sum0 = 0.f, sum1 = 0.f;
    uint32_t pk[64];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
            float p0 = v[ch * 32 + c], p1 = v[ch * 32 + c + 1];
            if (PK == 2) {
                p0 = fmaf(p0, sl2, neg_m);
                p1 = fmaf(p1, sl2, neg_m);
            } else {
                fma2(p0, p1, sl2, sl2, neg_m, neg_m);
            }
            if (EC > 0 && ((c / 2) % EP) >= EP - EC) {
                ex2_emu2(p0, p1);
            } else {
                p0 = ex2_approx(p0);
                p1 = ex2_approx(p1);
            }
            if (PK == 2) {
                sum0 += p0;
                sum1 += p1;
            } else {
                add2(sum0, sum1, p0, p1);
            }
            if (PK == 1) {
                asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(pk[ch * 16 + c / 2]) : "r"(__float_as_uint(p0)), "r"(__float_as_uint(p1)));
            } else {
                pk[ch * 16 + c / 2] = pack2<true>(p0, p1);
            }
        }
        if (ST == 0) tmem_st_x16(tP + ch * 16, pk + ch * 16);
        if (ST == 1 && (ch & 1)) tmem_st_x32(tP + (ch - 1) * 16, pk + (ch - 1) * 16);
        if (ST == 2 && ch == 3) tmem_st_x64(tP, pk);
        if (W34 && ch == 2) tmem_wait_st();
    }
    tmem_wait_st();
    row_sum = row_sum * acc_scale + (sum0 + sum1);
}

template <int EP, int EC, int LD, int ST, int CH, int W34, int PK = 0>
__global__ void __launch_bounds__(512, 1) k(float* out, const float* in, int iters, long long* cyc, int hog) {
    extern __shared__ uint8_t dyn_smem[];
    __shared__ uint32_t tmem_ptr;
    __shared__ uint64_t bars[2];
    __shared__ int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        stop = 0;
        mbar_fence_init();
    }
    if (warp == 12) tmem_alloc<512>(smem_u32(&tmem_ptr));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    if (warp < 8) {
        reg_inc<192>();
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_off + (warp >> 2) * 128;
        const uint32_t tP = tS + 64;
        const float sl2 = in[1];
        float m_ref = in[2], row_sum = 0.f;
        {
            uint32_t z[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(in[(lane * 32 + i) % 977]);
            for (int c = 0; c < 4; ++c) tmem_st_x32(tS + c * 32, z);
            tmem_wait_st();
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) tile<EP, EC, LD, ST, CH, W34, PK>(tS, tP, sl2, m_ref, row_sum);
        const long long t1 = clock64();
        if (threadIdx.x == 0) cyc[0] = t1 - t0;
        out[threadIdx.x] = row_sum + m_ref;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(&stop) = 1;
    } else if (warp == 12) {
        reg_dec<48>();
        const uint32_t sb = (smem_u32(dyn_smem) + 1023u) & ~1023u;
        constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t q_lo = ((sb & 0x3FFFFu) >> 4) | (1u << 16), k_lo = (((sb + 32768) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t v_lo = (((sb + 65536) & 0x3FFFFu) >> 4) | ((uint32_t)(16384 >> 4) << 16);
        constexpr uint32_t idesc_qk = umma_idesc_f16(true, 128, 128, false, false), idesc_pv = umma_idesc_f16(true, 128, 128, false, true);
        const uint32_t bar_m = smem_u32(&bars[0]);
        int n = 0;
        while (hog != 0 && *reinterpret_cast<volatile int*>(&stop) == 0) {
            if (hog > 1) {  // duty cycle: idle `hog` clocks per 2048 clocks of MMAs
                const long long t = clock64();
                while (clock64() - t < hog) {}
            }
            tc_fence_after();
            umma_issue_qk_d128(tmem_base + 256, q_lo, k_lo, kDescHi, kDescHi, idesc_qk);
            umma_issue_pv_k0_8(tmem_base + 384, tmem_base + 64, v_lo, 0, kDescHi, idesc_pv, 1u);
            umma_issue_qk_d128(tmem_base + 256, q_lo, k_lo, kDescHi, kDescHi, idesc_qk);
            umma_issue_pv_k0_8(tmem_base + 384, tmem_base + 192, v_lo, 0, kDescHi, idesc_pv, 1u);
            umma_commit_elect(bar_m);
            if (n >= 1) mbar_wait(bar_m, (n - 1) & 1);
            ++n;
        }
        if (n > 0) mbar_wait(bar_m, (n - 1) & 1);
        if (lane == 0) cyc[1] = n;
    } else {
        reg_dec<48>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) tmem_dealloc<512>(tmem_base);
}

template <int EP, int EC, int LD, int ST, int CH, int W34, int PK = 0>
void run(const char* name, float* out, float* in, long long* cyc) {
    auto kern = k<EP, EC, LD, ST, CH, W34, PK>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 400;
    long long c2[2];
    printf("%-58s", name);
    const int hogs[3] = {0, 700, 1};  // tensor pipe idle / ~75 % busy / saturated
    for (int h = 0; h < 3; ++h) {
        double best = 1e30, busy = 0;
        for (int r = 0; r < 3; ++r) {
            kern<<<1, 512, 100 * 1024>>>(out, in, iters, cyc, hogs[h]);
            cudaDeviceSynchronize();
            cudaMemcpy(c2, cyc, 16, cudaMemcpyDeviceToHost);
            if ((double)c2[0] < best) {
                best = (double)c2[0];
                busy = 100.0 * c2[1] * 2048.0 / (double)c2[0];
            }
        }
        printf("  %7.1f (%3.0f %%)", best / iters, busy);
    }
    printf("\n");
}

int main() {
    float *out, *in;
    long long* cyc;
    cudaMalloc(&out, 4096 * 4);
    cudaMalloc(&in, 4096 * 4);
    cudaMalloc(&cyc, 16);
    float h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = -0.01f * (i % 977);
    h[1] = 0.127f;
    h[2] = 0.5f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("clocks per two tiles (tensor pipe busy %%) with the tensor pipe idle / ~75 %% busy / saturated\n");
    //   EP EC LD ST CH W34
    run<4, 1, 0, 0, 4, 1>("kernel: 1/4 emu, 4x32 load, x16 st, 4 chains, 3/4 wait", out, in, cyc);
    run<4, 0, 0, 0, 4, 1>("no emulation", out, in, cyc);
    run<8, 1, 0, 0, 4, 1>("1/8 emulated", out, in, cyc);
    run<8, 3, 0, 0, 4, 1>("3/8 emulated", out, in, cyc);
    run<4, 1, 1, 0, 4, 1>("split load (max of 1st half under 2nd load)", out, in, cyc);
    run<4, 1, 0, 1, 4, 1>("two x32 stores", out, in, cyc);
    run<4, 1, 0, 2, 4, 0>("one x64 store, one wait", out, in, cyc);
    run<4, 1, 0, 0, 8, 1>("8 max chains", out, in, cyc);
    run<4, 1, 0, 0, 4, 0>("no wait after 3/4", out, in, cyc);
    run<4, 1, 0, 0, 4, 1, 1>("pack by PRMT (truncation) instead of F2FP", out, in, cyc);
    run<4, 1, 0, 0, 4, 1, 2>("scalar FFMA / FADD for scale and row sum", out, in, cyc);
    run<4, 1, 0, 1, 4, 0, 1>("PRMT + x32 stores + one wait", out, in, cyc);
    run<8, 1, 0, 1, 4, 0, 1>("PRMT + x32 stores + one wait, 1/8 emulated", out, in, cyc);
    run<8, 3, 0, 1, 4, 0, 1>("PRMT + x32 stores + one wait, 3/8 emulated", out, in, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
