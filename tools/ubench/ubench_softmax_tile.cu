// How fast can one SM sub-partition run the forward's softmax when its warps never wait for S? (bring-up tool)
// Each warp loops over "tiles": tcgen05.ld of 128 fp32 columns, row max (3-input max chains), lazy-rescale test,
// exp2 (1/4 emulated on the FMA pipe), row sum, 16-bit pack, tcgen05.st of 64 columns -- the same statements as the
// softmax warps of fwd_sm100.cuh, without any mbarrier. Reported: cycles per tile per warp and per scheduler, for
// 1 / 2 / 3 / 4 warps per scheduler all running concurrently.
//   nvcc -O3 -std=c++17 -arch=sm_100a -o ubench_softmax_tile ubench_softmax_tile.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../flash-attention-v100_b200/csrc/ptx_sm100.cuh"
#include "../../flash-attention-v100_b200/csrc/tmem_ldst_gen.cuh"
#include "../../flash-attention-v100_b200/csrc/umma_issue_gen.cuh"
using namespace fa;

#ifndef MAXT
#define MAXT 288  // 256 threads: up to 2 warps per scheduler with no register cap (the kernel gives its softmax warps 192)
#endif
#ifndef EMU_PERIOD
#define EMU_PERIOD 4
#endif
#ifndef EMU_COUNT
#define EMU_COUNT 1
#endif

template <int MODE, int OVH = 0, int MMA = 0>  // MMA = 1: an extra warp keeps the tensor pipe busy with the forward's two GEMMs
                              // (S = Q K^T SS-form from shared memory, O += P V TS-form with P read from TMEM) on dummy data. OVH (mode 0 only): bit 0 four syncwarp + elected mbarrier arrivals per tile, bit 1 three waits on
                              // completed mbarriers, bit 2 tcgen05 fences, bit 3 stats store, bit 4 mask vote.  MODE 0: full tile; 1: no max; 2: no TMEM traffic (registers only); 3: max computed but the exponentials
                     // do not depend on it; 4: 2-input max tree; 5: max over packed halves (8 chains)
__global__ void __launch_bounds__(MAXT, 1) k(float* out, const float* in, int iters, long long* cyc) {
    extern __shared__ uint8_t dyn_smem[];
    __shared__ uint32_t tmem_ptr;
    __shared__ uint64_t bars[4];
    __shared__ float s_scale[512];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bars[0]), 1000000u);  // arrive target: never completes
        mbar_init(smem_u32(&bars[1]), 1);         // wait target: completed once below, waited for with the old parity
        mbar_init(smem_u32(&bars[2]), 1);         // MMA hog: commit barrier
        bars[3] = 0;                              // MMA hog: stop flag
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_ptr));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    // each warp of a scheduler gets its own 128 columns (4 warps per scheduler -> 512 columns)
    const uint32_t tS = tmem_base + lane_off + (warp >> 2) * 128;
    const uint32_t tP = tS + 64;
    const float sl2 = in[1];
    float m_ref = in[2];
    float row_sum = 0.f;
    {  // some finite data in TMEM
        uint32_t z[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(in[(lane * 32 + i) % 977]);
        for (int c = 0; c < 4; ++c) tmem_st_x32(tS + c * 32, z);
        tmem_wait_st();
    }
    __syncthreads();
    if (threadIdx.x == 0) mbar_arrive(smem_u32(&bars[1]));
    __syncthreads();
    const uint32_t bar_a = smem_u32(&bars[0]), bar_w = smem_u32(&bars[1]);
    auto arrive = [&]() {
        if (OVH & 4) tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a);
    };
    auto wait_done = [&]() {
        mbar_wait(bar_w, 0);
        if (OVH & 4) tc_fence_after();
    };
    if (MMA && warp == (int)(blockDim.x >> 5) - 1) {
        // tensor-pipe hog: per "iteration" two tiles' worth of GEMMs (2 x (QK^T + PV) = 2048 tensor clocks), paced by a
        // commit barrier so that the issue queue never runs dry nor overflows; stops when the softmax warps are done
        const uint32_t sb = (smem_u32(dyn_smem) + 1023u) & ~1023u;
        constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t q_lo = ((sb & 0x3FFFFu) >> 4) | (1u << 16), k_lo = (((sb + 32768) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t v_lo = (((sb + 65536) & 0x3FFFFu) >> 4) | ((uint32_t)(16384 >> 4) << 16);
        constexpr uint32_t idesc_qk = umma_idesc_f16(true, 128, 128, false, false), idesc_pv = umma_idesc_f16(true, 128, 128, false, true);
        const uint32_t bar_m = smem_u32(&bars[2]);
        volatile int* stop = reinterpret_cast<volatile int*>(&bars[3]);
        int n = 0;
        while (*stop == 0) {
            tc_fence_after();
            umma_issue_qk_d128(tmem_base + 256, q_lo, k_lo, kDescHi, kDescHi, idesc_qk);
            umma_issue_pv_k0_8(tmem_base + 384, tmem_base + 64, v_lo, 0, kDescHi, idesc_pv, 1u);
            umma_issue_qk_d128(tmem_base + 256, q_lo, k_lo, kDescHi, kDescHi, idesc_qk);
            umma_issue_pv_k0_8(tmem_base + 384, tmem_base + 192, v_lo, 0, kDescHi, idesc_pv, 1u);
            umma_commit_elect(bar_m);
            if (n >= 1) mbar_wait(bar_m, (n - 1) & 1);  // at most two iterations in flight
            ++n;
        }
        mbar_wait(bar_m, (n - 1) & 1);
        if (lane == 0) cyc[1] = n;
    } else {
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float v[128];
        if (OVH & 2) wait_done();  // s_full
        if (OVH & 16) {
            const bool need = (it & 1023) == 1023 && lane == 5;
            if (__any_sync(0xffffffffu, need)) row_sum += 1.0f;
        }
        if (MODE != 2) {
            tmem_ld_4x32_wait(tS, reinterpret_cast<uint32_t*>(v));
            if (OVH & 1) arrive();  // sx_free
        } else {
#pragma unroll
            for (int i = 0; i < 128; ++i) v[i] = in[i] + (float)it;
        }
        float acc_scale = 1.0f;
        float m_side = 0.f;
        if (MODE == 4) {
            float mx[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) mx[a] = fmaxf(v[2 * a], v[2 * a + 1]);
#pragma unroll
            for (int c = 16; c < 128; c += 16) {
#pragma unroll
                for (int a = 0; a < 8; ++a) mx[a] = fmaxf(mx[a], fmaxf(v[c + 2 * a], v[c + 2 * a + 1]));
            }
            const float m_new = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
            const float d = (m_ref - fmaxf(m_new, m_ref)) * sl2;
            if (d < -8.0f) {
                acc_scale = ex2_approx(d);
                m_ref = m_new;
            }
        } else if (MODE == 5) {
            float mx[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) mx[a] = fmaxf(v[2 * a], v[2 * a + 1]);
#pragma unroll
            for (int c = 16; c < 128; c += 16) {
#pragma unroll
                for (int a = 0; a < 8; ++a) mx[a] = fmax3(mx[a], v[c + 2 * a], v[c + 2 * a + 1]);
            }
            const float m_new = fmaxf(m_ref, fmax3(fmax3(mx[0], mx[1], mx[2]), fmax3(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7])));
            const float d = (m_ref - m_new) * sl2;
            if (d < -8.0f) {
                acc_scale = ex2_approx(d);
                m_ref = m_new;
            }
        } else if (MODE != 1) {
            float mx[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) mx[a] = fmaxf(v[2 * a], v[2 * a + 1]);
#pragma unroll
            for (int c = 8; c < 128; c += 8) {
#pragma unroll
                for (int a = 0; a < 4; ++a) mx[a] = fmax3(mx[a], v[c + 2 * a], v[c + 2 * a + 1]);
            }
            const float m_new = fmaxf(m_ref, fmax3(fmaxf(mx[0], mx[1]), mx[2], mx[3]));
            const float d = (m_ref - m_new) * sl2;
            if (MODE == 3) {
                m_side = m_new;
            } else if (d < -8.0f) {
                acc_scale = ex2_approx(d);
                m_ref = m_new;
            }
        }
        if (OVH & 8) s_scale[threadIdx.x & 511] = acc_scale;
        if (OVH & 1) arrive();  // stats
        const float neg_m = -m_ref * sl2;
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                float p0 = v[ch * 32 + c], p1 = v[ch * 32 + c + 1];
                fma2(p0, p1, sl2, sl2, neg_m, neg_m);
                if (EMU_COUNT > 0 && ((c / 2) % EMU_PERIOD) >= EMU_PERIOD - EMU_COUNT) {
                    ex2_emu2(p0, p1);
                } else {
                    p0 = ex2_approx(p0);
                    p1 = ex2_approx(p1);
                }
                add2(sum0, sum1, p0, p1);
                pk[c / 2] = pack2<true>(p0, p1);
            }
            if (MODE != 2) {
                if (ch == 0 && (OVH & 2)) wait_done();  // p_free
                tmem_st_x16(tP + ch * 16, pk);
                if (ch >= 2) {
                    tmem_wait_st();
                    if (OVH & 1) arrive();  // p_full (3/4), p_last
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) row_sum += __uint_as_float(pk[i] & 0x3fffffffu);
            }
        }
        row_sum = row_sum * acc_scale + (sum0 + sum1) + m_side;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = row_sum + m_ref;
    if (MMA) {
        asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x - 32) : "memory");  // all softmax warps are done
        if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(&bars[3]) = 1;
    }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

template <int MODE, int OVH = 0, int MMA = 0>
void run(const char* name, float* out, float* in, long long* cyc) {
    printf("%-40s", name);
    for (int wps = 1; wps <= 2; ++wps) {
        const int iters = 200;
        cudaFuncSetAttribute(k<MODE, OVH, MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        k<MODE, OVH, MMA><<<1, wps * 128 + (MMA ? 32 : 0), 100 * 1024>>>(out, in, iters, cyc);
        cudaDeviceSynchronize();
        k<MODE, OVH, MMA><<<1, wps * 128 + (MMA ? 32 : 0), 100 * 1024>>>(out, in, iters, cyc);
        cudaDeviceSynchronize();
        long long c, c2[2];
        cudaMemcpy(c2, cyc, 16, cudaMemcpyDeviceToHost);
        c = c2[0];
        if (MMA) printf(" [tensor busy %4.1f%%]", 100.0 * c2[1] * 2048.0 / (double)c);
        printf("  w/s %d: %7.1f per warp-tile, %7.1f per tile per scheduler |", wps, (double)c / iters, (double)c / iters / wps);
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
    printf("emulated exponentials: %d of every %d pairs\n", EMU_COUNT, EMU_PERIOD);
    run<0>("full tile (ld, max, exp, sum, pack, st)", out, in, cyc);
    run<1>("no row max", out, in, cyc);
    run<2>("registers only (no TMEM traffic)", out, in, cyc);
    run<0, 1>("full tile + 4 arrivals", out, in, cyc);
    run<0, 2>("full tile + 2 waits on done barriers", out, in, cyc);
    run<0, 3>("full tile + arrivals + waits", out, in, cyc);
    run<0, 7>("full tile + arrivals + waits + tc fences", out, in, cyc);
    run<0, 15>("... + stats store", out, in, cyc);
    run<0, 31>("... + mask vote (all kernel hand-shakes)", out, in, cyc);
    run<0, 0, 1>("full tile + tensor pipe busy (QK^T SS, PV TS)", out, in, cyc);
    run<0, 31, 1>("all hand-shakes + tensor pipe busy", out, in, cyc);
    run<3>("max computed, exps independent of it", out, in, cyc);
    run<4>("2-input max, 8 chains", out, in, cyc);
    run<5>("3-input max, 8 chains", out, in, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
