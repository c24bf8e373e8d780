// Fused attention backward for sm_100a:  dQ, dK, dV from dO, Q, K, V, LSE and delta = rowsum(dO * O).
//
// What it replaces: the reference's backward kernels
//   flash_attention_backward_kernel          (reference kernel/fused_mha_backward.cu:26-505)
//   flash_attention_backward_varlen_kernel   (reference kernel/fused_mha_backward_varlen.cu)
// and their building blocks (row dot include/product.h:9-96, softmax gradient include/softmax.h:205-406,
// gradient GEMMs include/mat_mul.h:166-234). The reference runs two phases in one grid -- phase 1: one CTA per
// query block accumulates dQ over the KV tiles; phase 2: one CTA per KV block (per KV head, looping over the GQA
// group) accumulates dK and dV over the query tiles -- each recomputing S and dP, so no atomics are needed and
// the result is deterministic. That two-pass structure is kept (it is the right one for an exact drop-in:
// deterministic, no fp32 dQ workspace), the schedule is Blackwell-native:
//
//   * one kernel template, two instantiations: KV_STAT = false is the dQ pass, KV_STAT = true the dK/dV pass.
//     The "stationary" block (128 rows of Q+dO, or of K+V) sits in smem for the whole CTA, the other side
//     streams through a TMA ring in 128-row tiles.
//   * per streamed tile:  T1 = A1 B1^T  (S or S^T),  T2 = A2 B2^T  (dP or dP^T)   -- tcgen05.mma SS form
//                         P  = exp2(T1 * scale * log2e - lse * log2e),  dS = P * (dP - delta)
//                         out1 += dS B1   (dQ += dS K   /  dK += dS^T Q)          -- TS form, A read from TMEM
//                         out2 += P  B2   (             /  dV += P^T dO)
//     All accumulators live in TMEM: T1 [0,128) T2 [128,256) out1 [256,256+D) out2 [256+D,256+2D).
//     P (16-bit) is written back over T1, dS over T2, each warpgroup into the columns it read itself.
//   * warps 0-7: element-wise stage; thread == TMEM lane (a row of the stationary block); warpgroup 0 takes tile
//     columns [0,64), warpgroup 1 [64,128). Per-row statistics (dQ pass) are registers, per-column statistics
//     (dK/dV pass) come from a double-buffered smem table filled by the same warps.
//   * warp 8: TMA producer.  warp 9: tcgen05.mma issuer.  mbarriers order everything.
//
// The softmax scale is applied in the epilogue (dQ, dK *= scale), and so is dropout's 1/(1-p) on dV.
//
// head_dim 256 ("split-D"): two 128 x 256 stationary tiles fill 128 KB of shared memory and two 256-column
// accumulators would fill TMEM, so (a) the streamed side moves in 64-row tiles (T1, T2 are 128 x 64) through a
// 3-slot ring, and (b) each CTA owns one 128-column half of the outputs: grid.x = 2 x blocks, CTA 2b+h
// accumulates out[:, 128h .. 128h+128) of block b. T1 and T2 are contracted over all 256 dims by both CTAs.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "fwd_sm100.cuh"

#ifndef FA_BWD_EMU_PERIOD
#define FA_BWD_EMU_PERIOD 4  // of every FA_BWD_EMU_PERIOD pairs of exponentials ...
#endif
#ifndef FA_BWD_EMU_COUNT
#define FA_BWD_EMU_COUNT 1   // ... this many are evaluated on the FMA pipes instead of the MUFU
#endif

namespace fa {

struct alignas(64) BwdKernelParams {
    CUtensorMap tm_q;   // 4-D (head_dim, heads, rows, batch), box (64, 1, 128, 1), 128B swizzle
    CUtensorMap tm_k;
    CUtensorMap tm_v;
    CUtensorMap tm_do;
    CUtensorMap tm_q_stat;   // head_dim 256, dQ pass only: 128-row boxes for the stationary Q / dO tiles
    CUtensorMap tm_do_stat;  // (tm_q / tm_do / tm_k / tm_v then carry the 64-row boxes of the streamed side)
    void* dq;
    void* dk;
    void* dv;
    int64_t dq_stride_b, dq_stride_s, dq_stride_h;  // elements
    int64_t dk_stride_b, dk_stride_s, dk_stride_h;
    int64_t dv_stride_b, dv_stride_s, dv_stride_h;
    const float* lse;    // forward LSE (natural log); [B,H,Sq] or var-len [H,T]
    const float* delta;  // rowsum(dO * O), same layout as lse
    int64_t lse_stride_b, lse_stride_h;
    const int* cu_seqlens_q;
    const int* cu_seqlens_k;
    const float* alibi;
    int64_t alibi_stride_b;
    int seqlen_q;  // dense Sq / var-len max_seqlen_q
    int seqlen_k;  // dense Sk / var-len max_seqlen_k (also the dropout row length)
    int num_heads;
    int heads_per_kv;
    int head_dim;      // real head dim (multiple of 8, <= D): TMA zero-fills columns [head_dim, D) of every tile
    float scale;
    float scale_log2;
    float softcap;
    int window_left;   // -1 = unbounded
    int window_right;  // -1 = unbounded; causal is window_right = 0
    float rp_dropout;
    uint32_t drop_thr;
    uint64_t drop_seed;
    uint64_t drop_offset;
    int num_blocks;  // stationary blocks along the (longest) sequence
    int reverse;     // launch the last block first (dQ pass with a causal / right-window mask)
    // 1-D launch order: CTA w = (block rank, (head, batch) pair[, D-half]), numbered section by section -- a section is
    // `section_bh` (head, batch) pairs whose streamed tensors fit L2 together; inside a section every pair's
    // first-ranked (longest) block comes first. The hardware dispatches CTAs in this order, so the long blocks of a
    // section start first (with GQA the dK/dV blocks differ by up to 128 tiles) and the section's streams stay in L2.
    int num_bh;       // (grid heads) x batch; grid heads = KV heads in the dK/dV pass, query heads in the dQ pass
    int grid_heads;
    int section_bh;
};

template <int D>
struct BwdConfig {
    static constexpr bool kSplitD = (D == 256);           // see "split-D" in the header comment
    static constexpr int kBTS = kSplitD ? 64 : 128;       // rows of a streamed tile (= columns of T1, T2)
    static constexpr int kOW = kSplitD ? 128 : D;         // columns of one accumulator in this CTA
    static constexpr int kTileBytes = 128 * D * 2;        // stationary tile
    static constexpr int kHalfBytes = 128 * 128;          // one 64-column swizzle block of a stationary tile
    static constexpr int kStreamBytes = kBTS * D * 2;     // streamed tile
    static constexpr int kStreamBlk = kBTS * 128;         // one 64-column swizzle block of a streamed tile
    // Streamed tiles in flight (B1, B2 alternate). B1(t) is read by T1(t) early in step t-1 and by the out1 GEMM at
    // the very end of step t, so full overlap needs B1 of three steps and B2 of two resident at once: 5 slots.
    static constexpr int kRing = (D == 256) ? 3 : (D == 128) ? 5 : 10;
    static constexpr int kSmemStat = 2 * kTileBytes;
    static constexpr int kSmemRing = kRing * kStreamBytes;
    static constexpr int kNumBars = 2 + 2 * kRing + 6;
    static constexpr int kOffBars = kSmemStat + kSmemRing;
    static constexpr int kOffTmemPtr = kOffBars + 8 * kNumBars;
    static constexpr int kOffStats = (kOffTmemPtr + 16 + 15) & ~15;  // float [2 buffers][2 kinds][kBTS]
    static constexpr int kSmemUsed = kOffStats + 2 * 2 * kBTS * 4;
    // No slack for manual alignment: the dynamic shared window is declared __align__(1024) and the kernel traps
    // if the runtime ever hands it a base that is not (the 128B swizzle needs 1024-byte aligned tiles).
    static constexpr int kSmemBytes = kSmemUsed;
    static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA can have");
    static constexpr int kTmemT1 = 0, kTmemT2 = 128, kTmemOut1 = 256, kTmemOut2 = 256 + kOW;
};

FA_DEVICE uint32_t pack_half2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
FA_DEVICE float2 unpack_half2(uint32_t v) {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

// LSE as the kernel uses it: -lse * log2(e); rows without any visible key (sentinel / -inf) get a finite value
// (all of their scores are masked, so P is 0 whatever it is).
FA_DEVICE float neg_lse_log2(float lse) {
    return (lse > -1e29f && lse < 1e30f) ? -lse * kLog2e : 0.f;
}

// delta[row] = sum_d O[row, d] * dO[row, d] in fp32 (reference include/product.h:9-96); this is also the
// `softmax_d` tensor the operator returns. HBM-bound (reads O and dO once): a row is shared by LPR = head_dim/8
// lanes (rounded up to a power of two) with one 16-byte load each, so a warp covers 32/LPR rows per pass, and
// every warp keeps kDotUnroll passes of loads in flight.
constexpr int kDotUnroll = 4;
template <bool BF16>
__global__ void fa_bwd_dot_kernel(const uint16_t* __restrict__ o, const uint16_t* __restrict__ dout,
                                  float* __restrict__ delta, int head_dim, int64_t rows_total, int seqlen_q,
                                  int heads, int64_t o_sb, int64_t o_ss, int64_t o_sh, int64_t do_sb, int64_t do_ss,
                                  int64_t do_sh, int64_t d_sb, int64_t d_sh, int lpr) {
    const int lane = threadIdx.x & 31;
    const int rows_per_pass = 32 / lpr;
    const int sub = lane / lpr, chunk = lane - sub * lpr;
    const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t row0 = warp_id * (rows_per_pass * kDotUnroll) + sub;
    uint4 a[kDotUnroll], g[kDotUnroll];
    int64_t out_idx[kDotUnroll];
    // row enumerates (b, s, h) with h fastest: neighbouring rows are neighbouring heads of one token. One 32-bit
    // decomposition per thread, then carries (64-bit divisions per row made this kernel ALU-bound).
    const uint32_t r32 = (uint32_t)(row0 < rows_total ? row0 : 0);
    int h = (int)(r32 % (uint32_t)heads);
    const uint32_t bs = r32 / (uint32_t)heads;
    int sq = (int)(bs % (uint32_t)seqlen_q);
    int64_t b = bs / (uint32_t)seqlen_q;
#pragma unroll
    for (int u = 0; u < kDotUnroll; ++u) {
        const int64_t row = row0 + (int64_t)u * rows_per_pass;
        a[u] = make_uint4(0, 0, 0, 0);
        g[u] = make_uint4(0, 0, 0, 0);
        out_idx[u] = -1;
        if (row < rows_total) {
            out_idx[u] = b * d_sb + (int64_t)h * d_sh + sq;
            if (chunk * 8 < head_dim) {
                a[u] = *reinterpret_cast<const uint4*>(o + b * o_sb + (int64_t)sq * o_ss + (int64_t)h * o_sh + chunk * 8);
                g[u] = *reinterpret_cast<const uint4*>(dout + b * do_sb + (int64_t)sq * do_ss + (int64_t)h * do_sh + chunk * 8);
            }
        }
        h += rows_per_pass;
        while (h >= heads) {
            h -= heads;
            if (++sq == seqlen_q) {
                sq = 0;
                ++b;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kDotUnroll; ++u) {
        const uint32_t aw[4] = {a[u].x, a[u].y, a[u].z, a[u].w}, gw[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float2 x, y;
            if constexpr (BF16) {
                x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[e]));
                y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[e]));
            } else {
                x = __half22float2(*reinterpret_cast<const __half2*>(&aw[e]));
                y = __half22float2(*reinterpret_cast<const __half2*>(&gw[e]));
            }
            acc = fmaf(x.x, y.x, acc);
            acc = fmaf(x.y, y.y, acc);
        }
        for (int off = lpr >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (chunk == 0 && out_idx[u] >= 0) delta[out_idx[u]] = acc;
    }
}

template <int D, bool BF16, bool FEAT, bool KV_STAT, bool DROPOUT>
__global__ void __launch_bounds__(384, 1)
fa_bwd_sm100_kernel(const __grid_constant__ BwdKernelParams p) {
    using Cfg = BwdConfig<D>;
    constexpr int BT = 128;           // rows of the stationary block
    constexpr int BTS = Cfg::kBTS;    // rows of a streamed tile
    constexpr int CW = BTS / 2;       // tile columns one element-wise warpgroup handles
    constexpr bool SPLIT = Cfg::kSplitD;
    constexpr int RING = Cfg::kRing;

    // ------------------------------------------------------------------ geometry (uniform over the CTA)
    int w = (int)blockIdx.x;
    const int dhalf = SPLIT ? w & 1 : 0;  // split-D: which 128-column half of the outputs is mine
    if constexpr (SPLIT) w >>= 1;
    const int per_section = p.section_bh * p.num_blocks;
    const int sec = w / per_section;
    const int w_in = w - sec * per_section;
    const int sec_n = min(p.section_bh, p.num_bh - sec * p.section_bh);
    const int rank = w_in / sec_n;
    const int bh = sec * p.section_bh + (w_in - rank * sec_n);
    const int hh = bh % p.grid_heads;  // KV head (dK/dV pass) or query head (dQ pass)
    const int batch = bh / p.grid_heads;
    int q_off = 0, q_b = batch, seqlen_q = p.seqlen_q, k_off = 0, k_b = batch, seqlen_k = p.seqlen_k;
    if (p.cu_seqlens_q) {
        q_off = p.cu_seqlens_q[batch];
        seqlen_q = p.cu_seqlens_q[batch + 1] - q_off;
        q_b = 0;
        k_off = p.cu_seqlens_k[batch];
        seqlen_k = p.cu_seqlens_k[batch + 1] - k_off;
        k_b = 0;
    }
    const int o_b = p.cu_seqlens_q ? 0 : batch;
    const int G = p.heads_per_kv;
    const int off = seqlen_k - seqlen_q;
    const int blk = p.reverse ? p.num_blocks - 1 - rank : rank;
    const int x0 = blk * BT;  // first row of the stationary block (query position or key position)
    if (x0 >= (KV_STAT ? seqlen_k : seqlen_q)) return;
    const int kv_head = KV_STAT ? hh : hh / G;
    const int head0 = KV_STAT ? kv_head * G : hh;  // first (dK/dV pass) or only (dQ pass) query head

    // streamed tiles [t_lo, t_hi) along the other sequence, visible to at least one row of this block
    int t_lo = 0, t_hi;
    if constexpr (KV_STAT) {
        t_hi = (seqlen_q + BTS - 1) / BTS;
        const int j_last = min(x0 + BT, seqlen_k) - 1;
        if (p.window_right >= 0) t_lo = max(0, x0 - off - p.window_right) / BTS;
        if (p.window_left >= 0) {
            const int i_max = j_last - off + p.window_left;
            t_hi = min(t_hi, i_max < 0 ? 0 : i_max / BTS + 1);
        }
    } else {
        t_hi = (seqlen_k + BTS - 1) / BTS;
        const int i_last = min(x0 + BT, seqlen_q) - 1;
        if (p.window_right >= 0) {
            const int j_max = i_last + off + p.window_right;
            t_hi = min(t_hi, j_max < 0 ? 0 : j_max / BTS + 1);
        }
        if (p.window_left >= 0) t_lo = max(0, x0 + off - p.window_left) / BTS;
    }
    const int nt = max(t_hi - t_lo, 0);         // tiles per head
    const int n_tiles = KV_STAT ? nt * G : nt;  // the dK/dV pass streams the whole GQA group

    extern __shared__ __align__(1024) uint8_t smem_bwd_raw[];
    const uint32_t sbase = smem_u32(smem_bwd_raw);
    if ((sbase & 1023u) != 0u) __trap();  // see BwdConfig::kSmemBytes
    uint8_t* sgen = smem_bwd_raw;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const uint32_t sStat = sbase;                   // A1, A2
    const uint32_t sRing = sbase + Cfg::kSmemStat;  // B tiles
    const uint32_t bars = sbase + Cfg::kOffBars;
    auto bar_a_full = [&](int s) { return bars + 8 * s; };
    auto bar_ring_full = [&](int i) { return bars + 8 * (2 + i); };
    auto bar_ring_empty = [&](int i) { return bars + 8 * (2 + RING + i); };
    constexpr int kB0 = 2 + 2 * RING;
    const uint32_t bar_t1_full = bars + 8 * (kB0 + 0);   // MMA -> element-wise: T1 ready
    // MMA -> element-wise: T2 ready. The dQ pass has the out2 columns of TMEM to spare and keeps two T2 buffers
    // (tile t uses buffer t & 1), so dP of the next tile is computed while this tile's dS is still being made:
    // one barrier per buffer, because a barrier must never run two phases ahead of its waiter.
    // Only where the ring is deep enough (head_dim <= 64): the early dP GEMM needs B2(t+1) resident while B1(t), B2(t)
    // are still held, and with 5 slots (head_dim 128) waiting for it delays out1(t) and the slot releases behind it
    // (measured: -5 % at head_dim 128, +18 % at head_dim 64).
    constexpr bool T2DB = !KV_STAT && RING >= 6;
    auto bar_t2_full = [&](int t) { return bars + 8 * (T2DB && (t & 1) ? kB0 + 5 : kB0 + 1); };
    auto t2_parity = [&](int t) -> uint32_t { return T2DB ? (t >> 1) & 1 : t & 1; };
    auto t2_col = [&](int t) { return T2DB && (t & 1) ? Cfg::kTmemOut2 : Cfg::kTmemT2; };
    const uint32_t bar_p_ready = bars + 8 * (kB0 + 2);   // element-wise -> MMA: T1 consumed (and P written)
    const uint32_t bar_ds_ready = bars + 8 * (kB0 + 3);  // element-wise -> MMA: dS written over T2
    const uint32_t bar_out_full = bars + 8 * (kB0 + 4);  // MMA -> epilogue: accumulators final
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(sgen + Cfg::kOffTmemPtr);
    float* sTab = reinterpret_cast<float*>(sgen + Cfg::kOffStats);

    if (warp == 8 && lane == 0) {
        mbar_init(bar_a_full(0), 1);
        mbar_init(bar_a_full(1), 1);
        for (int i = 0; i < RING; ++i) {
            mbar_init(bar_ring_full(i), 1);
            mbar_init(bar_ring_empty(i), 1);
        }
        mbar_init(bar_t1_full, 1);
        mbar_init(bars + 8 * (kB0 + 1), 1);
        mbar_init(bars + 8 * (kB0 + 5), 1);
        mbar_init(bar_p_ready, 8);
        mbar_init(bar_ds_ready, 8);
        mbar_init(bar_out_full, 1);
        mbar_fence_init();
        tma_prefetch_desc(&p.tm_q);
        tma_prefetch_desc(&p.tm_k);
        tma_prefetch_desc(&p.tm_v);
        tma_prefetch_desc(&p.tm_do);
    }
    if (warp == 9) tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 8) {
        // ============================================================ TMA producer
        reg_dec<48>();
        if (n_tiles > 0) {
            auto load_tile = [&](const CUtensorMap* tm, uint32_t dst, uint32_t bar, int h, int row, int b) {
                mbar_arrive_expect_tx(bar, Cfg::kTileBytes);
#pragma unroll
                for (int c = 0; c < D / 64; ++c) tma_load_4d(dst + c * Cfg::kHalfBytes, tm, bar, c * 64, h, row, b);
            };
            auto load_stream = [&](const CUtensorMap* tm, uint32_t dst, uint32_t bar, int h, int row, int b) {
                mbar_arrive_expect_tx(bar, Cfg::kStreamBytes);  // the streamed maps' boxes have BTS rows
#pragma unroll
                for (int c = 0; c < D / 64; ++c) tma_load_4d(dst + c * Cfg::kStreamBlk, tm, bar, c * 64, h, row, b);
            };
            int ring = 0;
            auto produce = [&](const CUtensorMap* tm, int h, int row, int b) {
                const int slot = ring % RING;
                mbar_wait(bar_ring_empty(slot), ((ring / RING) & 1) ^ 1);
                if (lane == 0) load_stream(tm, sRing + slot * Cfg::kStreamBytes, bar_ring_full(slot), h, row, b);
                ++ring;
            };
            auto stream_tile = [&](int t, bool second) {
                const int g = KV_STAT ? t / nt : 0;
                const int ti = t_lo + (KV_STAT ? t - g * nt : t);
                if constexpr (KV_STAT) produce(second ? &p.tm_do : &p.tm_q, head0 + g, q_off + ti * BTS, q_b);
                else produce(second ? &p.tm_v : &p.tm_k, kv_head, k_off + ti * BTS, k_b);
            };
            if (lane == 0) {
                if constexpr (KV_STAT) load_tile(&p.tm_k, sStat, bar_a_full(0), kv_head, k_off + x0, k_b);
                else load_tile(SPLIT ? &p.tm_q_stat : &p.tm_q, sStat, bar_a_full(0), head0, q_off + x0, q_b);
            }
            stream_tile(0, false);
            if (lane == 0) {
                if constexpr (KV_STAT) load_tile(&p.tm_v, sStat + Cfg::kTileBytes, bar_a_full(1), kv_head, k_off + x0, k_b);
                else load_tile(SPLIT ? &p.tm_do_stat : &p.tm_do, sStat + Cfg::kTileBytes, bar_a_full(1), head0, q_off + x0, q_b);
            }
            stream_tile(0, true);
            for (int t = 1; t < n_tiles; ++t) {
                stream_tile(t, false);
                stream_tile(t, true);
            }
        }
    } else if (warp == 9) {
        // ============================================================ MMA issuer
        reg_dec<48>();
        if (n_tiles > 0) {
            constexpr uint32_t idesc_t = umma_idesc_f16(BF16, BT, BTS, false, false);
            constexpr uint32_t idesc_o = umma_idesc_f16(BF16, BT, Cfg::kOW, false, true);
            const uint32_t tT1 = tmem_base + Cfg::kTmemT1, tT2 = tmem_base + Cfg::kTmemT2;
            const uint32_t tO1 = tmem_base + Cfg::kTmemOut1, tO2 = tmem_base + Cfg::kTmemOut2;
            constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
            constexpr uint32_t kLoKmajor = 1u << 16;
            constexpr uint32_t kLoMn = (uint32_t)(Cfg::kStreamBlk >> 4) << 16;
            auto lo_addr = [](uint32_t saddr) { return (saddr & 0x3FFFFu) >> 4; };
            auto slot_addr = [&](int r) { return sRing + (r % RING) * Cfg::kStreamBytes; };
            // second GEMMs: B = the streamed tile as an MN-major operand; split-D takes this CTA's 128 columns of it
            auto mn_lo = [&](int r) { return lo_addr(slot_addr(r) + (SPLIT ? dhalf * 2 * Cfg::kStreamBlk : 0)) | kLoMn; };
            auto issue_o = [&](uint32_t d_tmem, uint32_t a_tmem, int r, uint32_t acc) {
                if constexpr (SPLIT) umma_issue_ts_split32(d_tmem, a_tmem, mn_lo(r), 0, kDescHi, idesc_o, acc);
                else umma_issue_ts_split(d_tmem, a_tmem, mn_lo(r), 0, kDescHi, idesc_o, acc);
            };
            auto wait_full = [&](int r) { mbar_wait(bar_ring_full(r % RING), (r / RING) & 1); };
            auto issue_t = [&](uint32_t d_tmem, int a_idx, int r) {  // d = A[a_idx] * B(ring r)^T
                const uint32_t a_lo = lo_addr(sStat + a_idx * Cfg::kTileBytes) | kLoKmajor;
                const uint32_t b_lo = lo_addr(slot_addr(r)) | kLoKmajor;
                if constexpr (D == 256) umma_issue_t_d256_n64(d_tmem, a_lo, b_lo, kDescHi, kDescHi, idesc_t);
                else if constexpr (D == 128) umma_issue_qk_d128(d_tmem, a_lo, b_lo, kDescHi, kDescHi, idesc_t);
                else umma_issue_qk_d64(d_tmem, a_lo, b_lo, kDescHi, kDescHi, idesc_t);
            };
            mbar_wait(bar_a_full(0), 0);
            wait_full(0);
            tc_fence_after();
            issue_t(tT1, 0, 0);
            umma_commit_elect(bar_t1_full);
            mbar_wait(bar_a_full(1), 0);
            wait_full(1);
            tc_fence_after();
            issue_t(tT2, 1, 1);
            umma_commit_elect(bar_t2_full(0));
            for (int t = 0; t < n_tiles; ++t) {
                const uint32_t ph = t & 1;
                mbar_wait(bar_p_ready, ph);  // T1(t) is in registers; in the dK/dV pass P(t) sits in T1's columns
                tc_fence_after();
                if constexpr (KV_STAT)
                    issue_o(tO2, tT1, 2 * t + 1, t > 0 ? 1u : 0u);
                if (t + 1 < n_tiles) {  // executes after the P read above (tcgen05 ops run in issue order)
                    wait_full(2 * t + 2);
                    tc_fence_after();
                    issue_t(tT1, 0, 2 * t + 2);
                    umma_commit_elect(bar_t1_full);
                    if constexpr (T2DB) {  // dP(t+1) goes into the other T2 buffer right away
                        wait_full(2 * t + 3);
                        tc_fence_after();
                        issue_t(tmem_base + t2_col(t + 1), 1, 2 * t + 3);
                        umma_commit_elect(bar_t2_full(t + 1));
                    }
                }
                mbar_wait(bar_ds_ready, ph);
                tc_fence_after();
                issue_o(tO1, tmem_base + t2_col(t), 2 * t, t > 0 ? 1u : 0u);
                umma_commit_elect(bar_ring_empty((2 * t) % RING));
                umma_commit_elect(bar_ring_empty((2 * t + 1) % RING));
                if (!T2DB && t + 1 < n_tiles) {
                    wait_full(2 * t + 3);
                    tc_fence_after();
                    issue_t(tT2, 1, 2 * t + 3);
                    umma_commit_elect(bar_t2_full(t + 1));
                }
            }
            umma_commit_elect(bar_out_full);
        }
    } else if (warp < 8) {
        // ============================================================ element-wise stage + epilogue
        reg_inc<216>();
        const int wg = warp >> 2;                  // column half of every tile
        const int r = (warp & 3) * 32 + lane;      // row of the stationary block == TMEM lane
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tT1 = tmem_base + lane_off + Cfg::kTmemT1 + wg * CW;
        const uint32_t tT2_0 = tmem_base + lane_off + wg * CW;  // + t2_col(t)
        const int x = x0 + r;  // this thread's query position (dQ pass) or key position (dK/dV pass)
        const float sl2 = FEAT ? 1.0f : p.scale_log2;

        // visible range [lo, lo + width) of the streamed index for this row
        int lo = 0, hi;
        if constexpr (KV_STAT) {
            hi = seqlen_q;
            if (p.window_right >= 0) lo = max(0, x - off - p.window_right);
            if (p.window_left >= 0) hi = min(hi, x - off + p.window_left + 1);
            if (x >= seqlen_k) hi = lo;
        } else {
            hi = seqlen_k;
            if (p.window_right >= 0) hi = min(hi, x + off + p.window_right + 1);
            if (p.window_left >= 0) lo = max(0, x + off - p.window_left);
            if (x >= seqlen_q) hi = lo;
        }
        const unsigned width = (unsigned)max(hi - lo, 0);

        float row_nl = 0.f, row_delta = 0.f;  // dQ pass: this row's -lse*log2e and delta
        if constexpr (!KV_STAT) {
            if (x < seqlen_q) {
                const int64_t idx = o_b * p.lse_stride_b + (int64_t)head0 * p.lse_stride_h + q_off + x;
                row_nl = neg_lse_log2(p.lse[idx]);
                row_delta = p.delta[idx];
            }
        }
        const float inv_cap = (FEAT && p.softcap > 0.f) ? 1.0f / p.softcap : 0.f;

        // dK/dV pass: column statistics of streamed tile t, one value per thread (2*BTS threads = BTS x {lse, delta})
        // The load returns the RAW value: converting it here would make the warp wait for the global load before it
        // starts on the current tile (ncu round 1: 8 % of the pass's samples); stat_conv runs one tile later.
        const bool stat_is_lse = (int)threadIdx.x < BTS;
        const float* stat_src = stat_is_lse ? p.lse : p.delta;
        auto load_stat = [&](int t) -> float {
            const int g = t / max(nt, 1);
            const int i = (t_lo + t - g * nt) * BTS + ((int)threadIdx.x % BTS);
            if (t >= n_tiles || i >= seqlen_q || (int)threadIdx.x >= 2 * BTS) return 0.f;
            return stat_src[o_b * p.lse_stride_b + (int64_t)(head0 + g) * p.lse_stride_h + q_off + i];
        };
        auto stat_conv = [&](float raw) -> float { return stat_is_lse ? neg_lse_log2(raw) : -raw; };
        float stat_next = 0.f;
        if constexpr (KV_STAT) stat_next = load_stat(0);

        for (int t = 0; t < n_tiles; ++t) {
            const int g = KV_STAT ? t / nt : 0;
            const int c0 = (t_lo + (KV_STAT ? t - g * nt : t)) * BTS + wg * CW;  // streamed index of my first column
            const float* tab = sTab + (t & 1) * (2 * BTS);
            if constexpr (KV_STAT) {
                if ((int)threadIdx.x < 2 * BTS) sTab[(t & 1) * (2 * BTS) + threadIdx.x] = stat_conv(stat_next);
                named_bar_sync(1, 256);
                stat_next = load_stat(t + 1);
            }
            float slope = 0.f;
            if constexpr (FEAT) {
                if (p.alibi) slope = p.alibi[batch * p.alibi_stride_b + head0 + g];
            }

            // ---- T1 -> P
            mbar_wait(bar_t1_full, t & 1);
            tc_fence_after();
            float pv[CW];
            if constexpr (CW == 64) tmem_ld_x64_wait(tT1, reinterpret_cast<uint32_t*>(pv));
            else tmem_ld_x32_wait(tT1, reinterpret_cast<uint32_t*>(pv));
            if constexpr (!KV_STAT) {  // dQ pass: nothing is written back over T1, release it right away
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_p_ready);
            }
            uint32_t fac[FEAT ? CW / 2 : 1];  // softcap: (1 - tanh^2) as f16 pairs
            if constexpr (FEAT) {
                // reference order (include/mat_mul.h:111-117): scale, ALiBi, softcap
                const int rel0 = KV_STAT ? c0 + off - x : x + off - c0;  // i + off - j at column 0
#pragma unroll
                for (int c = 0; c < CW; c += 2) {
                    float u[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        u[e] = pv[c + e] * p.scale - slope * fabsf((float)(KV_STAT ? rel0 + c + e : rel0 - c - e));
                        float f = 1.0f;
                        if (p.softcap > 0.f) {
                            const float th = tanh_approx(u[e] * inv_cap);
                            u[e] = p.softcap * th;
                            f = fmaf(-th, th, 1.0f);  // reference include/softmax.h:307-310
                        }
                        pv[c + e] = u[e] * kLog2e;
                        u[e] = f;
                    }
                    fac[c / 2] = pack_half2(u[0], u[1]);
                }
            }
            const bool need_mask = (c0 < lo) || ((unsigned)(c0 + CW - lo) > width);
            if (__any_sync(0xffffffffu, need_mask)) {
                const int base = c0 - lo;
#pragma unroll
                for (int c = 0; c < CW; ++c) pv[c] = ((unsigned)(base + c) < width) ? pv[c] : -INFINITY;
            }
#pragma unroll
            for (int c = 0; c < CW; c += 4) {
                float n0, n1, n2, n3;
                if constexpr (KV_STAT) {
                    const float4 nl = *reinterpret_cast<const float4*>(tab + wg * CW + c);
                    n0 = nl.x, n1 = nl.y, n2 = nl.z, n3 = nl.w;
                } else {
                    n0 = n1 = n2 = n3 = row_nl;
                }
                fma2(pv[c], pv[c + 1], sl2, sl2, n0, n1);
                fma2(pv[c + 2], pv[c + 3], sl2, sl2, n2, n3);
                // FA_BWD_EMU_COUNT of every FA_BWD_EMU_PERIOD pairs go to the FMA pipes (ex2_emu2), the rest to the MUFU
#pragma unroll
                for (int h = 0; h < 4; h += 2) {
                    if (FA_BWD_EMU_COUNT > 0 && (((c + h) / 2) % FA_BWD_EMU_PERIOD) >= FA_BWD_EMU_PERIOD - FA_BWD_EMU_COUNT) {
                        ex2_emu2(pv[c + h], pv[c + h + 1]);
                    } else {
                        pv[c + h] = ex2_approx(pv[c + h]);
                        pv[c + h + 1] = ex2_approx(pv[c + h + 1]);
                    }
                }
            }
            // dropout keep bits for my 64 columns (reference include/softmax.h:276-291: same index as the forward)
            uint32_t keep[2] = {0xffffffffu, 0xffffffffu};  // CW / 32 words are used
            if constexpr (DROPOUT) {
                const uint32_t k0 = (uint32_t)p.drop_seed, k1 = (uint32_t)(p.drop_seed >> 32);
                if constexpr (KV_STAT) {  // columns are query rows: every element has its own counter
#pragma unroll 1
                    for (int w = 0; w < CW / 32; ++w) {
                        uint32_t bits = 0u;
#pragma unroll 4
                        for (int c = 0; c < 32; ++c) {
                            const uint64_t idx = (uint64_t)(q_off + c0 + w * 32 + c) * (uint64_t)p.seqlen_k + (uint64_t)x;
                            const uint32_t k4 = philox_keep4(p.drop_offset + (idx >> 2), k0, k1, p.drop_thr);
                            bits |= ((k4 >> ((uint32_t)idx & 3u)) & 1u) << c;
                        }
                        keep[w] = bits;
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < CW / 32; ++w) {
                        const uint64_t idx0 = (uint64_t)(q_off + x) * (uint64_t)p.seqlen_k + (uint64_t)(c0 + w * 32);
                        const uint32_t sh = (uint32_t)idx0 & 3u;
                        const uint64_t ctr0 = p.drop_offset + (idx0 >> 2);
                        uint32_t lo_w = 0u, hi_w = 0u;
#pragma unroll
                        for (int q4 = 0; q4 < 8; ++q4) lo_w |= philox_keep4(ctr0 + q4, k0, k1, p.drop_thr) << (4 * q4);
                        if (sh != 0u) hi_w = philox_keep4(ctr0 + 8, k0, k1, p.drop_thr);
                        keep[w] = __funnelshift_r(lo_w, hi_w, sh);
                    }
                }
            }
            if constexpr (KV_STAT) {  // P (after dropout, unscaled) -> TMEM over my half of T1
                uint32_t pk[CW / 2];
#pragma unroll
                for (int c = 0; c < CW; c += 2) {
                    float a = pv[c], b = pv[c + 1];
                    if constexpr (DROPOUT) {
                        a = (keep[c >> 5] >> (c & 31)) & 1u ? a : 0.f;
                        b = (keep[c >> 5] >> ((c + 1) & 31)) & 1u ? b : 0.f;
                    }
                    pk[c / 2] = pack2<BF16>(a, b);
                }
                if constexpr (CW == 64) tmem_st_x32(tT1, pk);
                else tmem_st_x16(tT1, pk);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_p_ready);
            }

            // ---- T2 -> dS = P * (keep * rp * dP - delta)       (reference include/softmax.h:293-294)
            mbar_wait(bar_t2_full(t), t2_parity(t));
            tc_fence_after();
            const uint32_t tT2 = tT2_0 + t2_col(t);
            uint32_t dsk[CW / 2];
#pragma unroll
            for (int hc = 0; hc < CW / 32; ++hc) {
                float dp[32];
                tmem_ld_x32_wait(tT2 + hc * 32, reinterpret_cast<uint32_t*>(dp));
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const int cc = hc * 32 + c;
                    float a = dp[c], b = dp[c + 1];
                    if constexpr (DROPOUT) {
                        a = (keep[hc] >> c) & 1u ? a * p.rp_dropout : 0.f;
                        b = (keep[hc] >> (c + 1)) & 1u ? b * p.rp_dropout : 0.f;
                    }
                    float d0, d1;
                    if constexpr (KV_STAT) {
                        const float2 nd = *reinterpret_cast<const float2*>(tab + BTS + wg * CW + cc);
                        d0 = nd.x, d1 = nd.y;
                    } else {
                        d0 = d1 = -row_delta;
                    }
                    add2(a, b, d0, d1);
                    mul2(a, b, pv[cc], pv[cc + 1]);
                    if constexpr (FEAT) {
                        const float2 f = unpack_half2(fac[cc / 2]);
                        mul2(a, b, f.x, f.y);
                    }
                    dsk[cc / 2] = pack2<BF16>(a, b);
                }
            }
            if constexpr (CW == 64) tmem_st_x32(tT2, dsk);
            else tmem_st_x16(tT2, dsk);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ds_ready);
        }

        // ---- epilogue: warpgroup wg stores columns [wg*D/2, (wg+1)*D/2) of each accumulator
        const bool row_ok = x < (KV_STAT ? seqlen_k : seqlen_q);
        if (n_tiles > 0) {
            mbar_wait(bar_out_full, 0);
            tc_fence_after();
        }
        constexpr int HALF = Cfg::kOW / 2;
        const int col_base = SPLIT ? dhalf * Cfg::kOW : 0;  // first output column this CTA owns
        auto store_out = [&](uint32_t tmem_col, void* base, int64_t sb, int64_t ss, int64_t sh, int head, int row_off,
                             float mult) {
            uint16_t* dst = static_cast<uint16_t*>(base) + o_b * sb + (int64_t)(row_off + x) * ss + (int64_t)head * sh + col_base + wg * HALF;
            const bool wide_ok = __all_sync(0xffffffffu, (reinterpret_cast<uintptr_t>(dst) & 31) == 0);
#pragma unroll
            for (int c = 0; c < HALF; c += 32) {
                const int col = col_base + wg * HALF + c;  // columns [head_dim, D) are the tile's zero padding
                if (col >= p.head_dim) break;
                float o[32];
                if (n_tiles > 0) {
                    tmem_ld_x32_wait(tmem_base + lane_off + tmem_col + wg * HALF + c, reinterpret_cast<uint32_t*>(o));
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) o[e] = 0.f;
                }
                if (row_ok) {
                    uint32_t pk[16];
#pragma unroll
                    for (int e = 0; e < 32; e += 2) pk[e / 2] = pack2<BF16>(o[e] * mult, o[e + 1] * mult);
                    if (wide_ok && col + 32 <= p.head_dim) {
                        st_global_v8(dst + c, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
                        st_global_v8(dst + c + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; e += 4)
                            if (col + 2 * e < p.head_dim)
                                *reinterpret_cast<uint4*>(dst + c + 2 * e) = make_uint4(pk[e], pk[e + 1], pk[e + 2], pk[e + 3]);
                    }
                }
            }
        };
        if constexpr (KV_STAT) {
            store_out(Cfg::kTmemOut1, p.dk, p.dk_stride_b, p.dk_stride_s, p.dk_stride_h, kv_head, k_off, p.scale);
            store_out(Cfg::kTmemOut2, p.dv, p.dv_stride_b, p.dv_stride_s, p.dv_stride_h, kv_head, k_off,
                      DROPOUT ? p.rp_dropout : 1.0f);
        } else {
            store_out(Cfg::kTmemOut1, p.dq, p.dq_stride_b, p.dq_stride_s, p.dq_stride_h, head0, q_off, p.scale);
        }
    } else {
        reg_dec<48>();  // warps 10, 11: idle (they complete the third warpgroup for setmaxnreg)
    }

    // ------------------------------------------------------------------ teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc<512>(tmem_base);
}

}  // namespace fa
