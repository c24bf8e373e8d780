// Fused attention forward for sm_100a:  O = softmax(Q K^T * scale [+mask/bias/softcap]) V,  LSE.
//
// What it replaces: the reference's three forward kernels
//   flash_attention_forward_kernel          (reference kernel/fused_mha_forward.cu:25-224)
//   flash_attention_forward_varlen_kernel   (reference kernel/fused_mha_forward_varlen.cu:25-275)
//   flash_attention_kvcache_kernel, Sq > 1  (reference kernel/fused_mha_forward_kvcache.cu:24-295)
// and the per-tile building blocks they include (tile bounds include/template.h:36-112, score
// modifiers include/mat_mul.h:82-157, online softmax include/softmax.h:21-203, epilogue
// include/gemm_smem.h:119-205). Only the semantics are shared; the schedule is Blackwell-native:
//
//   * one CTA = 2 query tiles of 128 rows ("stages") x one (batch, head); KV streamed in 128-row tiles
//   * warp 13 : TMA producer + tile scheduler (decodes each work id once and publishes id + geometry through a
//               shared-memory mailbox; first K tile, Q, then K/V tiles into a ring of 128B-swizzled smem slots)
//   * warp 12 : tcgen05.mma issuer.  S = Q_s K^T (SS form) into the ONE S buffer the stages share, O_s += P_s V (TS
//               form, P read from TMEM, V consumed as the MN-major B operand).  Accumulators never leave TMEM.
//               Issue order per iteration: QK0(it) PV0(it-1) QK1(it) PV1(it-1).
//   * warps 0-3 / 4-7 : softmax for stage 0 / 1.  thread == row (tcgen05.ld 32x32b), so the row max / row sum are
//               thread-local; S is copied to registers (which frees the S buffer for the other stage's next
//               Q K^T), P is written to the stage's own P buffer in TMEM as 16-bit pairs.
//   * warps 8-11 : correction (lazy rescale of O in TMEM when the running max moved by > 2^8)
//               and the epilogue (O / l -> 16-bit -> global, LSE).
//   * warp 14 : zeroes the rows of a ragged V tile past seqlen_k;  warp 15 : watchdog.
//   * everything is ordered with mbarriers; tcgen05.commit signals MMA completion.
//
// TMEM map (512 columns), FA_SHARED_S = 1 (default):  S [0,128)  P0 [128,192)  P1 [192,256)  O0 [256,256+D)
// O1 [256+D,256+2D).  A softmax warp has its S tile in registers ~200 clocks after it lands, so one S buffer serves
// both stages in turn and Q K^T(j+1) of a stage never waits for that stage's softmax(j) -> P V(j) (see FwdConfig).
// FA_SHARED_S = 0 (round 1, kept for A/B): S0 [0,128) S1 [128,256), P_s aliases columns [64,128) of S_s.
//
// head_dim 256: an O accumulator of 256 columns per query tile leaves no room for two stages of different rows, so a
// work item is ONE 128-row query block and stage 0 runs alone: one S = Q K^T over all 256 dims, one softmax, one P V
// with N = 256 into O0|O1 (contiguous columns); the Q K^T of tile j+1 runs under the softmax of tile j as at head_dim
// 128 (shared S, private P). Stage 1's warps only take part in the work hand-out. One 64 KB K tile and one V tile
// fit beside Q, so they are handed over in 32 KB halves (FwdConfig::kHalfRing): Q K^T releases the first 128 dims of
// K while it works on the second, P V is issued per 128 output columns, and the loader refills each half as it frees
// up, K one tile ahead of V. K+V are 128 KB per tile: the SM's ingest rate, not the tensor pipe, bounds this path.
// FA_SPLIT_SINGLE = 0 (round 1, "split-D", kept for A/B): both stages work on the same rows, each computes S and
// the softmax again and owns one 128-column half of O (stage s multiplies P by V[:, 128 s .. 128 s + 128)).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "kvcache_prep.cuh"
#include "ptx_sm100.cuh"
#include "umma_issue_gen.cuh"

namespace fa {

struct alignas(64) FwdKernelParams {
    CUtensorMap tm_q;  // 4-D (head_dim, heads, rows, batch), box (64, 1, 128, 1), 128B swizzle
    CUtensorMap tm_k;
    CUtensorMap tm_v;
    void* out;
    float* lse;
    int64_t o_stride_b, o_stride_s, o_stride_h;  // elements
    int64_t lse_stride_b, lse_stride_h;          // elements; row stride is 1
    const int* cu_seqlens_q;
    const int* cu_seqlens_k;
    const int* seqused_k;
    const int* cache_seqlens;
    const int* cache_batch_idx;
    const int* leftpad_k;
    const int* block_table;
    const float* alibi;
    int64_t alibi_stride_b;
    int block_table_stride;
    int page_size;
    int seqlen_q;
    int seqlen_k;
    int seqlen_k_add;  // rows appended to the cache before this call (kvcache)
    int num_heads;
    int heads_per_kv;
    int head_dim;      // real head dim (multiple of 8, <= D): TMA zero-fills columns [head_dim, D) of every tile
    float scale;
    float scale_log2;
    float softcap;
    int window_left;   // -1 = unbounded
    int window_right;  // -1 = unbounded; causal is window_right = 0
    int reverse_m;     // launch the longest query blocks first (causal / local)
    // 1-D launch order (non-decode): CTAs are numbered section by section; a section is `section_bh`
    // (batch, head) pairs whose K/V fit L2 together, and inside a section all pairs' longest query block
    // come first (longest-processing-time-first across the section, not just inside one head).
    int num_m_blocks;
    int num_bh;
    int section_bh;
    // decode mode (packed GQA + split-KV): the 128 tile rows are (query position, head-in-group) pairs
    int gqa_pack;        // query heads per KV head packed into the row dimension
    int num_splits;      // CTAs along the KV length per (batch, kv head); > 1 => partial results
    float* o_partial;    // [split][batch][head][seqlen_q][D] fp32, normalised per split
    float* lse_partial;  // [split][batch][head][seqlen_q] fp32, -inf for an empty split
    // Tile scheduler (non-decode). Default: persistent grid (one CTA per SM); CTA c starts with work id c and draws
    // further ids from `sched[0]` (a counter private to this launch, zeroed on the launch stream right before the
    // kernel; NULL when the grid already covers every item). Ids are numbered longest-first, and the counter hands
    // them out in exactly that order. -DFA_SCHED_CLC=1 builds the stateless alternative: one CTA per work item, the
    // running CTAs cancel not-yet-started ones through cluster launch control and take their ids -- measured 3 %
    // (head_dim 128) to 9 % (head_dim 64) slower on causal shapes because the launch unit hands CTAs out only
    // approximately in order (tools/ubench/clc_order.cu, profiles/experiments/README.md).
    int* sched;
    // Debug counters (tests only; NULL in normal use): [0] += softmax rows whose running maximum moved past the
    // lazy-rescale threshold, [1] += O-accumulator rescales executed by the correction warps (per warp).
    unsigned long long* dbg_counters;
    // dropout (reference include/softmax.h:96-125, include/philox.h): keep iff Philox word <= drop_thr;
    // the 1/(1-p) factor is folded into the epilogue's 1/l. dmask: optional +-1.0 sign tensor.
    float rp_dropout;       // 1 / (1 - p); 1 when dropout is off
    uint32_t drop_thr;      // (1 - p) * (2^32 - 1), the reference's float expression
    uint64_t drop_seed;
    uint64_t drop_offset;
    uint16_t* dmask;
    int64_t dmask_stride_b, dmask_stride_h, dmask_stride_row;  // elements; row = q_off + position
};

// Philox4x32-10 (reference include/philox.h:13-64): counter = (c0, c1, 0, 0), key = (k0, k1).
FA_DEVICE uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
    uint32_t x0 = c0, x1 = c1, x2 = 0u, x3 = 0u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
        x0 = hi1 ^ x1 ^ k0;
        x1 = lo1;
        x2 = hi0 ^ x3 ^ k1;
        x3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(x0, x1, x2, x3);
}

// Keep bits of the 4 flat indices {4g .. 4g+3} (bit t = index 4g+t is kept).
FA_DEVICE uint32_t philox_keep4(uint64_t ctr, uint32_t k0, uint32_t k1, uint32_t thr) {
    const uint4 r = philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), k0, k1);
    return (r.x <= thr ? 1u : 0u) | (r.y <= thr ? 2u : 0u) | (r.z <= thr ? 4u : 0u) | (r.w <= thr ? 8u : 0u);
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kNegSentinel = -1e30f;  // reference NEG_INF (include/kernel.h:20) as the "no keys" LSE
constexpr float kRescaleThreshold = 8.0f;  // log2 units; O is rescaled only when the max moves more

// Tuning knobs (compile-time; csrc/build.sh can override them with -D for A/B runs on the GPU).
#ifndef FA_HALF_RING
#define FA_HALF_RING 1  // head_dim 256 single stage: K / V tiles handed over in 32 KB halves (0: whole 64 KB tiles)
#endif
#ifndef FA_SPLIT_SINGLE
#define FA_SPLIT_SINGLE 1  // head_dim 256: one Q K^T and one softmax per tile, one N=256 P V (0: round-1 split-D, both stages)
#endif
#ifndef FA_EMU_PERIOD
#define FA_EMU_PERIOD 4  // of every FA_EMU_PERIOD pairs of exponentials ...
#endif
#ifndef FA_EMU_COUNT
#define FA_EMU_COUNT 1   // ... this many are evaluated on the FMA pipes (ex2_emu2) instead of the MUFU
#endif
#ifndef FA_SPLIT_P
#define FA_SPLIT_P 1     // signal the MMA warp after 3/4 of P so P V starts before the last quarter
#endif
#ifndef FA_SHARED_S
#define FA_SHARED_S 1    // one S buffer shared by both stages + a private P buffer per stage (see "TMEM map"): the next
#endif                   // Q K^T of a stage no longer waits for that stage's softmax -> P V chain
#ifndef FA_LD_OVERLAP
#define FA_LD_OVERLAP 0  // softmax: split the S load so the row max of the first half overlaps the second half's load
#endif
#ifndef FA_SCHED_CLC
#define FA_SCHED_CLC 0   // 1: stateless tile scheduler through cluster launch control (see FwdKernelParams::sched)
#endif
#ifndef FA_EARLY_QK
#define FA_EARLY_QK 1    // head_dim <= 64: issue the left half of the next S = Q K^T as soon as the softmax warps
#endif                   // hold S in registers (see kEarlyQK)

// Optional event trace (build with -DFA_TRACE): lane 0 of every warp of CTA 0 appends (event, clock64)
// pairs to a global buffer set through fa_b200_debug_set_trace(); tools/trace_timeline.py prints the
// per-iteration timeline. Compiled out of the product build.
#ifdef FA_TRACE
__device__ long long* g_fa_trace = nullptr;
#define FA_TRACE_DECL int trace_n = 0; long long* trace_buf = (blockIdx.x == 0 && g_fa_trace) ? g_fa_trace + (threadIdx.x >> 5) * 4096 : nullptr
#define FA_TRACE_EV(ev)                                                   \
    do {                                                                  \
        if (trace_buf && (threadIdx.x & 31) == 0 && trace_n < 2040) {     \
            trace_buf[trace_n * 2] = (ev);                                \
            trace_buf[trace_n * 2 + 1] = clock64();                       \
            ++trace_n;                                                    \
        }                                                                 \
    } while (0)
#else
#define FA_TRACE_DECL
#define FA_TRACE_EV(ev)
#endif

template <int D>
struct FwdConfig {
    static constexpr int kBlockM = 128;
    static constexpr int kBlockN = 128;
    static constexpr int kTileBytes = kBlockN * D * 2;
    static constexpr int kHalfBytes = kBlockN * 128;  // one 64-column (128-byte) swizzle block
    static constexpr bool kSplitD = (D == 256);            // see "split-D" in the header comment
    static constexpr int kDO = kSplitD ? 128 : D;          // columns of one stage's O accumulator
    static constexpr int kItemRows = kSplitD ? kBlockM : 2 * kBlockM;  // query rows of one work item
    // head_dim 256, single stage: the K and V tiles (64 KB each, one of each fits) are handed over in halves of two
    // swizzle blocks (head dims [0,128) and [128,256)): Q K^T runs over the first half of the dims and releases it while it
    // works on the second, P V is issued per 128 output columns -- so the next tile's loads start half a GEMM earlier and
    // the ring has four 32 KB slots K-lo K-hi V-lo V-hi instead of two of 64 KB (same bytes, same layout in memory).
    static constexpr bool kHalfRing = kSplitD && FA_SPLIT_SINGLE && FA_HALF_RING;
    static constexpr int kKvStages = (D == 256) ? (kHalfRing ? 4 : 2) : (D == 128) ? 4 : 6;
    static constexpr int kSlotBytes = kHalfRing ? kTileBytes / 2 : kTileBytes;
    static constexpr int kEntriesPerTile = kHalfRing ? 4 : 2;  // ring entries of one KV tile
    static constexpr int kSmemQ = kSplitD ? kTileBytes : 2 * kTileBytes;
    static constexpr int kSmemKV = kKvStages * kSlotBytes;
    static constexpr int kNumBars = 2 + 2 * kKvStages + 6 * 2 + 1 + 2 + 4 + 1 + 2 + 1 + 1 + 2 + 2;
    static constexpr int kOffBars = kSmemQ + kSmemKV;
    static constexpr int kOffTmemPtr = kOffBars + 8 * kNumBars;
    static constexpr int kOffScale = (kOffTmemPtr + 16 + 15) & ~15;
    static constexpr int kOffRowSum = kOffScale + 2 * 2 * 128 * 4;  // (sScale: [tile parity][stage][128]) [item parity][stage][128]
    static constexpr int kOffRowMax = kOffRowSum + 2 * 2 * 128 * 4;  // [item parity][stage][128]
    static constexpr int kOffSched = kOffRowMax + 2 * 2 * 128 * 4;   // int[2] work ids (+ 8 bytes of padding)
    static constexpr int kOffClc = kOffSched + 16;                   // 16-byte cluster-launch-control response
    static constexpr int kOffGeom = kOffClc + 16;                   // 2 x WorkHead (48 bytes each): the work mailbox's payload
    static constexpr int kSmemUsed = kOffGeom + 2 * 48;
    static constexpr int kSmemBytes = kSmemUsed + 1024;  // slack for manual 1024-byte alignment
    // TMEM map. FA_SHARED_S = 0: S0 [0,128) S1 [128,256), P_s = upper half of S_s.
    // FA_SHARED_S = 1: ONE S buffer [0,128) that the stages use in turn (a softmax warp copies its S tile to registers
    // within ~200 clocks, after which the buffer is free for the other stage's next Q K^T) and a private P buffer per
    // stage, P0 [128,192) P1 [192,256). P_s(j) no longer shares columns with S_s(j+1), so Q K^T(j+1) of a stage is
    // issued BEFORE P V(j) of that stage and is usually complete by the time the stage's softmax warps finish tile j:
    // they go from tile to tile without waiting for the tensor pipe (round 1: one third of their time was that wait).
    static constexpr bool kSharedS = FA_SHARED_S != 0;
    static constexpr int kTmemS0 = 0, kTmemS1 = kSharedS ? 0 : 128, kTmemO0 = 256, kTmemO1 = 256 + kDO;
    static constexpr int kTmemP0 = kSharedS ? 128 : 64, kTmemP1 = kSharedS ? 192 : 128 + 64;
    static constexpr int kTmemPOff = 64;  // forward v2 (fwd2_sm100.cuh): P_i = upper half of S_i
    // Early left half of S: two N=64 MMAs read the Q tile from shared memory twice, and an SS-form 128x128x16 MMA
    // already needs the full 128 B/clk of shared-memory bandwidth -- measured -10 % at head_dim 128 (tensor pipe is
    // the bottleneck there), +7 % at head_dim 64 (the softmax is, and the pipe has slack).
    static constexpr bool kEarlyQK = FA_EARLY_QK && (D == 64) && !kSharedS;
};

// Per-sequence geometry shared by every role.
struct SeqGeom {
    int q_off, q_b, seqlen_q;
    int k_off, k_b, seqlen_k;
};

FA_DEVICE SeqGeom load_geom(const FwdKernelParams& p, int batch) {
    SeqGeom g;
    g.q_off = 0;
    g.q_b = batch;
    g.seqlen_q = p.seqlen_q;
    g.k_off = 0;
    g.k_b = batch;
    g.seqlen_k = p.seqlen_k;
    if (p.cu_seqlens_q) {
        g.q_off = p.cu_seqlens_q[batch];
        g.seqlen_q = p.cu_seqlens_q[batch + 1] - g.q_off;
        g.q_b = 0;
    }
    if (p.cu_seqlens_k) {
        const int s = p.cu_seqlens_k[batch];
        g.seqlen_k = p.cu_seqlens_k[batch + 1] - s;
        if (!p.block_table) {
            g.k_off = s;
            g.k_b = 0;
        }
    }
    if (p.cache_seqlens) g.seqlen_k = p.cache_seqlens[batch] + p.seqlen_k_add;
    if (p.seqused_k) {  // reference include/template.h:65-68: used > 0 ? min(len, used) : 0
        const int u = p.seqused_k[batch];
        g.seqlen_k = u > 0 ? min(g.seqlen_k, u) : 0;
    }
    if (p.cache_batch_idx) g.k_b = p.cache_batch_idx[batch];
    if (p.leftpad_k) g.k_off += p.leftpad_k[batch];
    return g;
}

// Geometry of one work item (a 256-row query block of one (batch, head); in DECODE mode one KV split
// of one (batch, kv head) GQA group). Every warp role recomputes it from the work id.
struct WorkGeom {
    SeqGeom g;
    int head, batch, kv_head, split;
    int m0, m_end;       // query positions [m0, m_end)
    int off;             // seqlen_k - seqlen_q (bottom-right alignment)
    int n_min, n_max;    // KV tiles [n_min, n_max)
    int n_tiles;
    int it_lo[2], it_hi[2];  // iterations each stage takes part in (iteration it = tile n_max-1-it)
    int o_b;
    bool ragged_tail;    // the first tile processed (n_max-1) extends past seqlen_k: its V rows need zeroing
    bool skip;           // nothing to do and nothing to write (query block past the sequence end)
};

// The expensive half of decoding a work id: the integer divisions of the launch order and the per-sequence loads.
// 12 words; the loader publishes it through the work mailbox, the other roles finish it with shifts and min/max.
struct alignas(16) WorkHead {
    SeqGeom g;
    int m_block, head, batch, kv_head, split, pad;
};
static_assert(sizeof(WorkHead) == 48, "WorkHead is copied as three 16-byte words");
// field-by-field on purpose: taking the struct's address would put it in local memory
FA_DEVICE void work_head_store(uint32_t smem, const WorkHead& h) {
    st_shared_v4(smem, (uint32_t)h.g.q_off, (uint32_t)h.g.q_b, (uint32_t)h.g.seqlen_q, (uint32_t)h.g.k_off);
    st_shared_v4(smem + 16, (uint32_t)h.g.k_b, (uint32_t)h.g.seqlen_k, (uint32_t)h.m_block, (uint32_t)h.head);
    st_shared_v4(smem + 32, (uint32_t)h.batch, (uint32_t)h.kv_head, (uint32_t)h.split, 0u);
}
FA_DEVICE WorkHead work_head_load(uint32_t smem) {
    uint32_t a0, a1, a2, a3, b0, b1, b2, b3, c0, c1, c2, c3;
    ld_shared_v4(smem, a0, a1, a2, a3);
    ld_shared_v4(smem + 16, b0, b1, b2, b3);
    ld_shared_v4(smem + 32, c0, c1, c2, c3);
    WorkHead h;
    h.g.q_off = (int)a0; h.g.q_b = (int)a1; h.g.seqlen_q = (int)a2; h.g.k_off = (int)a3;
    h.g.k_b = (int)b0; h.g.seqlen_k = (int)b1; h.m_block = (int)b2; h.head = (int)b3;
    h.batch = (int)c0; h.kv_head = (int)c1; h.split = (int)c2; h.pad = 0;
    return h;
}

template <bool DECODE>
FA_DEVICE WorkHead work_head(const FwdKernelParams& p, int work_id) {
    WorkHead h;
    const int G = DECODE ? p.gqa_pack : 1;
    h.m_block = 0;
    h.split = 0;
    h.pad = 0;
    if constexpr (DECODE) {
        h.split = (int)blockIdx.x;
        h.head = (int)blockIdx.y * G;  // first query head of the group
        h.batch = blockIdx.z;
        h.kv_head = blockIdx.y;
    } else {
        // sectioned longest-first order, see FwdKernelParams::section_bh
        const int per_section = p.section_bh * p.num_m_blocks;
        const int sec = work_id / per_section;
        const int r = work_id - sec * per_section;
        const int sec_n = min(p.section_bh, p.num_bh - sec * p.section_bh);
        const int m_rank = r / sec_n;
        const int bh = sec * p.section_bh + (r - m_rank * sec_n);
        h.m_block = p.reverse_m ? p.num_m_blocks - 1 - m_rank : m_rank;
        h.head = bh % p.num_heads;
        h.batch = bh / p.num_heads;
        h.kv_head = h.head / p.heads_per_kv;
    }
    h.g = load_geom(p, h.batch);
    return h;
}

template <bool DECODE, bool SPLIT = false>
FA_DEVICE WorkGeom finish_geom(const FwdKernelParams& p, const WorkHead& h) {
    constexpr int BM = 128, BN = 128;
    constexpr int ROWS = SPLIT ? BM : 2 * BM;  // query rows per work item
    WorkGeom w;
    const int m_block = h.m_block;
    w.split = h.split;
    w.head = h.head;
    w.batch = h.batch;
    w.kv_head = h.kv_head;
    w.g = h.g;
    w.m0 = m_block * ROWS;
    w.skip = w.m0 >= w.g.seqlen_q;  // over-provisioned varlen grid
    w.m_end = DECODE ? w.g.seqlen_q : min(w.m0 + ROWS, w.g.seqlen_q);
    w.off = w.g.seqlen_k - w.g.seqlen_q;
    w.o_b = p.cu_seqlens_q ? 0 : w.batch;
    int n_max = (w.g.seqlen_k + BN - 1) / BN;
    if (p.window_right >= 0) {
        const int max_col = w.m_end - 1 + w.off + p.window_right;
        n_max = min(n_max, max_col < 0 ? 0 : max_col / BN + 1);
    }
    int n_min = 0;
    if (p.window_left >= 0) {
        const int min_col = w.m0 + w.off - p.window_left;
        n_min = max(0, min_col >= 0 ? min_col / BN : 0);
    }
    if constexpr (DECODE) {  // this CTA's share of the KV tiles
        const int per = (max(n_max - n_min, 0) + p.num_splits - 1) / p.num_splits;
        const int lo = n_min + w.split * per;
        n_max = min(n_max, lo + per);
        n_min = min(lo, n_max);
    }
    w.n_min = n_min;
    w.n_max = n_max;
    w.n_tiles = w.skip ? 0 : max(n_max - n_min, 0);
    w.ragged_tail = w.n_tiles > 0 && n_max * BN > w.g.seqlen_k;
    // Stage s (rows m0+128s ..) only takes part in iterations [it_lo[s], it_hi[s]): tiles wholly above its
    // causal diagonal or wholly left of its window are skipped for that stage.
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int r0 = w.m0 + (SPLIT ? 0 : s * BM);
        int hi_n = n_max, lo_n = n_min;
        if (SPLIT && FA_SPLIT_SINGLE && s == 1) {
            hi_n = lo_n = n_min;  // head_dim 256: stage 0 alone (its P V covers all 256 output columns)
        } else if (DECODE) {
            if (s == 1 && !SPLIT) hi_n = lo_n = n_min;  // single tile of packed rows: stage 1 is idle
        } else if (r0 >= w.g.seqlen_q) {
            hi_n = lo_n = n_min;  // no valid row in this stage
        } else {
            if (p.window_right >= 0) {
                const int max_col = min(r0 + BM, w.g.seqlen_q) - 1 + w.off + p.window_right;
                hi_n = min(n_max, max_col < 0 ? 0 : max_col / BN + 1);
            }
            if (p.window_left >= 0) {
                const int min_col = r0 + w.off - p.window_left;
                lo_n = max(n_min, min_col >= 0 ? min_col / BN : 0);
            }
            hi_n = max(hi_n, lo_n);
        }
        w.it_lo[s] = n_max - hi_n;
        w.it_hi[s] = n_max - lo_n;
        if (w.n_tiles == 0) w.it_lo[s] = w.it_hi[s] = 0;
    }
    return w;
}

template <bool DECODE, bool SPLIT = false>
FA_DEVICE WorkGeom work_geom(const FwdKernelParams& p, int work_id) {
    return finish_geom<DECODE, SPLIT>(p, work_head<DECODE>(p, work_id));
}

// DECODE = false: persistent grid (one CTA per SM); work items = (256-row query block, head, batch) handed
//                 out longest-first by an atomic counter, so the prologue of the next item (Q/K loads, first
//                 QK^T) overlaps the epilogue of the current one and TMEM / barriers are set up once.
// DECODE = true : grid (KV splits, KV heads, batch), one item per CTA; ONE tile whose 128 rows are the
//                 (position, head) pairs of a whole GQA group (row = position * G + head_in_group), so a
//                 decode step reads each K/V byte once per group; only stage 0 runs. HBM-bound by construction.
template <int D, bool BF16, bool FEAT, bool DECODE = false, bool DROPOUT = false>
__global__ void __launch_bounds__(512, 1)
fa_fwd_sm100_kernel(const __grid_constant__ FwdKernelParams p) {
    using Cfg = FwdConfig<D>;
    constexpr int BM = Cfg::kBlockM, BN = Cfg::kBlockN;
    constexpr int KV = Cfg::kKvStages;
    constexpr bool SPLIT = Cfg::kSplitD;
    constexpr int DO = Cfg::kDO;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_raw_u32 = smem_u32(smem_raw);
    const uint32_t sbase = (smem_raw_u32 + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_raw_u32);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    FA_TRACE_DECL;
    const int G = DECODE ? p.gqa_pack : 1;
    const int total_work = DECODE ? 1 : p.num_m_blocks * p.num_bh;
    const int num_batch = DECODE ? (int)gridDim.z : p.num_bh / p.num_heads;
    const bool partial_out = DECODE && p.num_splits > 1;

    // ------------------------------------------------------------------ shared-memory carve-up
    const uint32_t sQ = sbase;
    const uint32_t sKV = sbase + Cfg::kSmemQ;
    const uint32_t bars = sbase + Cfg::kOffBars;
    auto bar_q_full = [&](int s) { return bars + 8 * s; };
    auto bar_kv_full = [&](int i) { return bars + 8 * (2 + i); };
    auto bar_kv_empty = [&](int i) { return bars + 8 * (2 + KV + i); };
    constexpr int kB0 = 2 + 2 * KV;
    auto bar_s_full = [&](int s) { return bars + 8 * (kB0 + s); };       // MMA -> softmax: S_s ready
    auto bar_p_full = [&](int s) { return bars + 8 * (kB0 + 2 + s); };   // softmax+correction -> MMA
    // softmax -> correction: scale; double-buffered by tile parity (the softmax of a stage may publish tile j+1 before
    // the correction warps have looked at tile j: they serve both stages in turn)
    auto bar_stats = [&](int s, int par) { return bars + 8 * (par ? kB0 + 26 + s : kB0 + 4 + s); };
    auto bar_o_full = [&](int s) { return bars + 8 * (kB0 + 6 + s); };   // MMA -> correction: O_s final
    auto bar_p_last = [&](int s) { return bars + 8 * (kB0 + 8 + s); };   // softmax -> MMA: last 1/4 of P
    auto bar_final = [&](int s, int b) { return bars + 8 * (kB0 + 10 + 2 * s + b); };  // softmax -> corr.: l, m
    const uint32_t bar_q_empty = bars + 8 * (kB0 + 14);                  // MMA -> loader: Q tiles consumed
    auto bar_sched_full = [&](int b) { return bars + 8 * (kB0 + 15 + b); };   // loader -> everyone: work id
    auto bar_sched_empty = [&](int b) { return bars + 8 * (kB0 + 17 + b); };  // everyone -> loader
    const uint32_t bar_vfix = bars + 8 * (kB0 + 19);  // sanitiser -> MMA: tail rows of the ragged V tile zeroed
    auto bar_s_loaded = [&](int s) { return bars + 8 * (kB0 + 20 + s); };  // softmax -> MMA: S_s is in registers
    const uint32_t bar_clc = bars + 8 * (kB0 + 22);  // launch unit -> loader: cancellation response landed
    const uint32_t bar_done = bars + 8 * (kB0 + 23);  // role warps -> watchdog: this warp has left its loop
    auto bar_p_free = [&](int s) { return bars + 8 * (kB0 + 24 + s); };  // MMA -> softmax, correction: P V_s of a tile done
    const uint32_t bar_sx_free = bar_s_loaded(0);  // shared-S mode: softmax (either stage) -> MMA: the S buffer is in registers
    static_assert(kB0 + 28 <= Cfg::kNumBars, "barrier table too small");
    const uint32_t sClc = sbase + Cfg::kOffClc;
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(sgen + Cfg::kOffTmemPtr);
    float* sScale = reinterpret_cast<float*>(sgen + Cfg::kOffScale);
    float* sRowSum = reinterpret_cast<float*>(sgen + Cfg::kOffRowSum);
    float* sRowMax = reinterpret_cast<float*>(sgen + Cfg::kOffRowMax);
    volatile int* sSched = reinterpret_cast<volatile int*>(sgen + Cfg::kOffSched);
    volatile uint32_t* sWatch = reinterpret_cast<volatile uint32_t*>(sgen + Cfg::kOffSched + 8);  // progress, warps done

    if (warp == 13 && lane == 0) {
        sWatch[0] = 0;
        sWatch[1] = 0;
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_q_full(s), 1);
            mbar_init(bar_s_full(s), 1);
            mbar_init(bar_p_full(s), 8);  // 4 softmax warps + 4 correction warps
            mbar_init(bar_stats(s, 0), 4);
            mbar_init(bar_stats(s, 1), 4);
            mbar_init(bar_o_full(s), 1);
            mbar_init(bar_p_last(s), 4);
            mbar_init(bar_s_loaded(s), 4);
            mbar_init(bar_final(s, 0), 4);
            mbar_init(bar_final(s, 1), 4);
            mbar_init(bar_p_free(s), 1);
            mbar_init(bar_sched_full(s), 1);
            mbar_init(bar_sched_empty(s), 14);  // MMA warp + 8 softmax + 4 correction warps + V sanitiser
        }
        mbar_init(bar_q_empty, 1);
        mbar_init(bar_vfix, 1);
        mbar_init(bar_clc, 1);
        mbar_init(bar_done, 15);
        for (int i = 0; i < KV; ++i) {
            mbar_init(bar_kv_full(i), 1);
            mbar_init(bar_kv_empty(i), 1);
        }
        mbar_fence_init();
        tma_prefetch_desc(&p.tm_q);
        tma_prefetch_desc(&p.tm_k);
        tma_prefetch_desc(&p.tm_v);
    }
    if (warp == 12) tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if constexpr (DECODE) {
        // programmatic dependent launch: everything above overlapped the kv-cache preparation kernel; its writes
        // (appended K/V rows, the rotated Q) are visible after the wait. The combine kernel may start likewise.
        pdl_wait_primary();
        pdl_launch_dependents();
    }

    // Consumer side of the scheduler: k-th work id of this CTA (>= total_work means "no more work").
    // The loader decodes a work id ONCE (work_geom: a dozen integer divisions and, for var-len, dependent loads of the
    // cu_seqlens entries -- ~2500 clocks in a 48-register warp, measured on the MMA warp between two items) and publishes
    // the decoded geometry next to the id; every other role copies it out of shared memory.
    const uint32_t sGeom = sbase + Cfg::kOffGeom;
    struct Work {
        int id;
        WorkHead h;
    };
    auto get_work = [&](int k) -> Work {
        Work r;
        if constexpr (DECODE) {
            r.id = k == 0 ? 0 : total_work;
            r.h = work_head<DECODE>(p, 0);
        } else {
            mbar_wait(bar_sched_full(k & 1), (k >> 1) & 1);
            r.id = sSched[k & 1];
            r.h = work_head_load(sGeom + (k & 1) * 48);  // (garbage behind the end-of-work id: never used)
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_sched_empty(k & 1));
        }
        return r;
    };

    if (warp == 13) {
        // ============================================================ TMA producer + tile scheduler
        reg_dec<48>();
        auto load_tile = [&](const CUtensorMap* tm, uint32_t dst, uint32_t bar, int h, int row, int b) {
            mbar_arrive_expect_tx(bar, Cfg::kTileBytes);
#pragma unroll
            for (int c = 0; c < D / 64; ++c)
                tma_load_4d(dst + c * Cfg::kHalfBytes, tm, bar, c * 64, h, row, b);
        };
        int ring = 0;  // K of iteration it is ring entry (base + 2*it), V is (base + 2*it + 1)
        int ka = 0;    // items with work so far
        int id = DECODE ? 0 : (int)blockIdx.x;
#if FA_SCHED_CLC
        // Work stealing through cluster launch control: ask for the blockIdx of a CTA that has not started
        // yet; once a request fails (every CTA of the grid has started or was cancelled) no more are issued.
        bool clc_more = !DECODE;
        int clc_n = 0;  // requests issued so far (phase of bar_clc)
        auto fetch_issue = [&]() {
            if (lane == 0) {
                mbar_arrive_expect_tx(bar_clc, 16);
                clc_try_cancel(sClc, bar_clc);
            }
        };
        auto fetch_read = [&]() -> int {
            mbar_wait(bar_clc, clc_n & 1);
            ++clc_n;
            uint32_t x = 0;
            const bool ok = clc_query(sClc, x);
            fence_proxy_async_smem();  // the response was read before the next request's async write
            __syncwarp();
            if (!ok) clc_more = false;
            return ok ? (int)x : total_work;
        };
#else
        // next unclaimed work id from this launch's counter (one atomic per CTA, broadcast to the warp)
        bool clc_more = !DECODE && p.sched != nullptr;
        int pending_raw = 0;  // lane 0: the counter value on its way back (nothing waits for it until fetch_read)
        auto fetch_issue = [&]() {
            if (lane == 0) pending_raw = atomicAdd(p.sched, 1) + (int)gridDim.x;
        };
        auto fetch_read = [&]() -> int {
            const int pending_id = __shfl_sync(0xffffffffu, pending_raw, 0);
            if (pending_id >= total_work) clc_more = false;
            return pending_id < total_work ? pending_id : total_work;
        };
#endif
        auto fetch = [&]() -> int {
            if (!clc_more) return total_work;
            fetch_issue();
            return fetch_read();
        };
        WorkGeom w;
        for (int k = 0;; ++k) {
            if constexpr (!DECODE) {
                // Query blocks past the end of their sequence (the var-len grid is sized for the longest
                // sequence) have nothing to compute or write: drop them here instead of sending every warp
                // through a scheduler hand-shake for them.
                WorkHead h;
                while (id < total_work) {
                    h = work_head<DECODE>(p, id);
                    w = finish_geom<DECODE, SPLIT>(p, h);
                    if (!w.skip) break;
                    id = fetch();
                }
                // publish the k-th work id and its geometry (slot k&1 is free once everyone consumed item k-2)
                if (k >= 2) mbar_wait(bar_sched_empty(k & 1), ((k >> 1) - 1) & 1);
                if (lane == 0) {
                    sSched[k & 1] = id;
                    if (id < total_work) work_head_store(sGeom + (k & 1) * 48, h);
                    mbar_arrive(bar_sched_full(k & 1));
                }
                __syncwarp();
            } else {
                if (k > 0) id = total_work;
                else w = work_geom<DECODE, SPLIT>(p, id);
            }
            if (id >= total_work) break;
            FA_TRACE_EV(303);  // loader: work id published
            const bool prefetching = clc_more;
            if (prefetching) fetch_issue();  // early: the request's latency hides behind the loads
            if (w.n_tiles > 0) {
                auto kv_coords = [&](int n, int& row, int& b) {
                    const int r = n * BN;
                    if (p.block_table) {
                        const int page = r / p.page_size;
                        b = p.block_table[(int64_t)w.batch * p.block_table_stride + page];
                        row = r - page * p.page_size;
                    } else {
                        b = w.g.k_b;
                        row = w.g.k_off + r;
                    }
                };
                constexpr int EPT = Cfg::kEntriesPerTile;
                const int base = ring;  // tile `it` of the item: K = ring entries base + EPT*it .., V = base + EPT*it + EPT/2 ..
                auto produce = [&](const CUtensorMap* tm, int n, int entry) {
                    int row = 0, b = 0;
                    if (lane == 0) kv_coords(n, row, b);
#pragma unroll
                    for (int half = 0; half < (Cfg::kHalfRing ? 2 : 1); ++half) {
                        const int slot = (entry + half) % KV;
                        const uint32_t parity = (((entry + half) / KV) & 1) ^ 1;
                        mbar_wait(bar_kv_empty(slot), parity);
                        FA_TRACE_EV(310);  // loader: ring slot free
                        if (lane == 0) {
                            const uint32_t dst = sKV + slot * Cfg::kSlotBytes, bar = bar_kv_full(slot);
                            if (Cfg::kHalfRing) {  // two of the tile's four 64-column blocks
                                // (head_dim <= 192: the fourth block would be all padding -- it is neither loaded nor read)
                                const int blocks = (half == 1 && p.head_dim <= 192) ? 1 : 2;
                                mbar_arrive_expect_tx(bar, blocks * Cfg::kHalfBytes);
                                for (int c = 0; c < blocks; ++c)
                                    tma_load_4d(dst + c * Cfg::kHalfBytes, tm, bar, (half * 2 + c) * 64, w.kv_head, row, b);
                            } else {
                                load_tile(tm, dst, bar, w.kv_head, row, b);
                            }
                        }
                    }
                };
                auto k_entry = [&](int it) { return base + EPT * it; };
                auto v_entry = [&](int it) { return base + EPT * it + EPT / 2; };
                // The first K tile only needs a ring slot (free long before the previous item ends): it goes out first, so that
                // at the item boundary only the Q tiles are still to come (an SM takes in ~40-60 B/clk: Q + K + V of a first
                // iteration are 128 KB = ~2800 clocks after the Q buffer frees up, measured; Q alone is half of that).
                produce(&p.tm_k, w.n_max - 1, k_entry(0));
                // Q tiles of the previous item must have been consumed by its last QK^T
                FA_TRACE_EV(300);  // loader: waiting for the Q buffer
                mbar_wait(bar_q_empty, (ka & 1) ^ 1);
                FA_TRACE_EV(301);  // loader: Q buffer free
                // DECODE: tm_q's box is (64, G, 128/G, 1), so one load brings the G heads of 128/G positions
                if (lane == 0) load_tile(&p.tm_q, sQ, bar_q_full(0), w.head, w.g.q_off + w.m0, w.g.q_b);
                if (!DECODE && !SPLIT && lane == 0)
                    load_tile(&p.tm_q, sQ + Cfg::kTileBytes, bar_q_full(1), w.head, w.g.q_off + w.m0 + BM, w.g.q_b);
                if (Cfg::kHalfRing) {
                    // One K and one V tile fit: K goes one tile ahead of V, K(0) | Q | V(0) K(1) | K(2) V(1) | K(3) V(2) .. --
                    // the order in which the MMA warp frees the slots (Q K^T(it+1) is issued before P V(it)); with K(it+1)
                    // behind V(it) in this in-order queue it would wait a whole P V longer than its slot does. (V(0)'s slots
                    // are free once the previous item's last P V is done, long before Q K^T(0) frees K(1)'s.)
                    produce(&p.tm_v, w.n_max - 1, v_entry(0));
                    FA_TRACE_EV(302);  // loader: the first iteration's tiles are on their way
                    for (int it = 0; it < w.n_tiles; ++it) {
                        if (it + 1 < w.n_tiles) produce(&p.tm_k, w.n_max - 2 - it, k_entry(it + 1));
                        if (it > 0) produce(&p.tm_v, w.n_max - 1 - it, v_entry(it));
                    }
                } else {
                    produce(&p.tm_v, w.n_max - 1, v_entry(0));
                    FA_TRACE_EV(302);  // loader: Q0, K, Q1, V of the first iteration issued
                    for (int it = 1; it < w.n_tiles; ++it) {
                        produce(&p.tm_k, w.n_max - 1 - it, k_entry(it));
                        produce(&p.tm_v, w.n_max - 1 - it, v_entry(it));
                    }
                }
                ring = base + EPT * w.n_tiles;
                ++ka;
            }
            id = prefetching ? fetch_read() : total_work;
            if (lane == 0) watchdog_progress(sWatch);
        }
        watchdog_role_done(bar_done);
    } else if (warp == 12) {
        // ============================================================ MMA issuer
        reg_dec<48>();
        constexpr uint32_t idesc_qk = umma_idesc_f16(BF16, BM, BN, false, false);
        constexpr uint32_t idesc_qk_half = umma_idesc_f16(BF16, BM, BN / 2, false, false);
        constexpr bool SPLIT1 = SPLIT && FA_SPLIT_SINGLE;  // one stage, P V with N = 256 into O0|O1 (contiguous columns)
        constexpr uint32_t idesc_pv = umma_idesc_f16(BF16, BM, SPLIT1 ? 2 * DO : DO, false, true);
        constexpr uint32_t idesc_pv_half = umma_idesc_f16(BF16, BM, DO, false, true);  // half ring: 128 output columns at a time
        constexpr uint32_t idesc_pv_quarter = umma_idesc_f16(BF16, BM, 64, false, true);  // head_dim <= 192: columns [128,192)
        const uint32_t tS[2] = {tmem_base + Cfg::kTmemS0, tmem_base + Cfg::kTmemS1};
        const uint32_t tO[2] = {tmem_base + Cfg::kTmemO0, tmem_base + Cfg::kTmemO1};
        // Descriptor words (ptx_sm100.cuh): lo = addr>>4 | (LBO>>4)<<16, hi = SBO>>4 | version | swizzle.
        // Q, K: K-major, 128B swizzle, 8-row groups 1024 B apart (LBO unused = 1).
        // V: MN-major B operand: 64-column blocks kHalfBytes apart (LBO), 8-row groups 1024 B apart (SBO).
        constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
        constexpr uint32_t kLoKmajor = 1u << 16;
        constexpr uint32_t kLoVmn = (uint32_t)(Cfg::kHalfBytes >> 4) << 16;
        auto lo_addr = [](uint32_t saddr) { return (saddr & 0x3FFFFu) >> 4; };
        // The whole warp stays convergent; each issue block elects one lane (umma_issue_gen.cuh).
        auto issue_qk = [&](int s, uint32_t k_smem) {
            const uint32_t a_lo = lo_addr(sQ + (SPLIT ? 0 : s) * Cfg::kTileBytes) | kLoKmajor;
            const uint32_t b_lo = lo_addr(k_smem) | kLoKmajor;
            if constexpr (D == 256) umma_issue_qk_d256(tS[s], a_lo, b_lo, kDescHi, kDescHi, idesc_qk);
            else if constexpr (D == 128) umma_issue_qk_d128(tS[s], a_lo, b_lo, kDescHi, kDescHi, idesc_qk);
            else umma_issue_qk_d64(tS[s], a_lo, b_lo, kDescHi, kDescHi, idesc_qk);
        };
        // One half of S_s = Q_s K^T: keys [64 h, 64 h + 64) of the tile -> columns [64 h, 64 h + 64) of S_s. The left
        // half does not touch the columns P_s lives in, so it can run while the previous P_s is still being made.
        auto issue_qk_half = [&](int s, uint32_t k_smem, int h) {
            const uint32_t a_lo = lo_addr(sQ + (SPLIT ? 0 : s) * Cfg::kTileBytes) | kLoKmajor;
            const uint32_t b_lo = lo_addr(k_smem + h * 64 * 128) | kLoKmajor;  // 64 rows x 128 B further in every block
            const uint32_t d = tS[s] + h * 64;
            if constexpr (D == 256) umma_issue_qk_d256(d, a_lo, b_lo, kDescHi, kDescHi, idesc_qk_half);
            else if constexpr (D == 128) umma_issue_qk_d128(d, a_lo, b_lo, kDescHi, kDescHi, idesc_qk_half);
            else umma_issue_qk_d64(d, a_lo, b_lo, kDescHi, kDescHi, idesc_qk_half);
        };
        // head_dim 256, half ring: dims [128 h, 128 h + 128) of S = Q K^T (two swizzle blocks of Q and the K half-slot)
        const bool narrow = Cfg::kHalfRing && p.head_dim <= 192;  // dims [192,256) are padding: skip their MMAs
        auto issue_qk_dims = [&](uint32_t k_smem, int h) {
            const uint32_t a_lo = lo_addr(sQ + h * 2 * Cfg::kHalfBytes) | kLoKmajor;
            const uint32_t b_lo = lo_addr(k_smem) | kLoKmajor;
            if (h == 1 && narrow) umma_issue_qk_quarter256(tS[0], a_lo, b_lo, kDescHi, kDescHi, idesc_qk, 1u);
            else umma_issue_qk_half256(tS[0], a_lo, b_lo, kDescHi, kDescHi, idesc_qk, h ? 1u : 0u);
        };
        auto slot_addr = [&](int r) { return sKV + (r % KV) * Cfg::kSlotBytes; };
        constexpr int EPT = Cfg::kEntriesPerTile;  // ring entries per KV tile: K, V -- or K-lo, K-hi, V-lo, V-hi
        auto wait_full = [&](int r) { mbar_wait(bar_kv_full(r % KV), (r / KV) & 1); };

        const uint32_t tPs[2] = {tmem_base + Cfg::kTmemP0, tmem_base + Cfg::kTmemP1};
        int ring = 0;           // ring entries consumed by earlier items
        int ka = 0;             // items with work so far
        int kfix = 0;           // items with a ragged tail so far
        int steps[2] = {0, 0};  // softmax steps of earlier items, per stage (barrier phase bookkeeping)
        int sx_uses = 0;        // shared-S mode: Q K^T tiles written into the S buffer so far
        for (int k = 0;; ++k) {
            const Work wk = get_work(k);
            const int id = wk.id;
            if (id >= total_work) break;
            const WorkGeom w = finish_geom<DECODE, SPLIT>(p, wk.h);
            if (w.n_tiles <= 0) continue;
            // the item's last Q K^T in issue order (stage 1 is issued behind stage 0 within an iteration)
            const int last0 = w.it_hi[0] > w.it_lo[0] ? w.it_hi[0] - 1 : -1, last1 = w.it_hi[1] > w.it_lo[1] ? w.it_hi[1] - 1 : -1;
            const int last_qk_it = max(last0, last1);
            const int last_qk_s = (last1 >= 0 && last1 == last_qk_it) ? 1 : 0;
            FA_TRACE_EV(130);  // MMA: new item
            mbar_wait(bar_q_full(0), ka & 1);
            FA_TRACE_EV(131);  // MMA: Q0 landed
            // Issue order per iteration. Separate S buffers: PV0(it-1) QK0(it) PV1(it-1) QK1(it) -- tcgen05 ops execute in
            // issue order, so S_s(it) may overwrite the columns P_s(it-1) lives in without a barrier. Shared S buffer:
            // QK0(it) PV0(it-1) QK1(it) PV1(it-1) -- a Q K^T only waits until the previous S tile (the other stage's, as a
            // rule) has been copied to registers, so it runs while its own stage is still busy with the tile before.
            // Either way the first QK^T of the next item may follow the last P V of this one directly.
            for (int it = 0; it <= w.n_tiles; ++it) {
                if (!Cfg::kHalfRing && it < w.n_tiles) wait_full(ring + EPT * it);  // (half ring: each half right before its MMAs)
                bool v_ready = false;  // V(it-1) is only waited for by the first P V of the iteration (a Q K^T does not need it)
                if (!DECODE && !SPLIT && it == 0) mbar_wait(bar_q_full(1), ka & 1);
                const bool qk0_this_it = it < w.n_tiles && it >= w.it_lo[0] && it < w.it_hi[0];
                const bool qk1_this_it = it < w.n_tiles && it >= w.it_lo[1] && it < w.it_hi[1];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const bool do_pv = it > 0 && (it - 1) >= w.it_lo[s] && (it - 1) < w.it_hi[s];
                    const bool do_qk = it < w.n_tiles && it >= w.it_lo[s] && it < w.it_hi[s];
                    auto qk = [&]() {
                        if (Cfg::kSharedS) {  // the S tile written before this one sits in its owner's registers
                            if (sx_uses > 0) mbar_wait(bar_sx_free, (sx_uses - 1) & 1);
                            ++sx_uses;
                        }
                        tc_fence_after();
                        if (Cfg::kHalfRing) {
                            // the first half of the dims, its K half-slot released right behind it, then the second half
                            wait_full(ring + EPT * it);
                            tc_fence_after();
                            issue_qk_dims(slot_addr(ring + EPT * it), 0);
                            umma_commit_elect(bar_kv_empty((ring + EPT * it) % KV));
                            wait_full(ring + EPT * it + 1);
                            tc_fence_after();
                            issue_qk_dims(slot_addr(ring + EPT * it + 1), 1);
                            umma_commit_elect(bar_kv_empty((ring + EPT * it + 1) % KV));
                        } else if (Cfg::kEarlyQK && do_pv) {
                            issue_qk_half(s, slot_addr(ring + EPT * it), 1);  // left half went ahead
                        } else {
                            issue_qk(s, slot_addr(ring + EPT * it));
                        }
                        umma_commit_elect(bar_s_full(s));
                        // the item's last Q K^T: the Q buffer is free as soon as it completes (a commit at the end of the
                        // iteration would also wait for the P V issued behind it, i.e. for a whole softmax)
                        if (it == last_qk_it && s == last_qk_s) umma_commit_elect(bar_q_empty);
                        // likewise the K tile: free once the iteration's last Q K^T has read it (shared-S order only: there
                        // a P V of the previous iteration is issued BEHIND this Q K^T and would hold the slot for a softmax)
                        if (!Cfg::kHalfRing && Cfg::kSharedS && (s == 1 || !qk1_this_it)) umma_commit_elect(bar_kv_empty((ring + EPT * it) % KV));
                        FA_TRACE_EV(120 + s);  // MMA: QK_s issued
                    };
                    if (Cfg::kSharedS && do_qk) qk();
                    if (Cfg::kHalfRing && do_pv) {
                        // head_dim 256, half ring (stage 0 alone): O0 += P V[:, 0:128) from the V-lo slot, released right behind
                        // it, then O1 += P V[:, 128:256) from the V-hi slot
                        const int ve = ring + EPT * (it - 1) + 2;
                        wait_full(ve);
                        if (it == 1 && w.ragged_tail) mbar_wait(bar_vfix, kfix & 1);  // V rows past seqlen_k are zero now
                        const int j = it - 1 - w.it_lo[s];
                        const uint32_t ph = (steps[s] + j) & 1;
                        const uint32_t acc = j > 0 ? 1u : 0u;
                        const uint32_t v0 = lo_addr(slot_addr(ve)) | kLoVmn, v1 = lo_addr(slot_addr(ve + 1)) | kLoVmn;
                        mbar_wait(bar_p_full(s), ph);
                        tc_fence_after();
                        FA_TRACE_EV(100 + s);  // MMA: P_s (3/4) + O rescale observed
                        umma_issue_pv_k0_6(tO[0], tPs[s], v0, 0, kDescHi, idesc_pv_half, acc);
                        mbar_wait(bar_p_last(s), ph);
                        tc_fence_after();
                        FA_TRACE_EV(110 + s);  // MMA: last quarter of P_s observed
                        umma_issue_pv_k6_8(tO[0], tPs[s], v0, 0, kDescHi, idesc_pv_half, 1u);
                        umma_commit_elect(bar_kv_empty(ve % KV));
                        wait_full(ve + 1);
                        tc_fence_after();
                        umma_issue_pv_k0_8(tO[1], tPs[s], v1, 0, kDescHi, narrow ? idesc_pv_quarter : idesc_pv_half, acc);
                        umma_commit_elect(bar_kv_empty((ve + 1) % KV));
                        umma_commit_elect(bar_p_free(s));  // P_s may be overwritten, O_s rescaled
                        if (it == w.it_hi[s]) umma_commit_elect(bar_o_full(s));
                    } else if (do_pv) {
                        if (!v_ready) {
                            wait_full(ring + 2 * it - 1);
                            if (it == 1 && w.ragged_tail) mbar_wait(bar_vfix, kfix & 1);  // V rows past seqlen_k are zero now
                            v_ready = true;
                        }
                        const int j = it - 1 - w.it_lo[s];
                        const uint32_t ph = (steps[s] + j) & 1;
                        const uint32_t tP = tPs[s];
                        // split-D: stage s multiplies by its own 128-column half of V (two swizzle blocks further)
                        const uint32_t v_lo = lo_addr(slot_addr(ring + 2 * it - 1) + (SPLIT && !SPLIT1 ? s * 2 * Cfg::kHalfBytes : 0)) | kLoVmn;
                        if (Cfg::kEarlyQK) {
                            // S_s(j) sits in the softmax warps' registers: columns [0,64) of S_s are free, so the left
                            // half of the next S_s can be computed while P_s(j) is still being made
                            mbar_wait(bar_s_loaded(s), ph);
                            if (do_qk) {
                                tc_fence_after();
                                issue_qk_half(s, slot_addr(ring + 2 * it), 0);
                            }
                        }
                        // P_s(j) written, O_s rescaled; for j == 0 this also means the correction warps
                        // finished reading O_s of the previous item (their arrival comes after that epilogue)
                        mbar_wait(bar_p_full(s), ph);
                        tc_fence_after();
                        FA_TRACE_EV(100 + s);  // MMA: P_s (3/4) + O rescale observed
                        if (FA_SPLIT_P) {
                            umma_issue_pv_k0_6(tO[s], tP, v_lo, 0, kDescHi, idesc_pv, j > 0 ? 1u : 0u);
                            mbar_wait(bar_p_last(s), ph);
                            tc_fence_after();
                            FA_TRACE_EV(110 + s);  // MMA: last quarter of P_s observed
                            umma_issue_pv_k6_8(tO[s], tP, v_lo, 0, kDescHi, idesc_pv, 1u);
                        } else {
                            umma_issue_pv_k0_8(tO[s], tP, v_lo, 0, kDescHi, idesc_pv, j > 0 ? 1u : 0u);
                        }
                        if (Cfg::kSharedS) umma_commit_elect(bar_p_free(s));  // P_s may be overwritten, O_s rescaled
                        if (it == w.it_hi[s]) umma_commit_elect(bar_o_full(s));
                    }
                    if (!Cfg::kSharedS && do_qk) qk();
                }
                if (Cfg::kHalfRing) {
                    // (every tile of a head_dim-256 item is used by stage 0, which releases the half-slots itself; a tile
                    // nobody multiplies with would still have to land before its slots go back to the loader)
                    if (it < w.n_tiles && !qk0_this_it) {
                        wait_full(ring + EPT * it);
                        wait_full(ring + EPT * it + 1);
                        umma_commit_elect(bar_kv_empty((ring + EPT * it) % KV));
                        umma_commit_elect(bar_kv_empty((ring + EPT * it + 1) % KV));
                    }
                    const bool pv0_this_it = it > 0 && (it - 1) >= w.it_lo[0] && (it - 1) < w.it_hi[0];
                    if (it > 0 && !pv0_this_it) {
                        wait_full(ring + EPT * (it - 1) + 2);
                        wait_full(ring + EPT * (it - 1) + 3);
                        umma_commit_elect(bar_kv_empty((ring + EPT * (it - 1) + 2) % KV));
                        umma_commit_elect(bar_kv_empty((ring + EPT * (it - 1) + 3) % KV));
                    }
                } else {
                    if (it > 0) umma_commit_elect(bar_kv_empty((ring + 2 * it - 1) % KV));
                    if (it < w.n_tiles && !(Cfg::kSharedS && (qk0_this_it || qk1_this_it))) umma_commit_elect(bar_kv_empty((ring + 2 * it) % KV));
                }
                // (bar_q_empty is committed right behind the item's last Q K^T, see qk(); an item none of whose stages has a
                // tile -- possible with Sq > Sk and a window -- releases the Q buffer here)
                if (last_qk_it < 0 && it == w.n_tiles - 1) umma_commit_elect(bar_q_empty);
            }
            ring += EPT * w.n_tiles;
            steps[0] += w.it_hi[0] - w.it_lo[0];
            steps[1] += w.it_hi[1] - w.it_lo[1];
            kfix += w.ragged_tail ? 1 : 0;
            ++ka;
            if (lane == 0) watchdog_progress(sWatch);
        }
        watchdog_role_done(bar_done);
    } else if (warp < 8) {
        // ============================================================ softmax (stage = warp / 4)
        reg_inc<192>();
        const int s = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_off + (s == 0 ? Cfg::kTmemS0 : Cfg::kTmemS1);
        const uint32_t tP = tmem_base + lane_off + (s == 0 ? Cfg::kTmemP0 : Cfg::kTmemP1);
        // FEAT: S is transformed to log2 units in registers (sl2 = 1), except for soft-capping alone, where the registers
        // keep tanh(score * scale / cap) and cap * log2(e) is folded into the exponent's FMA like the plain scale
        const bool cap_only = FEAT && p.softcap > 0.f && p.alibi == nullptr;
        const float sl2 = FEAT ? (cap_only ? p.softcap * kLog2e : 1.0f) : p.scale_log2;
        int steps = 0;  // softmax steps of earlier items (barrier phase bookkeeping)
        int items = 0;  // earlier items in which this stage took part

        for (int k = 0;; ++k) {
            const Work wk = get_work(k);
            const int id = wk.id;
            if (id >= total_work) break;
            const WorkGeom w = finish_geom<DECODE, SPLIT>(p, wk.h);
            const int my_lo = s == 0 ? w.it_lo[0] : w.it_lo[1];
            const int my_n = (s == 0 ? w.it_hi[0] : w.it_hi[1]) - my_lo;
            if (my_n <= 0) continue;
            const int i_glob = DECODE ? row / G : w.m0 + (SPLIT ? 0 : s * BM) + row;  // query position inside the sequence

            // visible key range of this row: [col_lo, col_hi)
            int col_hi = w.g.seqlen_k;
            if (p.window_right >= 0) col_hi = min(col_hi, i_glob + w.off + p.window_right + 1);
            int col_lo = 0;
            if (p.window_left >= 0) col_lo = max(0, i_glob + w.off - p.window_left);
            const unsigned col_width = (unsigned)max(col_hi - col_lo, 0);

            float slope = 0.f, inv_cap = 0.f;
            if constexpr (FEAT) {
                if (p.alibi) slope = p.alibi[w.batch * p.alibi_stride_b + w.head + (DECODE ? row % G : 0)];
                if (p.softcap > 0.f) inv_cap = 1.0f / p.softcap;
            }

            float m_ref = -INFINITY;  // running reference max (raw score units; log2 units if FEAT)
            float row_sum = 0.f;

            // dropout: flat index of (this row, column 0) in the reference's numbering, and the dmask row
            uint64_t drop_row_idx = 0;
            uint16_t* dm_row = nullptr;
            if constexpr (DROPOUT) {
                drop_row_idx = (uint64_t)(w.g.q_off + i_glob) * (uint64_t)p.seqlen_k;
                if (p.dmask && i_glob < w.g.seqlen_q && !(SPLIT && s == 1))  // split-D: both stages hold the same mask
                    dm_row = p.dmask + w.o_b * p.dmask_stride_b + w.head * p.dmask_stride_h +
                             (int64_t)(w.g.q_off + i_glob) * p.dmask_stride_row;
            }

            // DECODE packs (position, head-in-group) pairs into the tile's rows: with one new token and a group of 4 only rows
            // 0..3 are real. A warp all of whose 32 rows are padding keeps every hand-shake but skips the arithmetic, which
            // leaves the sub-partitions to the one warp that has rows (its softmax is the decode chain's longest link).
            // P and O of padded rows are never stored, and the MMA keeps rows independent.
            const bool lazy_warp = DECODE && (warp & 3) * 32 >= G * w.g.seqlen_q;
            if (lazy_warp) {  // "no rescale" for both tile parities (the correction warps read sScale of every row)
                sScale[(0 * 2 + s) * BM + row] = 1.0f;
                sScale[(1 * 2 + s) * BM + row] = 1.0f;
                __syncwarp();
            }

            for (int j = 0; j < my_n; ++j) {
                const int j0 = (w.n_max - 1 - (my_lo + j)) * BN;
                mbar_wait(bar_s_full(s), (steps + j) & 1);
                if (lazy_warp) {
                    __syncwarp();
                    if (lane == 0) {
                        if (Cfg::kEarlyQK || Cfg::kSharedS) mbar_arrive(Cfg::kSharedS ? bar_sx_free : bar_s_loaded(s));
                        mbar_arrive(bar_stats(s, (steps + j) & 1));  // (sScale of padded rows is never looked at: see below)
                    }
                    if (Cfg::kSharedS && steps + j > 0) mbar_wait(bar_p_free(s), (steps + j - 1) & 1);
                    if (lane == 0) {
                        mbar_arrive(bar_p_full(s));
                        if (FA_SPLIT_P) mbar_arrive(bar_p_last(s));
                    }
                    continue;
                }
                tc_fence_after();
                FA_TRACE_EV(1);  // softmax: S observed
                float v[BN];
                const bool need_mask = (j0 + BN > col_hi) || (j0 < col_lo);
                const bool any_mask = __any_sync(0xffffffffu, need_mask);
                // FA_LD_OVERLAP: the row max of columns [0,64) runs while columns [64,128) are still on their way
                const bool split_ld = FA_LD_OVERLAP && !FEAT && !Cfg::kEarlyQK && !Cfg::kSharedS && !any_mask;
                if (split_ld) {
                    tmem_ld_2x32_wait(tS, reinterpret_cast<uint32_t*>(v));
                    tmem_ld_2x32_nowait(tS + 64, reinterpret_cast<uint32_t*>(v + 64));
                } else {
                    tmem_ld_4x32_wait(tS, reinterpret_cast<uint32_t*>(v));
                }
                FA_TRACE_EV(2);  // softmax: S in registers
                if (Cfg::kEarlyQK || Cfg::kSharedS) {  // (shared-S mode: bar_s_loaded(0) is the buffer's "free" barrier)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(Cfg::kSharedS ? bar_sx_free : bar_s_loaded(s));
                }

                if constexpr (FEAT) {
                    // reference order (include/mat_mul.h:111-117): scale, ALiBi, then softcap
                    const int rel0 = i_glob + w.off - j0;  // column of this tile that sits on the row's diagonal
                    // ALiBi alone is AFFINE in the column index on every tile that does not straddle the diagonal of any
                    // row of the warp: -slope |rel0 - c| = -/+ slope (rel0 - c). Score, scale and bias then fold into one
                    // packed FMA per pair plus one immediate-operand FMA per element for the bias (a * c + b0, c a
                    // compile-time constant): ~2.5 instructions per pair instead of ~12 (the general path below converts
                    // an integer per element). Only the diagonal tiles of a causal walk take the general path.
                    const bool affine = p.softcap <= 0.f;
                    const bool all_left = affine && __all_sync(0xffffffffu, rel0 >= BN - 1);  // every column <= diagonal
                    const bool all_right = affine && __all_sync(0xffffffffu, rel0 <= 0);      // every column >= diagonal
                    if (cap_only) {
                        // one packed multiply per pair and one MUFU.TANH per element (the general path below spends ~8
                        // instructions per element on the same thing)
                        const float c1 = p.scale * inv_cap;
#pragma unroll
                        for (int c = 0; c < BN; c += 2) {
                            mul2(v[c], v[c + 1], c1, c1);
                            v[c] = tanh_approx(v[c]);
                            v[c + 1] = tanh_approx(v[c + 1]);
                        }
                    } else if (all_left || all_right) {
                        const float a = (all_left ? slope : -slope) * kLog2e;
                        const float b0 = -a * (float)rel0;
                        const float s2 = p.scale * kLog2e;
#pragma unroll
                        for (int c = 0; c < BN; c += 2) {
                            const float bias0 = fmaf(a, (float)c, b0), bias1 = fmaf(a, (float)(c + 1), b0);
                            fma2(v[c], v[c + 1], s2, s2, bias0, bias1);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < BN; ++c) {
                            float u = v[c] * p.scale;
                            u -= slope * fabsf((float)(rel0 - c));
                            if (p.softcap > 0.f) u = p.softcap * tanh_approx(u * inv_cap);
                            v[c] = u * kLog2e;
                        }
                    }
                }
                if (any_mask) {
                    const int base = j0 - col_lo;
#pragma unroll
                    for (int c = 0; c < BN; ++c)
                        v[c] = ((unsigned)(base + c) < col_width) ? v[c] : -INFINITY;
                }

                // row max: four independent 3-input max chains
                float mx[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) mx[a] = fmaxf(v[2 * a], v[2 * a + 1]);
#pragma unroll
                for (int c = 8; c < BN / 2; c += 8) {
#pragma unroll
                    for (int a = 0; a < 4; ++a) mx[a] = fmax3(mx[a], v[c + 2 * a], v[c + 2 * a + 1]);
                }
                if (split_ld) tmem_wait_ld_x64(reinterpret_cast<uint32_t*>(v + 64));
#pragma unroll
                for (int c = BN / 2; c < BN; c += 8) {
#pragma unroll
                    for (int a = 0; a < 4; ++a) mx[a] = fmax3(mx[a], v[c + 2 * a], v[c + 2 * a + 1]);
                }
                const float m_new = fmaxf(m_ref, fmax3(fmaxf(mx[0], mx[1]), mx[2], mx[3]));
                const float m_new_safe = (m_new == -INFINITY) ? 0.f : m_new;

                float acc_scale = 1.0f;
                if (j == 0) {
                    m_ref = m_new;
                } else {
                    const float d = (m_ref - m_new_safe) * sl2;  // <= 0, -inf if nothing was visible yet
                    if (d < -kRescaleThreshold) {
                        acc_scale = ex2_approx(d);
                        m_ref = m_new;
                        if (p.dbg_counters) atomicAdd(p.dbg_counters, 1ull);
                    }
                }
                const int tpar = (steps + j) & 1;
                sScale[(tpar * 2 + s) * BM + row] = acc_scale;
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_stats(s, tpar));
                FA_TRACE_EV(3);  // softmax: row max done, stats published

                const float m_used = (m_ref == -INFINITY) ? 0.f : m_ref;
                const float neg_m = -m_used * sl2;
                float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
                for (int ch = 0; ch < BN / 32; ++ch) {
                    uint32_t pk[16];
                    uint32_t keep = 0xffffffffu;  // bit c = column ch*32+c survives dropout
                    if constexpr (DROPOUT) {
                        // counter = offset + (flat index >> 2), word = flat index & 3 (reference softmax.h:97-104)
                        const uint64_t idx0 = drop_row_idx + (uint64_t)(j0 + ch * 32);
                        const uint32_t sh = (uint32_t)idx0 & 3u;
                        const uint64_t ctr0 = p.drop_offset + (idx0 >> 2);
                        const uint32_t k0 = (uint32_t)p.drop_seed, k1 = (uint32_t)(p.drop_seed >> 32);
                        uint32_t lo = 0u, hi = 0u;
#pragma unroll
                        for (int g = 0; g < 8; ++g) lo |= philox_keep4(ctr0 + g, k0, k1, p.drop_thr) << (4 * g);
                        if (sh != 0u) hi = philox_keep4(ctr0 + 8, k0, k1, p.drop_thr);
                        keep = __funnelshift_r(lo, hi, sh);
                    }
#pragma unroll
                    for (int c = 0; c < 32; c += 2) {
                        float p0 = v[ch * 32 + c], p1 = v[ch * 32 + c + 1];
                        fma2(p0, p1, sl2, sl2, neg_m, neg_m);
                        if (FA_EMU_COUNT > 0 && ((c / 2) % FA_EMU_PERIOD) >= FA_EMU_PERIOD - FA_EMU_COUNT) {
                            ex2_emu2(p0, p1);
                        } else {
                            p0 = ex2_approx(p0);
                            p1 = ex2_approx(p1);
                        }
                        add2(sum0, sum1, p0, p1);  // the row sum uses P before dropout (reference softmax.h:94)
                        if constexpr (DROPOUT) {
                            p0 = (keep >> c) & 1u ? p0 : 0.f;
                            p1 = (keep >> (c + 1)) & 1u ? p1 : 0.f;
                        }
                        pk[c / 2] = pack2<BF16>(p0, p1);
                    }
                    if constexpr (DROPOUT) {
                        if (dm_row) {  // +1.0 kept / -1.0 dropped (reference softmax.h:116-124), columns < seqlen_k only
                            constexpr uint32_t kOne2 = BF16 ? 0x3F803F80u : 0x3C003C00u;
                            const int c_base = j0 + ch * 32;
                            uint16_t* dm = dm_row + c_base;
                            const bool vec_ok = (reinterpret_cast<uintptr_t>(dm) & 15) == 0;
#pragma unroll
                            for (int c = 0; c < 32; c += 8) {
                                uint32_t wd[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    wd[e] = kOne2 | (((~keep >> (c + 2 * e)) & 1u) << 15) | (((~keep >> (c + 2 * e + 1)) & 1u) << 31);
                                if (vec_ok && c_base + c + 8 <= w.g.seqlen_k) {
                                    *reinterpret_cast<uint4*>(dm + c) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
                                } else {
#pragma unroll
                                    for (int e = 0; e < 8; ++e)
                                        if (c_base + c + e < w.g.seqlen_k) dm[c + e] = (uint16_t)(wd[e >> 1] >> (16 * (e & 1)));
                                }
                            }
                        }
                    }
                    if (Cfg::kSharedS && ch == 0 && steps + j > 0) {
                        // This stage's softmax runs ahead of its P V: before P_s is overwritten, the P V of the stage's
                        // previous tile must have read it. Its completion also means the MMA warp consumed the previous
                        // p_full / p_last phases, so those barriers never get two phases ahead of their waiter.
                        mbar_wait(bar_p_free(s), (steps + j - 1) & 1);
                        tc_fence_after();
                        FA_TRACE_EV(7);  // softmax: previous P V of this stage complete
                    }
                    tmem_st_x16(tP + ch * 16, pk);
                    if (FA_SPLIT_P && ch == BN / 32 - 2) {  // 3/4 of P is on its way: let P V start
                        tmem_wait_st();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_p_full(s));
                        FA_TRACE_EV(4);  // softmax: 3/4 of P published
                    }
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(FA_SPLIT_P ? bar_p_last(s) : bar_p_full(s));
                FA_TRACE_EV(5);  // softmax: all of P published
                row_sum = row_sum * acc_scale + (sum0 + sum1);
            }
            // final statistics, double-buffered by item parity (the next item's may be ready before the
            // correction warps have read this one's)
            const int fb = items & 1;
            sRowSum[(fb * 2 + s) * BM + row] = row_sum;
            sRowMax[(fb * 2 + s) * BM + row] = ((m_ref == -INFINITY) ? 0.f : m_ref) * sl2;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_final(s, fb));
            steps += my_n;
            ++items;
        }
        watchdog_role_done(bar_done);
    } else if (warp < 12) {
        // ============================================================ correction + epilogue
        reg_dec<80>();
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tO[2] = {tmem_base + lane_off + Cfg::kTmemO0, tmem_base + lane_off + Cfg::kTmemO1};
        constexpr bool SPLIT1 = SPLIT && FA_SPLIT_SINGLE;
        constexpr int kStageCols = SPLIT1 ? 2 * DO : DO;  // O columns one stage accumulates (head_dim 256, single stage: O0|O1)
        uint16_t* outp = reinterpret_cast<uint16_t*>(p.out);
        int steps[2] = {0, 0};
        int items[2] = {0, 0};

        for (int k = 0;; ++k) {
            const Work wk = get_work(k);
            const int id = wk.id;
            if (id >= total_work) break;
            const WorkGeom w = finish_geom<DECODE, SPLIT>(p, wk.h);
            if (w.skip) continue;
            if (w.n_tiles <= 0) {
                // No visible key for any row of this block: out = 0, lse = sentinel (reference
                // kernel/fused_mha_forward_varlen.cu:100-111). A decode split with no tile writes an
                // ignorable partial (lse = -inf).
                const int rows = (w.m_end - w.m0) * G;  // packed rows: position-major, head-in-group minor
                const int t = (warp - 8) * 32 + lane;
                if (partial_out) {
                    for (int r = t; r < rows; r += 128)
                        p.lse_partial[(((int64_t)w.split * num_batch + w.batch) * p.num_heads + w.head + r % G) *
                                          w.g.seqlen_q + r / G] = -INFINITY;
                    continue;
                }
                const int hd8 = p.head_dim / 8;
                for (int idx = t; idx < rows * hd8; idx += 128) {
                    const int r = idx / hd8, c = idx % hd8;
                    uint16_t* dst = outp + w.o_b * p.o_stride_b + (int64_t)(w.g.q_off + w.m0 + r / G) * p.o_stride_s +
                                    (w.head + r % G) * p.o_stride_h + c * 8;
                    *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
                }
                for (int r = t; r < rows; r += 128)
                    p.lse[w.o_b * p.lse_stride_b + (w.head + r % G) * p.lse_stride_h + w.g.q_off + w.m0 + r / G] = kNegSentinel;
                continue;
            }
            for (int it = 0; it < w.n_tiles; ++it) {
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if (it < w.it_lo[s] || it >= w.it_hi[s]) continue;
                    const int j = it - w.it_lo[s];
                    const int tpar = (steps[s] + j) & 1;
                    mbar_wait(bar_stats(s, tpar), ((steps[s] + j) >> 1) & 1);
                    FA_TRACE_EV(200 + s);  // correction: stats observed
                    const float sc = sScale[(tpar * 2 + s) * BM + row];
                    if (Cfg::kSharedS && steps[s] + j > 0) {
                        // Shared-S mode: the softmax of a stage publishes the stats of tile j while P V_s(j-1) may not even
                        // have been issued. Each correction warp therefore waits for P V_s(j-1) to COMPLETE before it
                        // touches O_s or arrives on p_full for tile j: (a) O_s is not rescaled under a running MMA, and
                        // (b) p_full has collected all eight arrivals of tile j-1 -- otherwise a correction warp that is a
                        // tile ahead of a sibling would fill the sibling's slot in the previous phase (found as a hang
                        // under ncu's SASS-patching passes, which slow the warps of a CTA very unevenly).
                        mbar_wait(bar_p_free(s), (steps[s] + j - 1) & 1);
                    }
                    if (j > 0 && __any_sync(0xffffffffu, sc != 1.0f)) {
                        if (p.dbg_counters && lane == 0) atomicAdd(p.dbg_counters + 1, 1ull);
                        tc_fence_after();
#pragma unroll
                        for (int c = 0; c < kStageCols / 32; ++c) {
                            float o[32];
                            tmem_ld_x32_wait(tO[s] + c * 32, reinterpret_cast<uint32_t*>(o));
#pragma unroll
                            for (int e = 0; e < 32; ++e) o[e] *= sc;
                            tmem_st_x32(tO[s] + c * 32, reinterpret_cast<uint32_t*>(o));
                        }
                        tmem_wait_st();
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_p_full(s));
                }
            }
            // epilogue: out = O / l, lse = m + ln(l)   (reference kernel/fused_mha_forward.cu:215-223)
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (DECODE && !SPLIT && s == 1) continue;  // packed-row mode has a single tile
                if (SPLIT1 && s == 1) continue;             // stage 0 writes all 256 columns
                const int i_glob = DECODE ? row / G : w.m0 + (SPLIT ? 0 : s * BM) + row;
                const int h_row = DECODE ? w.head + row % G : w.head;
                const bool valid = i_glob < w.g.seqlen_q;
                constexpr int kColStep = (SPLIT && !SPLIT1) ? DO : 0;  // two-stage split-D: stage s owns output columns [s*128, s*128+128)
                const int col0 = s * kColStep;
                const bool own_lse = !(SPLIT && s == 1);
                uint16_t* dst = outp + w.o_b * p.o_stride_b + (int64_t)(w.g.q_off + i_glob) * p.o_stride_s +
                                h_row * p.o_stride_h + col0;
                float* lse_dst = p.lse + w.o_b * p.lse_stride_b + h_row * p.lse_stride_h + w.g.q_off + i_glob;
                if (w.it_hi[s] <= w.it_lo[s]) {  // this stage saw no KV tile: no key is visible to its rows
                    if (valid) {
#pragma unroll
                        for (int c = 0; c < kStageCols; c += 8)
                            if (col0 + c < p.head_dim) *reinterpret_cast<uint4*>(dst + c) = make_uint4(0, 0, 0, 0);
                        if (own_lse) *lse_dst = kNegSentinel;
                    }
                    continue;
                }
                const int fb = items[s] & 1;
                mbar_wait(bar_final(s, fb), (items[s] >> 1) & 1);
                const float l = sRowSum[(fb * 2 + s) * BM + row];
                const float mx = sRowMax[(fb * 2 + s) * BM + row];
                mbar_wait(bar_o_full(s), items[s] & 1);
                tc_fence_after();
                FA_TRACE_EV(210 + s);  // correction: final O observed
                const float inv = l > 0.f ? (DROPOUT ? p.rp_dropout : 1.0f) / l : 0.f;
                const bool wide_ok = __all_sync(0xffffffffu, (reinterpret_cast<uintptr_t>(dst) & 31) == 0);
                if (partial_out) {
                    // split-KV partial: normalised fp32 O and this split's LSE; fa_combine_kernel merges them
                    const int64_t prow = (((int64_t)w.split * num_batch + w.batch) * p.num_heads + h_row) * w.g.seqlen_q + i_glob;
#pragma unroll
                    for (int c = 0; c < kStageCols / 32; ++c) {
                        float o[32];
                        tmem_ld_x32_wait(tO[s] + c * 32, reinterpret_cast<uint32_t*>(o));
                        if (valid) {
#pragma unroll
                            for (int e = 0; e < 32; e += 4)
                                *reinterpret_cast<float4*>(p.o_partial + prow * D + col0 + c * 32 + e) =
                                    make_float4(o[e] * inv, o[e + 1] * inv, o[e + 2] * inv, o[e + 3] * inv);
                        }
                    }
                    if (valid && own_lse) p.lse_partial[prow] = l > 0.f ? (mx + lg2_approx(l)) * kLn2 : -INFINITY;
                } else {
#pragma unroll
                    for (int c = 0; c < kStageCols / 32; ++c) {
                        if (col0 + c * 32 >= p.head_dim) break;  // columns [head_dim, D) are the tile's zero padding
                        float o[32];
                        tmem_ld_x32_wait(tO[s] + c * 32, reinterpret_cast<uint32_t*>(o));
                        if (valid) {
                            uint32_t pk[16];
#pragma unroll
                            for (int e = 0; e < 32; e += 2) pk[e / 2] = pack2<BF16>(o[e] * inv, o[e + 1] * inv);
                            if (wide_ok && col0 + c * 32 + 32 <= p.head_dim) {  // 2 x 32 B per 32 columns instead of 4 x 16 B
                                st_global_v8(dst + c * 32, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
                                st_global_v8(dst + c * 32 + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
                            } else {
#pragma unroll
                                for (int e = 0; e < 16; e += 4)
                                    if (col0 + c * 32 + 2 * e < p.head_dim)
                                        *reinterpret_cast<uint4*>(dst + c * 32 + 2 * e) = make_uint4(pk[e], pk[e + 1], pk[e + 2], pk[e + 3]);
                            }
                        }
                    }
                    if (valid && own_lse) *lse_dst = l > 0.f ? (mx + lg2_approx(l)) * kLn2 : kNegSentinel;
                }
                steps[s] += w.it_hi[s] - w.it_lo[s];
                ++items[s];
                FA_TRACE_EV(220 + s);  // correction: epilogue of stage s stored
            }
        }
        watchdog_role_done(bar_done);
    } else if (warp == 14) {
        // ============================================================ V sanitiser
        // P is exactly 0 for key columns past seqlen_k, but 0 * NaN = NaN: a KV cache is allowed to hold
        // uninitialised memory beyond its valid length, and TMA cannot clip rows inside a tile. So for the one
        // ragged tile of an item (always the first one processed) this warp zeroes the smem rows of V past
        // seqlen_k before the MMA warp may read them. (K needs nothing: masked scores are replaced, not scaled.)
        reg_dec<48>();
        int ring = 0;
        for (int k = 0;; ++k) {
            const Work wk = get_work(k);
            const int id = wk.id;
            if (id >= total_work) break;
            const WorkGeom w = finish_geom<DECODE, SPLIT>(p, wk.h);
            if (w.n_tiles <= 0) continue;
            if (w.ragged_tail) {
                const int v_entry = ring + Cfg::kEntriesPerTile / 2;
                if (Cfg::kHalfRing) mbar_wait(bar_kv_full((v_entry + 1) % KV), ((v_entry + 1) / KV) & 1);  // both halves of V
                const int valid = w.g.seqlen_k - (w.n_max - 1) * BN;  // rows of the tail tile that hold keys
                mbar_wait(bar_kv_full(v_entry % KV), (v_entry / KV) & 1);
                const uint32_t v_smem = sKV + (v_entry % KV) * Cfg::kSlotBytes;  // (half ring: V-lo | V-hi are adjacent)
                // rows are 128 B long inside each 64-column block; the swizzle only permutes 16 B chunks in a row
                for (int idx = lane; idx < (BN - valid) * 8 * (D / 64); idx += 32) {
                    const int c16 = idx & 7, r = valid + ((idx >> 3) % (BN - valid)), blk = (idx >> 3) / (BN - valid);
                    st_shared_v4(v_smem + blk * Cfg::kHalfBytes + r * 128 + c16 * 16, 0u, 0u, 0u, 0u);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_vfix);
            }
            ring += Cfg::kEntriesPerTile * w.n_tiles;
        }
        watchdog_role_done(bar_done);
    } else {
        reg_dec<48>();  // warp 15: watchdog (ptx_sm100.cuh) -- traps the kernel if this CTA stops making progress
        watchdog_run(sWatch, bar_done);
    }

    // ------------------------------------------------------------------ teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 12) tmem_dealloc<512>(tmem_base);
}

// Merge the split-KV partials of the decode path: out = sum_i w_i O_i / sum_i w_i with
// w_i = exp(lse_i - max lse), lse = max + ln(sum w_i). One warp per (batch, head, position) row.
// Latency, not bandwidth, is what this kernel costs (a B=1 decode step has 32 rows): lane i fetches the LSE of
// split i (and i + 32), so all of them are in flight at once, the max and the weight sum are warp reductions, and the
// partial rows are fetched four splits at a time before any of them is used.
template <int D, bool BF16>
__global__ void fa_combine_kernel(const float* __restrict__ o_partial, const float* __restrict__ lse_partial,
                                  uint16_t* __restrict__ out, float* __restrict__ lse, int num_splits, int batch,
                                  int heads, int seqlen_q, int64_t o_stride_b, int64_t o_stride_s,
                                  int64_t o_stride_h, int head_dim) {
    const int lane = threadIdx.x & 31;
    const int64_t rows = (int64_t)batch * heads * seqlen_q;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    pdl_wait_primary();  // launched with programmatic stream serialisation behind the decode kernel
    if (row >= rows) return;
    const int pos = row % seqlen_q, h = (row / seqlen_q) % heads, b = row / ((int64_t)seqlen_q * heads);
    // num_splits <= 64 (decode_num_splits)
    const float l0 = lane < num_splits ? lse_partial[(int64_t)lane * rows + row] : -INFINITY;
    const float l1 = lane + 32 < num_splits ? lse_partial[(int64_t)(lane + 32) * rows + row] : -INFINITY;
    float mx = fmaxf(l0, l1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float w0 = (mx > -INFINITY && l0 > -INFINITY) ? __expf(l0 - mx) : 0.f;
    const float w1 = (mx > -INFINITY && l1 > -INFINITY) ? __expf(l1 - mx) : 0.f;
    float wsum = w0 + w1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    constexpr int E = D / 32;  // elements per lane
    float acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;
    const float* src0 = o_partial + row * D + lane * E;
    for (int i0 = 0; i0 < num_splits; i0 += 4) {
        float part[4][E];
        float wi[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u;
            const float wa = __shfl_sync(0xffffffffu, w0, i & 31), wb = __shfl_sync(0xffffffffu, w1, i & 31);
            wi[u] = i < num_splits ? (i < 32 ? wa : wb) : 0.f;
            if (wi[u] > 0.f) {
                const float* src = src0 + (int64_t)i * rows * D;
                if constexpr (E >= 4) {
#pragma unroll
                    for (int e = 0; e < E; e += 4) {
                        const float4 t = *reinterpret_cast<const float4*>(src + e);
                        part[u][e] = t.x; part[u][e + 1] = t.y; part[u][e + 2] = t.z; part[u][e + 3] = t.w;
                    }
                } else {
                    const float2 t = *reinterpret_cast<const float2*>(src);
                    part[u][0] = t.x; part[u][1] = t.y;
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) part[u][e] = 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] = fmaf(wi[u], part[u][e], acc[e]);
        }
    }
    const float inv = wsum > 0.f ? 1.0f / wsum : 0.f;
    uint16_t* dst = out + b * o_stride_b + (int64_t)pos * o_stride_s + h * o_stride_h + lane * E;
#pragma unroll
    for (int e = 0; e < E; e += 2)
        if (lane * E + e < head_dim) *reinterpret_cast<uint32_t*>(dst + e) = pack2<BF16>(acc[e] * inv, acc[e + 1] * inv);
    if (lane == 0) lse[((int64_t)b * heads + h) * seqlen_q + pos] = wsum > 0.f ? mx + __logf(wsum) : kNegSentinel;
}

}  // namespace fa
