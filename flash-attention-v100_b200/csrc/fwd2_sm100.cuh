// Forward v2 for head_dim 128 (plain softmax: no ALiBi / softcap / dropout / decode) -- the kernel behind BASELINE
// configs 2, 3 and 5. Same semantics and parameter block as fwd_sm100.cuh (reference kernel/fused_mha_forward.cu:25-224,
// kernel/fused_mha_forward_varlen.cu:25-275, kernel/fused_mha_forward_kvcache.cu:24-295 for Sq > 1); what changes is the
// schedule. Round 1's kernel runs two query tiles per CTA, each with ONE score buffer that its probabilities
// overwrite, so per tile "softmax -> P V -> next Q K^T -> next softmax" is a serial chain and the softmax warps idle
// ~30 % of the time waiting for their next S (profiles/timeline_r01_full_4096_1thread_per_row.txt). Here:
//
//   * one CTA = ONE 128-row query tile, ONE accumulator O, THREE score buffers S0 S1 S2 in TMEM
//     (3 x 128 + 128 = 512 columns; P_i aliases columns [64,128) of S_i as before);
//   * the MMA warp runs two tiles ahead: ... Q K^T(t+2) -> S_{(t+2)%3}, then O += P(t) V(t) ... ; S_{(t+2)%3} was
//     released when P V(t-1) was issued one step earlier, so Q K^T never waits for a softmax;
//   * two softmax groups (warps 0-3 / 4-7, thread == row in both) take alternate tiles, so each has two tiles'
//     worth of tensor time per tile of its own and its next S is always waiting for it. The running row maximum is
//     handed from group to group through shared memory (one value + one mbarrier arrival per warp and tile); each
//     group keeps its own partial row sum relative to the maximum it last used and the epilogue merges the two;
//   * the tile stream runs straight across work items: the first two Q K^T of the next item are issued during the
//     last two steps of the current one (Q is double buffered), and the three S buffers let the softmax run ahead
//     while the epilogue of the previous item drains O.
//
// K/V are then streamed once per 128 query rows instead of once per 256 -- 64 B/clk per SM from L2 against the ~42 the
// L2 delivers across 148 SMs -- so the kernel runs as CTA PAIRS (CL = 2, thread-block clusters of two): the pair
// takes the two 128-row tiles of one 256-row block of a head, walks the union of their KV ranges, and each CTA
// TMA-loads HALF of every K/V tile with .multicast::cluster into both CTAs' shared memory. A ring slot is released
// to the loaders only when both CTAs are done with it (tcgen05.commit multicast onto both CTAs' barriers; a CTA that
// skips a tile of the union -- above its causal diagonal / left of its window -- releases it with plain arrivals
// as soon as it has landed).
// The leader CTA draws the work ids and posts them into the peer's inbox through distributed shared memory.
// CL = 1 is the same kernel without a partner (every tile loaded by the CTA itself): L2-bound on large shapes,
// kept for A/B measurements and for devices whose SM count is odd.
#pragma once
#include "fwd_sm100.cuh"

namespace fa {

struct Fwd2Config {
    static constexpr int kD = 128, kBlockM = 128, kBlockN = 128;
    static constexpr int kTileBytes = kBlockN * kD * 2;
    static constexpr int kHalfBytes = kBlockN * 128;
#ifndef FA_FWD2_QBUFS
#define FA_FWD2_QBUFS 2
#endif
#ifndef FA_FWD2_KV
#define FA_FWD2_KV 4
#endif
    static constexpr int kQBufs = FA_FWD2_QBUFS;   // 2: the next item's Q K^T start before this item ends; 1: no run-ahead across items
    static constexpr int kKvStages = FA_FWD2_KV;
    static constexpr int kSmemQ = kQBufs * kTileBytes;
    static constexpr int kSmemKV = kKvStages * kTileBytes;
    // barrier table
    static constexpr int kBarQFull = 0;                      // [2]        loader -> MMA
    static constexpr int kBarQEmpty = 2;                     // [2]        MMA -> loader
    static constexpr int kBarKvFull = 4;                     // [KV]
    static constexpr int kBarKvEmpty = kBarKvFull + kKvStages;  // [KV]    count CL: the MMA warp of every CTA of the pair
    static constexpr int kBarSFull = kBarKvEmpty + kKvStages;  // [3]      MMA -> softmax: S slot written
    static constexpr int kBarPFull = kBarSFull + 3;          // [3]        softmax (4) + correction (4) -> MMA
    static constexpr int kBarStats = kBarPFull + 3;          // [3][4]     softmax warp -> correction warp (same rows)
    static constexpr int kBarMax = kBarStats + 12;           // [4][4]     softmax warp -> its twin in the other group
    static constexpr int kBarPvDone = kBarMax + 16;          //            MMA -> correction: P V(t) complete
    static constexpr int kBarFinal = kBarPvDone + 1;         // [2][2][4]  softmax warp -> correction warp: l, m
    static constexpr int kBarSchedFull = kBarFinal + 16;     // [2]
    static constexpr int kBarSchedEmpty = kBarSchedFull + 2; // [2]
    static constexpr int kBarVfix = kBarSchedEmpty + 2;
    static constexpr int kBarInboxFull = kBarVfix + 1;       // [2]        leader's loader -> peer's loader (work id posted)
    static constexpr int kBarInboxEmpty = kBarInboxFull + 2; // [2]        peer's loader -> leader's loader (work id read)
    static constexpr int kBarDone = kBarInboxEmpty + 2;      //            role warps -> watchdog
    static constexpr int kNumBars = kBarDone + 1;
    static constexpr int kOffBars = kSmemQ + kSmemKV;
    static constexpr int kOffTmemPtr = kOffBars + 8 * kNumBars;
    static constexpr int kOffScale = (kOffTmemPtr + 16 + 15) & ~15;  // [3 slots][128]   O rescale factor of a tile
    static constexpr int kOffMref = kOffScale + 3 * 128 * 4;         // [4][128]         running reference maximum after tile t (t & 3)
    static constexpr int kOffFinalL = kOffMref + 4 * 128 * 4;        // [group][parity][128]
    static constexpr int kOffFinalM = kOffFinalL + 4 * 128 * 4;
    static constexpr int kOffSched = kOffFinalM + 4 * 128 * 4;       // int[2] work ids, then the watchdog's two words
    static constexpr int kOffInbox = kOffSched + 16;                 // int[2] work ids posted by the leader CTA
    static constexpr int kOffWaitDbg = kOffInbox + 16;               // uint32[16]: line each warp waits at (FA_WAIT_DEBUG)
    static constexpr int kSmemUsed = kOffWaitDbg + 64;
    static constexpr int kSmemBytes = kSmemUsed + 1024;
    static constexpr int kTmemO = 384, kTmemPOff = 64;
};

// Ring entries of one item with n KV tiles (the union of the pair's ranges), in the order the loaders issue them and
// the MMA warps first need them: K0 K1 (K2 V0) (K3 V1) ... (K(n-1) V(n-3)) V(n-2) V(n-1).
FA_DEVICE int fwd2_entry_k(int u) { return u < 2 ? u : 2 * u - 2; }
FA_DEVICE int fwd2_entry_v(int u, int n) { return min(2 * u + 3, n + u); }

// One CTA's view of a work item: the union tile walk (n steps; step u = KV tile n_max-1-u) and the steps this CTA
// takes part in, [lo, hi). CL == 1: the item is one 128-row block and the CTA takes every step. CL == 2: the item
// is a 256-row block (the round-1 kernel's item, so the work numbering is shared with it); CTA `rank` owns rows
// m0 + 128 rank .. and its range is what that kernel calls stage `rank`'s.
struct Item2 {
    WorkGeom w;
    int n, lo, hi;
    int m0;  // first query position of this CTA's tile
};
template <int CL>
FA_DEVICE Item2 fwd2_item(const FwdKernelParams& p, int id, int rank) {
    Item2 it;
    if constexpr (CL == 2) {
        it.w = work_geom<false, false>(p, id);
        it.lo = rank ? it.w.it_lo[1] : it.w.it_lo[0];
        it.hi = rank ? it.w.it_hi[1] : it.w.it_hi[0];
        it.m0 = it.w.m0 + rank * 128;
    } else {
        it.w = work_geom<false, true>(p, id);
        it.lo = it.w.it_lo[0];
        it.hi = it.w.it_hi[0];
        it.m0 = it.w.m0;
    }
    it.n = it.w.n_tiles;
    if (it.n <= 0) it.lo = it.hi = 0;
    return it;
}

template <bool BF16, int CL>
__global__ void __launch_bounds__(512, 1)
fa_fwd2_sm100_kernel(const __grid_constant__ FwdKernelParams p) {
    using Cfg = Fwd2Config;
    constexpr int BM = Cfg::kBlockM, BN = Cfg::kBlockN, D = Cfg::kD;
    constexpr int KV = Cfg::kKvStages;
    static_assert(CL == 1 || CL == 2, "a CTA runs alone or as one of a pair");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_raw_u32 = smem_u32(smem_raw);
    const uint32_t sbase = (smem_raw_u32 + 1023u) & ~1023u;  // the same offset in both CTAs of a pair
    uint8_t* sgen = smem_raw + (sbase - smem_raw_u32);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = CL == 2 ? (int)cluster_ctarank() : 0;
    const uint32_t peer = (uint32_t)(rank ^ 1);
    FA_TRACE_DECL;
    const int total_work = p.num_m_blocks * p.num_bh;

    const uint32_t sQ = sbase;
    const uint32_t sKV = sbase + Cfg::kSmemQ;
    const uint32_t bars = sbase + Cfg::kOffBars;
    auto bar = [&](int i) { return bars + 8 * i; };
    auto bar_q_full = [&](int b) { return bar(Cfg::kBarQFull + b); };
    auto bar_q_empty = [&](int b) { return bar(Cfg::kBarQEmpty + b); };
    auto bar_kv_full = [&](int i) { return bar(Cfg::kBarKvFull + i); };
    auto bar_kv_empty = [&](int i) { return bar(Cfg::kBarKvEmpty + i); };
    auto bar_s_full = [&](int s) { return bar(Cfg::kBarSFull + s); };
    auto bar_p_full = [&](int s) { return bar(Cfg::kBarPFull + s); };
    auto bar_stats = [&](int s, int w) { return bar(Cfg::kBarStats + s * 4 + w); };
    auto bar_max = [&](int t4, int w) { return bar(Cfg::kBarMax + t4 * 4 + w); };
    const uint32_t bar_pv_done = bar(Cfg::kBarPvDone);
    auto bar_final = [&](int grp, int fb, int w) { return bar(Cfg::kBarFinal + (grp * 2 + fb) * 4 + w); };
    auto bar_sched_full = [&](int b) { return bar(Cfg::kBarSchedFull + b); };
    auto bar_sched_empty = [&](int b) { return bar(Cfg::kBarSchedEmpty + b); };
    const uint32_t bar_vfix = bar(Cfg::kBarVfix);
    auto bar_inbox_full = [&](int b) { return bar(Cfg::kBarInboxFull + b); };
    auto bar_inbox_empty = [&](int b) { return bar(Cfg::kBarInboxEmpty + b); };
    const uint32_t bar_done = bar(Cfg::kBarDone);
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(sgen + Cfg::kOffTmemPtr);
    float* sScale = reinterpret_cast<float*>(sgen + Cfg::kOffScale);
    float* sMref = reinterpret_cast<float*>(sgen + Cfg::kOffMref);
    float* sFinalL = reinterpret_cast<float*>(sgen + Cfg::kOffFinalL);
    float* sFinalM = reinterpret_cast<float*>(sgen + Cfg::kOffFinalM);
    volatile int* sSched = reinterpret_cast<volatile int*>(sgen + Cfg::kOffSched);
    volatile uint32_t* sWatch = reinterpret_cast<volatile uint32_t*>(sgen + Cfg::kOffSched + 8);  // progress, warps done
    volatile int* sInbox = reinterpret_cast<volatile int*>(sgen + Cfg::kOffInbox);
    volatile uint32_t* sWaitDbg = reinterpret_cast<volatile uint32_t*>(sgen + Cfg::kOffWaitDbg);
#define FA_WAIT(bar_, parity_) do { FA_WAIT_MARK(sWaitDbg); mbar_wait(bar_, parity_); } while (0)
#define FA_WAIT_CL(bar_, parity_) do { FA_WAIT_MARK(sWaitDbg); mbar_wait_cluster(bar_, parity_); } while (0)

    if (warp == 13 && lane == 0) {
        sWatch[0] = 0;
        sWatch[1] = 0;
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_q_full(b), 1);
            mbar_init(bar_q_empty(b), 1);
            mbar_init(bar_sched_full(b), 1);
            mbar_init(bar_sched_empty(b), 14);  // MMA warp + 8 softmax + 4 correction warps + V sanitiser
            mbar_init(bar_inbox_full(b), 1);
            mbar_init(bar_inbox_empty(b), 1);
        }
        for (int i = 0; i < KV; ++i) {
            mbar_init(bar_kv_full(i), 1);
            mbar_init(bar_kv_empty(i), CL);
        }
        for (int s = 0; s < 3; ++s) {
            mbar_init(bar_s_full(s), 1);
            mbar_init(bar_p_full(s), 8);  // 4 softmax warps of the tile's group + 4 correction warps
            for (int w = 0; w < 4; ++w) mbar_init(bar_stats(s, w), 1);
        }
        for (int t = 0; t < 4; ++t)
            for (int w = 0; w < 4; ++w) mbar_init(bar_max(t, w), 1);
        for (int i = 0; i < 16; ++i) mbar_init(bar(Cfg::kBarFinal + i), 1);
        mbar_init(bar_pv_done, 1);
        mbar_init(bar_vfix, 1);
        mbar_init(bar_done, 15);
        mbar_fence_init();
        tma_prefetch_desc(&p.tm_q);
        tma_prefetch_desc(&p.tm_k);
        tma_prefetch_desc(&p.tm_v);
    }
    if (warp == 12) tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)));
    tc_fence_before();
    __syncthreads();
    if constexpr (CL == 2) cluster_sync_all();  // the peer's barriers exist before anything is multicast onto them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // k-th work id of this CTA (>= total_work: no more work). get_work consumes the mailbox slot, peek_work does not.
    auto peek_work = [&](int k) -> int {
        FA_WAIT(bar_sched_full(k & 1), (k >> 1) & 1);
        return sSched[k & 1];
    };
    auto get_work = [&](int k) -> int {
        const int id = peek_work(k);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sched_empty(k & 1));
        return id;
    };

    if (warp == 13) {
        // ============================================================ TMA producer + tile scheduler
        reg_dec<48>();
        auto load_q = [&](uint32_t dst, uint32_t br, int h, int row, int b) {
            mbar_arrive_expect_tx(br, Cfg::kTileBytes);
            tma_load_4d(dst, &p.tm_q, br, 0, h, row, b);
            tma_load_4d(dst + Cfg::kHalfBytes, &p.tm_q, br, 64, h, row, b);
        };
        // A K/V tile is two 64-column halves; in a pair each CTA fetches one half for both (multicast), so each
        // CTA's barrier still expects the whole tile.
        auto load_kv = [&](const CUtensorMap* tm, uint32_t dst, uint32_t br, int h, int row, int b) {
            mbar_arrive_expect_tx(br, Cfg::kTileBytes);
            if constexpr (CL == 2) {
                tma_load_4d_mc(dst + rank * Cfg::kHalfBytes, tm, br, rank * 64, h, row, b, (uint16_t)0x3);
            } else {
                tma_load_4d(dst, tm, br, 0, h, row, b);
                tma_load_4d(dst + Cfg::kHalfBytes, tm, br, 64, h, row, b);
            }
        };
        int ring = 0, ka = 0;
        const int num_units = (int)gridDim.x / CL;  // CTAs (or pairs) drawing work
        int id = (int)blockIdx.x / CL;
        bool more = p.sched != nullptr;
        auto fetch = [&]() -> int {
            if (!more) return total_work;
            int nid = 0;
            if (lane == 0) nid = atomicAdd(p.sched, 1) + num_units;
            nid = __shfl_sync(0xffffffffu, nid, 0);
            if (nid >= total_work) {
                more = false;
                nid = total_work;
            }
            return nid;
        };
        const bool draws = (CL == 1) || rank == 0;  // the leader of a pair draws the ids and posts them to its peer
        for (int k = 0;; ++k) {
            if (draws) {
                while (id < total_work && fwd2_item<CL>(p, id, rank).w.skip) id = fetch();
                if constexpr (CL == 2) {
                    if (k >= 2) FA_WAIT_CL(bar_inbox_empty(k & 1), ((k >> 1) - 1) & 1);  // the peer has read slot k & 1
                    if (lane == 0) {
                        st_cluster_u32(mapa_u32(smem_u32(const_cast<int*>(sInbox + (k & 1))), peer), (uint32_t)id);
                        mbar_arrive_cluster(mapa_u32(bar_inbox_full(k & 1), peer));
                    }
                    __syncwarp();
                }
            } else {
                FA_WAIT_CL(bar_inbox_full(k & 1), (k >> 1) & 1);
                id = sInbox[k & 1];
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mapa_u32(bar_inbox_empty(k & 1), peer));
            }
            if (k >= 2) FA_WAIT(bar_sched_empty(k & 1), ((k >> 1) - 1) & 1);
            if (lane == 0) {
                sSched[k & 1] = id;
                mbar_arrive(bar_sched_full(k & 1));
                watchdog_progress(sWatch);
            }
            __syncwarp();
            if (id >= total_work) break;
            const Item2 it = fwd2_item<CL>(p, id, rank);
            const WorkGeom& w = it.w;
            int next_id = total_work;
            if (draws) next_id = fetch();  // early: the atomic's latency hides behind the loads
            if (it.n > 0) {
                const int n = it.n;
                auto produce = [&](const CUtensorMap* tm, int u) {
                    const int slot = ring % KV;
                    const uint32_t parity = ((ring / KV) & 1) ^ 1;
                    FA_WAIT(bar_kv_empty(slot), parity);
                    FA_TRACE_EV(310);
                    if (lane == 0) {
                        const int r = (w.n_max - 1 - u) * BN;
                        int row, b;
                        if (p.block_table) {
                            const int page = r / p.page_size;
                            b = p.block_table[(int64_t)w.batch * p.block_table_stride + page];
                            row = r - page * p.page_size;
                        } else {
                            b = w.g.k_b;
                            row = w.g.k_off + r;
                        }
                        load_kv(tm, sKV + slot * Cfg::kTileBytes, bar_kv_full(slot), w.kv_head, row, b);
                    }
                    ++ring;
                };
                if (it.hi > it.lo) {  // this CTA has rows and keys to work on: its Q tile
                    const int qb = ka % Cfg::kQBufs;
                    FA_WAIT(bar_q_empty(qb), ((ka / Cfg::kQBufs) & 1) ^ 1);  // the previous user of this buffer has issued its last Q K^T
                    if (lane == 0) load_q(sQ + qb * Cfg::kTileBytes, bar_q_full(qb), w.head, w.g.q_off + it.m0, w.g.q_b);
                    ++ka;
                }
                produce(&p.tm_k, 0);
                if (n > 1) produce(&p.tm_k, 1);
                for (int u = 0; u < n; ++u) {
                    if (u + 2 < n) produce(&p.tm_k, u + 2);
                    produce(&p.tm_v, u);
                }
            }
            id = next_id;
        }
        watchdog_role_done(bar_done);
    } else if (warp == 12) {
        // ============================================================ MMA issuer
        reg_dec<48>();
        constexpr uint32_t idesc_qk = umma_idesc_f16(BF16, BM, BN, false, false);
        constexpr uint32_t idesc_pv = umma_idesc_f16(BF16, BM, D, false, true);
        constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
        constexpr uint32_t kLoKmajor = 1u << 16;
        constexpr uint32_t kLoVmn = (uint32_t)(Cfg::kHalfBytes >> 4) << 16;
        auto lo_addr = [](uint32_t saddr) { return (saddr & 0x3FFFFu) >> 4; };
        const uint32_t tO = tmem_base + Cfg::kTmemO;
        auto slot_addr = [&](int e) { return sKV + (e % KV) * Cfg::kTileBytes; };
        auto wait_full = [&](int e) { FA_WAIT(bar_kv_full(e % KV), (e / KV) & 1); };
        // A ring slot goes back to the loaders of BOTH CTAs once every MMA issued so far by this CTA has completed.
        // (The multicast commit costs the issuing warp ~270 cycles per tile; releasing locally and letting the two
        // loader warps forward the releases to each other was measured far slower still -- 603 vs 840 TFLOP/s on
        // config 2 -- because the round trip then sits in the slot-recycling path of a 4-slot ring.)
        auto commit_slot = [&](int e) {
            if constexpr (CL == 2) umma_commit_elect_mc(bar_kv_empty(e % KV), (uint16_t)0x3);
            else umma_commit_elect(bar_kv_empty(e % KV));
        };
        // ... or as soon as it has landed, for a tile of the union this CTA takes no part in. (Waiting for the load
        // keeps this CTA's arrivals in step with the ring: entry e + KV cannot be loaded before both CTAs released
        // entry e, so an arrival for it can never be counted into the phase of entry e.)
        auto release_slot = [&](int e) {
            wait_full(e);
            if (lane == 0) {
                mbar_arrive(bar_kv_empty(e % KV));
                if constexpr (CL == 2) mbar_arrive_cluster(mapa_u32(bar_kv_empty(e % KV), peer));
            }
            __syncwarp();
        };
        // S_{gt % 3} = Q[qb] K(e)^T; releases the K slot, announces S, and the Q buffer after the item's last one
        auto issue_qk = [&](int gt, int qb, int e, bool last_of_item) {
            wait_full(e);
            FA_TRACE_EV(140);
            tc_fence_after();
            const uint32_t a_lo = lo_addr(sQ + qb * Cfg::kTileBytes) | kLoKmajor;
            const uint32_t b_lo = lo_addr(slot_addr(e)) | kLoKmajor;
            umma_issue_qk_d128(tmem_base + (gt % 3) * 128, a_lo, b_lo, kDescHi, kDescHi, idesc_qk);
            umma_commit_elect(bar_s_full(gt % 3));
            commit_slot(e);
            if (last_of_item) umma_commit_elect(bar_q_empty(qb));
            FA_TRACE_EV(120 + gt % 3);
        };

        int ring = 0;   // ring entries of earlier items
        int ka = 0;     // earlier items this CTA took part in (Q buffer = ka & 1)
        int g = 0;      // KV tiles this CTA has taken part in (index of this item's first own tile)
        int kfix = 0;   // ragged tails sanitised for this CTA so far
        int pre = 0;    // K entries (union steps) of the current item already handled by the previous item's tail
        bool q_ready = false;  // the current item's Q tile has been waited for (by the previous item's tail)
        for (int k = 0;; ++k) {
            const int id = get_work(k);
            if (id >= total_work) break;
            const Item2 it = fwd2_item<CL>(p, id, rank);
            const int n = it.n, lo = it.lo, hi = it.hi;
            if (n <= 0) {
                pre = 0;
                q_ready = false;
                continue;
            }
            const bool own_any = hi > lo;
            const int qb = ka % Cfg::kQBufs;
            // K entry of union step u of THIS item
            auto do_k = [&](int u) {
                const int e = ring + fwd2_entry_k(u);
                if (u >= lo && u < hi) {
                    if (!q_ready) {
                        FA_WAIT(bar_q_full(qb), (ka / Cfg::kQBufs) & 1);
                        q_ready = true;
                    }
                    issue_qk(g + (u - lo), qb, e, u == hi - 1);
                } else {
                    release_slot(e);
                }
            };
            for (int u = pre; u < min(n, 2); ++u) do_k(u);
            // the next item, looked at when the walk gets within two steps of this item's end
            int n_next = -1, lo_next = 0, hi_next = 0;
            int issued_next = 0;
            bool q_ready_next = false;
            const int ka_next = ka + (own_any ? 1 : 0);
            const int g_next = g + (hi - lo);
            for (int u = 0; u < n; ++u) {
                if (u + 2 < n) {
                    do_k(u + 2);
                } else {
                    // run ahead into the next item: its steps 0 and 1 take the flat positions n and n + 1
                    if (n_next < 0) {
                        const int nid = Cfg::kQBufs > 1 ? peek_work(k + 1) : total_work;
                        n_next = 0;
                        if (nid < total_work) {
                            const Item2 nx = fwd2_item<CL>(p, nid, rank);
                            n_next = nx.n > 0 ? nx.n : 0;
                            lo_next = nx.lo;
                            hi_next = nx.hi;
                        }
                    }
                    while (issued_next < min(n_next, 2) && n + issued_next <= u + 2) {
                        const int un = issued_next;
                        const int e = ring + 2 * n + fwd2_entry_k(un);
                        if (un >= lo_next && un < hi_next) {
                            if (!q_ready_next) {
                                FA_WAIT(bar_q_full(ka_next % Cfg::kQBufs), (ka_next / Cfg::kQBufs) & 1);
                                q_ready_next = true;
                            }
                            issue_qk(g_next + (un - lo_next), ka_next % Cfg::kQBufs, e, un == hi_next - 1);
                        } else {
                            release_slot(e);
                        }
                        ++issued_next;
                    }
                }
                // O (+)= P(u) V(u)
                const int ev = ring + fwd2_entry_v(u, n);
                if (u >= lo && u < hi) {
                    const int gt = g + (u - lo);
                    wait_full(ev);
                    FA_TRACE_EV(141);
                    if (u == 0 && it.w.ragged_tail) FA_WAIT(bar_vfix, kfix & 1);  // V rows past seqlen_k are zero now
                    FA_WAIT(bar_p_full(gt % 3), (gt / 3) & 1);
                    tc_fence_after();
                    FA_TRACE_EV(100 + gt % 3);
                    const uint32_t v_lo = lo_addr(slot_addr(ev)) | kLoVmn;
                    umma_issue_pv_k0_8(tO, tmem_base + (gt % 3) * 128 + Cfg::kTmemPOff, v_lo, 0, kDescHi, idesc_pv, u > lo ? 1u : 0u);
                    commit_slot(ev);
                    umma_commit_elect(bar_pv_done);
                    FA_TRACE_EV(142);
                } else {
                    release_slot(ev);
                }
            }
            ring += 2 * n;
            g = g_next;
            ka = ka_next;
            kfix += (it.w.ragged_tail && lo == 0 && own_any) ? 1 : 0;
            pre = issued_next;
            q_ready = q_ready_next;
            if (lane == 0) watchdog_progress(sWatch);
        }
        watchdog_role_done(bar_done);
    } else if (warp < 8) {
        // ============================================================ softmax groups (group = warp / 4 takes tiles of its parity)
        reg_inc<192>();
        const int X = warp >> 2;
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        const float sl2 = p.scale_log2;
        int g = 0;       // KV tiles this CTA has taken part in before this item
        int items = 0;   // earlier items in which this group took part

        for (int k = 0;; ++k) {
            const int id = get_work(k);
            if (id >= total_work) break;
            const Item2 it = fwd2_item<CL>(p, id, rank);
            const WorkGeom& w = it.w;
            const int no = it.hi - it.lo;  // own tiles
            if (no <= 0) continue;
            const int i_glob = it.m0 + row;
            int col_hi = w.g.seqlen_k;
            if (p.window_right >= 0) col_hi = min(col_hi, i_glob + w.off + p.window_right + 1);
            int col_lo = 0;
            if (p.window_left >= 0) col_lo = max(0, i_glob + w.off - p.window_left);
            const unsigned col_width = (unsigned)max(col_hi - col_lo, 0);

            float m_own = -INFINITY;  // the reference maximum this group's partial row sum is relative to
            float row_sum = 0.f;
            bool took_part = false;

            for (int j = (X ^ (g & 1)); j < no; j += 2) {  // own tiles whose running index has this group's parity
                const int gt = g + j;
                const int slot = gt % 3;
                const uint32_t tS = tmem_base + lane_off + slot * 128;
                const uint32_t tP = tS + Cfg::kTmemPOff;
                const int j0 = (w.n_max - 1 - (it.lo + j)) * BN;
                FA_WAIT(bar_s_full(slot), (gt / 3) & 1);
                tc_fence_after();
                FA_TRACE_EV(1);
                float v[BN];
                tmem_ld_4x32_wait(tS, reinterpret_cast<uint32_t*>(v));
                FA_TRACE_EV(2);
                const bool need_mask = (j0 + BN > col_hi) || (j0 < col_lo);
                if (__any_sync(0xffffffffu, need_mask)) {
                    const int base = j0 - col_lo;
#pragma unroll
                    for (int c = 0; c < BN; ++c) v[c] = ((unsigned)(base + c) < col_width) ? v[c] : -INFINITY;
                }
                float mx[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) mx[a] = fmaxf(v[2 * a], v[2 * a + 1]);
#pragma unroll
                for (int c = 8; c < BN; c += 8) {
#pragma unroll
                    for (int a = 0; a < 4; ++a) mx[a] = fmax3(mx[a], v[c + 2 * a], v[c + 2 * a + 1]);
                }
                const float tile_max = fmax3(fmaxf(mx[0], mx[1]), mx[2], mx[3]);

                // the running reference maximum after tile gt - 1, from the twin warp of the other group
                float m_prev = -INFINITY;
                if (j > 0) {
                    FA_WAIT(bar_max((gt - 1) & 3, wq), ((gt - 1) >> 2) & 1);
                    m_prev = sMref[((gt - 1) & 3) * BM + row];
                }
                const float m_new = fmaxf(m_prev, tile_max);
                const float m_new_safe = (m_new == -INFINITY) ? 0.f : m_new;
                float acc_scale = 1.0f;
                float m_ref = m_prev;
                if (j == 0) {
                    m_ref = m_new;
                } else {
                    const float d = (m_prev - m_new_safe) * sl2;  // <= 0, -inf if nothing was visible yet
                    if (d < -kRescaleThreshold) {
                        acc_scale = ex2_approx(d);
                        m_ref = m_new;
                        if (p.dbg_counters) atomicAdd(p.dbg_counters, 1ull);
                    }
                }
                sMref[(gt & 3) * BM + row] = m_ref;
                sScale[slot * BM + row] = acc_scale;
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar_max(gt & 3, wq));
                    mbar_arrive(bar_stats(slot, wq));
                }
                FA_TRACE_EV(3);

                const float m_used = (m_ref == -INFINITY) ? 0.f : m_ref;
                if (!took_part) {
                    m_own = m_ref;
                    took_part = true;
                } else if (m_ref != m_own) {
                    // the other group (or this tile) moved the reference: bring the partial sum along
                    const float m_own_safe = (m_own == -INFINITY) ? 0.f : m_own;
                    row_sum *= ex2_approx((m_own_safe - m_used) * sl2);
                    m_own = m_ref;
                }
                const float neg_m = -m_used * sl2;
                float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
                for (int ch = 0; ch < BN / 32; ++ch) {
                    uint32_t pk[16];
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
                        add2(sum0, sum1, p0, p1);
                        pk[c / 2] = pack2<BF16>(p0, p1);
                    }
                    tmem_st_x16(tP + ch * 16, pk);
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_p_full(slot));
                FA_TRACE_EV(5);
                row_sum += sum0 + sum1;
            }
            if (took_part) {
                const int fb = items & 1;
                sFinalL[(X * 2 + fb) * BM + row] = row_sum;
                sFinalM[(X * 2 + fb) * BM + row] = ((m_own == -INFINITY) ? 0.f : m_own) * sl2;
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_final(X, fb, wq));
                ++items;
            }
            g += no;
        }
        watchdog_role_done(bar_done);
    } else if (warp < 12) {
        // ============================================================ correction + epilogue
        reg_dec<80>();
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        const uint32_t tO = tmem_base + lane_off + Cfg::kTmemO;
        uint16_t* outp = reinterpret_cast<uint16_t*>(p.out);
        int g = 0;
        int items_a = 0, items_b = 0;  // items each softmax group took part in (scalars: no dynamically indexed array)

        for (int k = 0;; ++k) {
            const int id = get_work(k);
            if (id >= total_work) break;
            const Item2 it = fwd2_item<CL>(p, id, rank);
            const WorkGeom& w = it.w;
            if (w.skip) continue;
            const int no = it.hi - it.lo;
            if (no <= 0) {
                // No visible key for any row of this CTA's tile: out = 0, lse = sentinel (reference
                // kernel/fused_mha_forward_varlen.cu:100-111); nothing at all if the tile lies past the sequence end.
                const int rows = min(it.m0 + BM, w.g.seqlen_q) - it.m0;
                const int t = (warp - 8) * 32 + lane;
                const int hd8 = p.head_dim / 8;
                for (int idx = t; idx < rows * hd8; idx += 128) {
                    const int r = idx / hd8, c = idx % hd8;
                    uint16_t* dst = outp + w.o_b * p.o_stride_b + (int64_t)(w.g.q_off + it.m0 + r) * p.o_stride_s + w.head * p.o_stride_h + c * 8;
                    *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
                }
                for (int r = t; r < rows; r += 128)
                    p.lse[w.o_b * p.lse_stride_b + w.head * p.lse_stride_h + w.g.q_off + it.m0 + r] = kNegSentinel;
                continue;
            }
            for (int j = 0; j < no; ++j) {
                const int gt = g + j;
                const int slot = gt % 3;
                FA_WAIT(bar_stats(slot, wq), (gt / 3) & 1);
                const float sc = sScale[slot * BM + row];
                if (j > 0) {
                    FA_WAIT(bar_pv_done, (gt - 1) & 1);  // O holds tiles 0 .. j-1 and nothing is writing it
                    if (__any_sync(0xffffffffu, sc != 1.0f)) {
                        if (p.dbg_counters && lane == 0) atomicAdd(p.dbg_counters + 1, 1ull);
                        tc_fence_after();
#pragma unroll
                        for (int c = 0; c < D / 32; ++c) {
                            float o[32];
                            tmem_ld_x32_wait(tO + c * 32, reinterpret_cast<uint32_t*>(o));
#pragma unroll
                            for (int e = 0; e < 32; ++e) o[e] *= sc;
                            tmem_st_x32(tO + c * 32, reinterpret_cast<uint32_t*>(o));
                        }
                        tmem_wait_st();
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_p_full(slot));
            }
            // epilogue: merge the two groups' partial sums, out = O / l, lse = m + ln(l)
            const int gl = g + no - 1;
            const int xl = gl & 1;  // group of the last tile: its reference maximum is the final one
            const int items_l = xl ? items_b : items_a;
            const int fbl = items_l & 1;
            FA_WAIT(bar_final(xl, fbl, wq), (items_l >> 1) & 1);
            float l = sFinalL[(xl * 2 + fbl) * BM + row];
            const float mx = sFinalM[(xl * 2 + fbl) * BM + row];
            if (no >= 2) {
                const int xo = xl ^ 1;
                const int items_o = xo ? items_b : items_a;
                const int fbo = items_o & 1;
                FA_WAIT(bar_final(xo, fbo, wq), (items_o >> 1) & 1);
                const float lo = sFinalL[(xo * 2 + fbo) * BM + row];
                const float mo = sFinalM[(xo * 2 + fbo) * BM + row];
                l += lo * ex2_approx(mo - mx);  // mo <= mx (both already in scaled log2 units)
                ++items_a;
                ++items_b;
            } else if (xl) {
                ++items_b;
            } else {
                ++items_a;
            }
            FA_WAIT(bar_pv_done, gl & 1);
            tc_fence_after();
            FA_TRACE_EV(210);
            const int i_glob = it.m0 + row;
            const bool valid = i_glob < w.g.seqlen_q;
            uint16_t* dst = outp + w.o_b * p.o_stride_b + (int64_t)(w.g.q_off + i_glob) * p.o_stride_s + w.head * p.o_stride_h;
            const float inv = l > 0.f ? 1.0f / l : 0.f;
            const bool wide_ok = __all_sync(0xffffffffu, (reinterpret_cast<uintptr_t>(dst) & 31) == 0);
#pragma unroll
            for (int c = 0; c < D / 32; ++c) {
                if (c * 32 >= p.head_dim) break;  // columns [head_dim, D) are the tile's zero padding
                float o[32];
                tmem_ld_x32_wait(tO + c * 32, reinterpret_cast<uint32_t*>(o));
                if (valid) {
                    uint32_t pk[16];
#pragma unroll
                    for (int e = 0; e < 32; e += 2) pk[e / 2] = pack2<BF16>(o[e] * inv, o[e + 1] * inv);
                    if (wide_ok && c * 32 + 32 <= p.head_dim) {
                        st_global_v8(dst + c * 32, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
                        st_global_v8(dst + c * 32 + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; e += 4)
                            if (c * 32 + 2 * e < p.head_dim)
                                *reinterpret_cast<uint4*>(dst + c * 32 + 2 * e) = make_uint4(pk[e], pk[e + 1], pk[e + 2], pk[e + 3]);
                    }
                }
            }
            if (valid)
                p.lse[w.o_b * p.lse_stride_b + w.head * p.lse_stride_h + w.g.q_off + i_glob] =
                    l > 0.f ? (mx + lg2_approx(l)) * kLn2 : kNegSentinel;
            FA_TRACE_EV(220);
            g += no;
        }
        watchdog_role_done(bar_done);
    } else if (warp == 14) {
        // ============================================================ V sanitiser (see fwd_sm100.cuh)
        reg_dec<48>();
        int ring = 0;
        for (int k = 0;; ++k) {
            const int id = get_work(k);
            if (id >= total_work) break;
            const Item2 it = fwd2_item<CL>(p, id, rank);
            const WorkGeom& w = it.w;
            const int n = it.n;
            if (n <= 0) continue;
            if (w.ragged_tail && it.lo == 0 && it.hi > 0) {  // this CTA multiplies by the ragged tile (union step 0)
                const int v_entry = ring + fwd2_entry_v(0, n);
                const int valid = w.g.seqlen_k - (w.n_max - 1) * BN;
                FA_WAIT(bar_kv_full(v_entry % KV), (v_entry / KV) & 1);
                const uint32_t v_smem = sKV + (v_entry % KV) * Cfg::kTileBytes;
                for (int idx = lane; idx < (BN - valid) * 8 * (D / 64); idx += 32) {
                    const int c16 = idx & 7, r = valid + ((idx >> 3) % (BN - valid)), blk = (idx >> 3) / (BN - valid);
                    st_shared_v4(v_smem + blk * Cfg::kHalfBytes + r * 128 + c16 * 16, 0u, 0u, 0u, 0u);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_vfix);
            }
            ring += 2 * n;
        }
        watchdog_role_done(bar_done);
    } else {
        reg_dec<48>();  // warp 15: watchdog (ptx_sm100.cuh)
        watchdog_run(sWatch, bar_done, sWaitDbg, p.dbg_counters);
    }
#undef FA_WAIT
#undef FA_WAIT_CL

    tc_fence_before();
    __syncthreads();
    if constexpr (CL == 2) cluster_sync_all();  // the peer may still multicast into this CTA's memory until it is done too
    if (warp == 12) tmem_dealloc<512>(tmem_base);
}

}  // namespace fa
