// KV-cache preparation kernels: append k_new/v_new into the cache (RoPE on K) and RoPE on Q.
//
// Replaces WMMA_GEMM_UPDATE_KVCACHE (reference include/rotary.h:14-149) and
// WMMA_GEMM_TILE_ROTARY (reference include/rotary.h:154-264). The reference runs the append inside
// every (q-tile, q-head) CTA of the attention kernel, i.e. H_Q/H_K times redundantly; here it is one
// small HBM-bound kernel per call, launched before the attention kernel on the same stream.
//
// Arithmetic follows the reference exactly (fp32, result rounded to nearest-even 16-bit):
//   first  element of a pair: y0 = fma(x0, cos, -(x1 * sin))
//   second element of a pair: y1 = fma(x0, sin,  (x1 * cos))
// interleaved: pairs (x[2i], x[2i+1]), angle i;  NeoX: pairs (x[i], x[i + rotary_dim/2]), angle i.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace fa {

template <bool BF16>
__device__ __forceinline__ float load16(const uint16_t* p) {
    if constexpr (BF16) return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(p));
    else return __half2float(*reinterpret_cast<const __half*>(p));
}
template <bool BF16>
__device__ __forceinline__ uint16_t store16(float x) {
    if constexpr (BF16) {
        __nv_bfloat16 h = __float2bfloat16_rn(x);
        return *reinterpret_cast<uint16_t*>(&h);
    } else {
        __half h = __float2half_rn(x);
        return *reinterpret_cast<uint16_t*>(&h);
    }
}

// rotate one head vector x[0..D) (already in registers/local) at `pos`; d = element index
template <bool BF16>
__device__ __forceinline__ float rope_elem(const uint16_t* x, int d, int rotary_dim, bool interleaved,
                                           const uint16_t* cos_row, const uint16_t* sin_row) {
    if (d >= rotary_dim) return load16<BF16>(x + d);
    const int half = rotary_dim >> 1;
    int first, second, ang;
    bool is_first;
    if (interleaved) {
        first = d & ~1;
        second = first + 1;
        ang = d >> 1;
        is_first = (d & 1) == 0;
    } else {
        is_first = d < half;
        first = is_first ? d : d - half;
        second = first + half;
        ang = first;
    }
    const float x0 = load16<BF16>(x + first), x1 = load16<BF16>(x + second);
    const float c = load16<BF16>(cos_row + ang), s = load16<BF16>(sin_row + ang);
    return is_first ? __fmaf_rn(x0, c, -(x1 * s)) : __fmaf_rn(x0, s, x1 * c);
}

struct AppendParams {
    const uint16_t* k_new;
    const uint16_t* v_new;
    uint16_t* k_cache;
    uint16_t* v_cache;
    const uint16_t* cos;
    const uint16_t* sin;
    const int* cache_seqlens;
    const int* cache_batch_idx;
    const int* leftpad;
    const int* block_table;
    int64_t knew_sb, knew_ss, knew_sh;
    int64_t vnew_sb, vnew_ss, vnew_sh;
    int64_t kc_sb, kc_ss, kc_sh;
    int64_t vc_sb, vc_ss, vc_sh;
    int batch, seqlen_new, heads_k, head_dim;
    int rotary_dim, interleaved;
    int block_table_stride, page_size;
};

// Programmatic dependent launch (sm_90+): the attention kernel that follows on the stream is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so it may start (and run its set-up: barrier init, TMEM
// allocation, tensor-map prefetch) while this kernel is still running; it executes griddepcontrol.wait before it
// touches anything written here. Both instructions are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_primary() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// one thread per 8 consecutive elements of one (batch, new row, kv head)
template <bool BF16>
__device__ __forceinline__ void kv_append_body(const AppendParams& p, int block, int num_blocks) {
    const int chunks = p.head_dim / 8;
    const int64_t total = (int64_t)p.batch * p.seqlen_new * p.heads_k * chunks;
    for (int64_t idx = block * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)num_blocks * blockDim.x) {
        const int chunk = idx % chunks;
        const int hk = (idx / chunks) % p.heads_k;
        const int r = (idx / ((int64_t)chunks * p.heads_k)) % p.seqlen_new;
        const int b = idx / ((int64_t)chunks * p.heads_k * p.seqlen_new);
        const int len = p.cache_seqlens ? p.cache_seqlens[b] : 0;
        const int pos = len + (p.leftpad ? p.leftpad[b] : 0) + r;  // reference include/rotary.h:62
        int64_t coff_k, coff_v;
        if (p.block_table) {
            const int page = pos / p.page_size;
            const int phys = p.block_table[(int64_t)b * p.block_table_stride + page];
            const int in_page = pos - page * p.page_size;
            coff_k = (int64_t)phys * p.kc_sb + (int64_t)in_page * p.kc_ss + (int64_t)hk * p.kc_sh;
            coff_v = (int64_t)phys * p.vc_sb + (int64_t)in_page * p.vc_ss + (int64_t)hk * p.vc_sh;
        } else {
            const int cb = p.cache_batch_idx ? p.cache_batch_idx[b] : b;
            coff_k = (int64_t)cb * p.kc_sb + (int64_t)pos * p.kc_ss + (int64_t)hk * p.kc_sh;
            coff_v = (int64_t)cb * p.vc_sb + (int64_t)pos * p.vc_ss + (int64_t)hk * p.vc_sh;
        }
        const uint16_t* ksrc = p.k_new + (int64_t)b * p.knew_sb + (int64_t)r * p.knew_ss + (int64_t)hk * p.knew_sh;
        const uint16_t* vsrc = p.v_new + (int64_t)b * p.vnew_sb + (int64_t)r * p.vnew_ss + (int64_t)hk * p.vnew_sh;
        const int d0 = chunk * 8;
        *reinterpret_cast<uint4*>(p.v_cache + coff_v + d0) = *reinterpret_cast<const uint4*>(vsrc + d0);
        if (p.rotary_dim > 0 && d0 < p.rotary_dim) {
            const uint16_t* cos_row = p.cos + (int64_t)pos * (p.rotary_dim / 2);
            const uint16_t* sin_row = p.sin + (int64_t)pos * (p.rotary_dim / 2);
            alignas(16) uint16_t o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
                o[e] = store16<BF16>(rope_elem<BF16>(ksrc, d0 + e, p.rotary_dim, p.interleaved != 0, cos_row, sin_row));
            *reinterpret_cast<uint4*>(p.k_cache + coff_k + d0) = *reinterpret_cast<const uint4*>(o);
        } else {
            *reinterpret_cast<uint4*>(p.k_cache + coff_k + d0) = *reinterpret_cast<const uint4*>(ksrc + d0);
        }
    }
}

struct QRotaryParams {
    const uint16_t* q;
    uint16_t* q_out;  // contiguous (batch, seqlen_q, heads, head_dim)
    const uint16_t* cos;
    const uint16_t* sin;
    const int* cache_seqlens;
    const int* leftpad;
    int64_t q_sb, q_ss, q_sh;
    int batch, seqlen_q, heads, head_dim;
    int rotary_dim, interleaved;
    int per_row_pos;  // causal or local: row r rotates at position len + r; else every row at len
};

template <bool BF16>
__device__ __forceinline__ void q_rotary_body(const QRotaryParams& p, int block, int num_blocks) {
    const int chunks = p.head_dim / 8;
    const int64_t total = (int64_t)p.batch * p.seqlen_q * p.heads * chunks;
    for (int64_t idx = block * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)num_blocks * blockDim.x) {
        const int chunk = idx % chunks;
        const int h = (idx / chunks) % p.heads;
        const int r = (idx / ((int64_t)chunks * p.heads)) % p.seqlen_q;
        const int b = idx / ((int64_t)chunks * p.heads * p.seqlen_q);
        const int len = p.cache_seqlens ? p.cache_seqlens[b] : 0;
        // reference include/rotary.h:176-177, 201-202
        const int pos = len + (p.leftpad ? p.leftpad[b] : 0) + (p.per_row_pos ? r : 0);
        const uint16_t* src = p.q + (int64_t)b * p.q_sb + (int64_t)r * p.q_ss + (int64_t)h * p.q_sh;
        uint16_t* dst = p.q_out + (((int64_t)b * p.seqlen_q + r) * p.heads + h) * p.head_dim;
        const int d0 = chunk * 8;
        if (d0 < p.rotary_dim) {
            const uint16_t* cos_row = p.cos + (int64_t)pos * (p.rotary_dim / 2);
            const uint16_t* sin_row = p.sin + (int64_t)pos * (p.rotary_dim / 2);
            alignas(16) uint16_t o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
                o[e] = store16<BF16>(rope_elem<BF16>(src, d0 + e, p.rotary_dim, p.interleaved != 0, cos_row, sin_row));
            *reinterpret_cast<uint4*>(dst + d0) = *reinterpret_cast<const uint4*>(o);
        } else {
            *reinterpret_cast<uint4*>(dst + d0) = *reinterpret_cast<const uint4*>(src + d0);
        }
    }
}

template <bool BF16>
__global__ void kv_append_kernel(const AppendParams p) {
    pdl_launch_dependents();
    kv_append_body<BF16>(p, (int)blockIdx.x, (int)gridDim.x);
}
template <bool BF16>
__global__ void q_rotary_kernel(const QRotaryParams p) {
    pdl_launch_dependents();
    q_rotary_body<BF16>(p, (int)blockIdx.x, (int)gridDim.x);
}
// append + RoPE(K) and RoPE(Q) in ONE launch (a decode step needs both; each launch costs the host ~2.5 us and the
// GPU a dependent-launch gap). The first `append_blocks` CTAs append, the others rotate Q: both jobs are chains of
// dependent loads (cache_seqlens -> block table -> rows and angles), so they run side by side, not one after the other.
template <bool BF16>
__global__ void kv_prep_kernel(const AppendParams ap, const QRotaryParams rp, int append_blocks) {
    pdl_launch_dependents();
    if ((int)blockIdx.x < append_blocks) kv_append_body<BF16>(ap, (int)blockIdx.x, append_blocks);
    else q_rotary_body<BF16>(rp, (int)blockIdx.x - append_blocks, (int)gridDim.x - append_blocks);
}

}  // namespace fa
