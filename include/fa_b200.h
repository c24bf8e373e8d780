/*
 * fa_b200.h -- C ABI of the B200-native FlashAttention-2 forward and backward (libfa_b200.so).
 *
 * This is the drop-in boundary for the attention hot path of ai-bond/flash-attention-v100.
 * Each entry point replaces one function of the reference's operator layer (the pybind11 module
 * `flash_attn_v100_cuda`, reference kernel/fused_mha_api.cpp:17-33; C++ declarations in reference
 * include/mha.h):
 *
 *   fa_b200_fwd          <->  flash_attention_forward         include/mha.h:27-41
 *                              (kernel/fused_mha_forward.cu:301-432)
 *   fa_b200_varlen_fwd   <->  flash_attention_varlen_forward  include/mha.h:116-139
 *                              (kernel/fused_mha_forward_varlen.cu:371-566)
 *   fa_b200_kvcache_fwd  <->  flash_attention_kvcache         include/mha.h:224-245
 *                              (kernel/fused_mha_forward_kvcache.cu:416-652)
 *   fa_b200_bwd          <->  flash_attention_backward        include/mha.h:67-87
 *                              (kernel/fused_mha_backward.cu:590-721)
 *   fa_b200_varlen_bwd   <->  flash_attention_varlen_backward include/mha.h:170-195
 *                              (kernel/fused_mha_backward_varlen.cu)
 *
 * The reference passes at::Tensor objects and allocates its outputs inside the wrapper; a C ABI
 * cannot, so here every tensor is a raw device pointer plus element strides, and the caller
 * allocates `out`, `lse` and `workspace` (sizes below). The library never synchronises the stream and
 * makes exactly one device allocation per device for its whole lifetime (an 8 KB tile-scheduler ring,
 * the first time that device is used); every call only enqueues kernels on `stream`.
 *
 * Error convention: 0 = success; <0 = invalid argument (FA_B200_EINVAL ...); >0 = cudaError_t of
 * a failed runtime call / launch. fa_b200_last_error() returns a thread-local message for the last
 * non-zero status (the reference raises c10::Error with such a message from TORCH_CHECK).
 */
#ifndef FA_B200_H_
#define FA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FA_B200_ABI_VERSION 3

#if defined(__GNUC__)
#define FA_B200_API __attribute__((visibility("default")))
#else
#define FA_B200_API
#endif

#define FA_B200_DTYPE_FP16 0
#define FA_B200_DTYPE_BF16 1

#define FA_B200_EINVAL (-1)       /* bad argument / unsupported combination */
#define FA_B200_EUNSUPPORTED (-2) /* valid in the reference API but not built yet */
#define FA_B200_EARCH (-3)        /* device is not sm_100 */

/*
 * One parameter block serves the three entry points; fields an entry point does not use must be
 * zero / NULL. All strides are in ELEMENTS of the tensor's dtype; the last (head_dim) stride of
 * q/k/v/out/k_new/v_new must be 1 (reference check kernel/fused_mha_forward.cu:327).
 *
 * Logical layouts (any strides):
 *   dense    q,out:(batch, seqlen_q, num_heads, head_dim)  k,v:(batch, seqlen_k, num_heads_k, head_dim)
 *            lse:(batch, num_heads, seqlen_q) fp32
 *   varlen   q,out:(total_q, num_heads, head_dim)  k,v:(total_k, num_heads_k, head_dim)
 *            [q_stride_b/k_stride_b unused]  lse:(num_heads, total_q) fp32
 *            or paged k,v:(num_pages, page_size, num_heads_k, head_dim) with block_table
 *   kvcache  q,out:(batch, seqlen_q, num_heads, head_dim)
 *            k,v = the CACHE:(batch_cache, seqlen_k, num_heads_k, head_dim), or paged as above
 *            k_new,v_new:(batch, seqlen_new, num_heads_k, head_dim) appended in place first
 *            lse:(batch, num_heads, seqlen_q) fp32
 */
typedef struct fa_b200_params {
    int32_t struct_bytes; /* = sizeof(fa_b200_params_t); checked */
    int32_t dtype;        /* FA_B200_DTYPE_* (the reference is fp16-only; bf16 is an addition) */
    int32_t device;       /* CUDA device ordinal the pointers live on */
    int32_t reserved0;

    /* sizes */
    int32_t batch;       /* number of sequences */
    int32_t seqlen_q;    /* dense/kvcache: Sq; varlen: max_seqlen_q */
    int32_t seqlen_k;    /* dense: Sk; varlen: max_seqlen_k; kvcache: cache capacity per sequence
                            (paged: block_table_cols * page_size) */
    int32_t num_heads;   /* query heads */
    int32_t num_heads_k; /* key/value heads; num_heads % num_heads_k == 0 */
    int32_t head_dim;    /* any multiple of 8 up to 256 (reference kernel/fused_mha_forward.cu:335-336); the
                            kernels' tiles are 64 / 128 / 256 wide, TMA zero-fills the columns in between */
    int32_t total_q;     /* varlen: rows of q; else 0 */
    int32_t total_k;     /* varlen non-paged: rows of k; else 0 */

    /* tensors */
    const void* q;
    const void* k;
    const void* v;
    void* out;
    float* lse;
    int64_t q_stride_b, q_stride_s, q_stride_h;
    int64_t k_stride_b, k_stride_s, k_stride_h; /* paged: _b = page stride, _s = row-in-page stride */
    int64_t v_stride_b, v_stride_s, v_stride_h;
    int64_t o_stride_b, o_stride_s, o_stride_h;
    int32_t batch_k; /* kvcache: rows of the cache's batch dim (>= batch when cache_batch_idx) */
    int32_t reserved1;

    /* varlen (reference kernel/fused_mha_forward_varlen.cu:452-467) */
    const int32_t* cu_seqlens_q; /* (batch+1) */
    const int32_t* cu_seqlens_k; /* (batch+1) */
    const int32_t* seqused_k;    /* (batch) or NULL: clamps each key length */

    /* paged KV (reference kernel/fused_mha_forward_varlen.cu:434-449, ..._kvcache.cu:484-500) */
    const int32_t* block_table; /* (batch, block_table_cols) page ids, or NULL */
    int32_t block_table_stride; /* elements between rows of block_table */
    int32_t page_size;          /* rows per page; a multiple of 128 (one KV tile never straddles two pages). The
                                   reference requires 256 (kernel/fused_mha_forward_kvcache.cu:506-509); every
                                   page size it accepts is accepted here, and 128 / 384 / ... in addition */
    int32_t num_pages;
    int32_t reserved2;

    /* kv-cache (reference kernel/fused_mha_forward_kvcache.cu:79-86, include/rotary.h:53-76) */
    const int32_t* cache_seqlens;   /* (batch) current lengths; NULL = full cache (upstream API) */
    const int32_t* cache_batch_idx; /* (batch) cache row per sequence, or NULL */
    const int32_t* cache_leftpad;   /* (batch) first valid cache row, or NULL */
    const void* k_new;
    const void* v_new;
    int64_t knew_stride_b, knew_stride_s, knew_stride_h;
    int64_t vnew_stride_b, vnew_stride_s, vnew_stride_h;
    int32_t seqlen_new;
    int32_t rotary_dim; /* 0 = no rotary; else <= head_dim, multiple of 16 */
    const void* rotary_cos; /* (rotary_seqlen, rotary_dim/2), same dtype as q, contiguous */
    const void* rotary_sin;
    int32_t rotary_seqlen;
    int32_t rotary_interleaved; /* 1 = GPT-J pairs (x[2i],x[2i+1]); 0 = NeoX halves */

    /* score modifiers (reference include/mat_mul.h:82-157) */
    const float* alibi_slopes; /* (num_heads) or (batch, num_heads) fp32, or NULL */
    int64_t alibi_stride_b;    /* 0 for the (num_heads) form */
    float softmax_scale;
    float softcap;       /* 0 = off */
    int32_t is_causal;   /* bottom-right aligned */
    int32_t window_left; /* -1 = unbounded */
    int32_t window_right;
    int32_t num_splits; /* kvcache decode: 0 = choose; the reference rejects > 1 */

    /* scratch for split-KV partial results; see fa_b200_workspace_bytes */
    void* workspace;
    int64_t workspace_bytes;

    /* dropout (dense and varlen; reference include/softmax.h:96-125, include/philox.h:59-119):
     * element (row, col) is kept iff word[(idx & 3)] of Philox4x32-10(key = seed, counter = offset + (idx >> 2))
     * is <= (1 - p) * (2^32 - 1), with idx = row * dropout_cols + col; kept P is scaled by 1/(1-p), the row
     * sum uses P before dropout. As in the reference the index ignores batch and head. */
    float p_dropout;       /* 0 = off */
    int32_t reserved3;
    uint64_t dropout_seed;
    uint64_t dropout_offset;
    void* dmask;           /* optional: +1.0 kept / -1.0 dropped in q's dtype; dense (B,H,Sq,Sk), varlen (total_q,H,max_seqlen_k) */

    /* backward (fa_b200_bwd / fa_b200_varlen_bwd; reference include/mha.h:67-87, 170-195). Inputs: q, k, v,
     * out, lse (from the forward), dout (layout of out). Outputs: dq, dk, dv (layouts of q, k, v; dk/dv have
     * num_heads_k heads: the GQA group is summed inside the kernel) and softmax_d = rowsum(dout * out), fp32
     * with the layout of lse. With p_dropout > 0 pass the forward's dropout_seed / dropout_offset (its rng_state). */
    const void* dout;
    void* dq;
    void* dk;
    void* dv;
    float* softmax_d;
    int64_t do_stride_b, do_stride_s, do_stride_h;
    int64_t dq_stride_b, dq_stride_s, dq_stride_h;
    int64_t dk_stride_b, dk_stride_s, dk_stride_h;
    int64_t dv_stride_b, dv_stride_s, dv_stride_h;
} fa_b200_params_t;

/* kinds for fa_b200_workspace_bytes */
#define FA_B200_KIND_DENSE 0
#define FA_B200_KIND_VARLEN 1
#define FA_B200_KIND_KVCACHE 2

FA_B200_API int fa_b200_abi_version(void);

/* One-time per-device set-up: allocates the library's only device memory, 192 KB of per-launch tile-scheduler
 * counters. Optional -- the first launch on a device does it implicitly -- except when that first launch would
 * happen while a CUDA graph is being captured (cudaMalloc is not capturable): call it once before capturing.
 * Replaces nothing in the reference (its kernels have no scheduler state). Returns 0 or an error code. */
FA_B200_API int fa_b200_init(int device);

/* Thread-local text of the last error returned on this thread ("" if none). */
FA_B200_API const char* fa_b200_last_error(void);

/* Bytes of device scratch the call described by (params, kind) needs (0 for most shapes). */
FA_B200_API int64_t fa_b200_workspace_bytes(const fa_b200_params_t* params, int kind);

/* Dense forward: replaces flash_attention_forward (reference include/mha.h:27-41). */
FA_B200_API int fa_b200_fwd(const fa_b200_params_t* params, void* cuda_stream);

/* Packed variable-length forward: replaces flash_attention_varlen_forward (include/mha.h:116-139). */
FA_B200_API int fa_b200_varlen_fwd(const fa_b200_params_t* params, void* cuda_stream);

/* KV-cache forward (append + rotary + attention): replaces flash_attention_kvcache
 * (include/mha.h:224-245). Mutates the cache in place, like the reference. */
FA_B200_API int fa_b200_kvcache_fwd(const fa_b200_params_t* params, void* cuda_stream);

/* Dense backward: replaces flash_attention_backward (reference include/mha.h:67-87). Deterministic. */
FA_B200_API int fa_b200_bwd(const fa_b200_params_t* params, void* cuda_stream);

/* Packed variable-length backward: replaces flash_attention_varlen_backward (include/mha.h:170-195). */
FA_B200_API int fa_b200_varlen_bwd(const fa_b200_params_t* params, void* cuda_stream);

/* Number of CUDA kernels this library has launched in this process (all threads). */
FA_B200_API int64_t fa_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FA_B200_H_ */
