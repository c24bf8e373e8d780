// libfa_b200.so: the C ABI declared in include/fa_b200.h.
//
// Host side of the forward hot path: argument validation and the normalisations the reference does
// in its C++ wrappers (kernel/fused_mha_forward.cu:301-432, kernel/fused_mha_forward_varlen.cu:
// 371-566, kernel/fused_mha_forward_kvcache.cu:416-652), TMA tensor-map encoding, and the launches.
// No torch / ATen types, no device allocation, no stream synchronisation.
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/fa_b200.h"
#include "fwd_sm100.cuh"
#include "fwd2_sm100.cuh"
#include "bwd_sm100.cuh"
#include "kvcache_prep.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
}
#define CHECK_ARG(cond, ...) \
    do {                     \
        if (!(cond)) return fail(FA_B200_EINVAL, __VA_ARGS__); \
    } while (0)

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

// 4-D view (head_dim, heads, rows, batch) of a 16-bit tensor with unit head_dim stride;
// box = 64 x 1 x 128 x 1 with the 128-byte swizzle the UMMA descriptors expect.
// Encoded maps are cached per thread, keyed by everything that goes into them: a serving / training loop calls
// with the same buffers and shapes over and over, and cuTensorMapEncodeTiled is ~1 us of driver time per operand
// (three per forward call, nine per backward) -- a third of the host cost of a small call.
struct TmapKey {
    const void* ptr;
    int64_t heads, rows, batch, stride_h, stride_s, stride_b;
    int dtype, head_dim, box_heads, box_rows;
    bool operator==(const TmapKey& o) const {
        return ptr == o.ptr && heads == o.heads && rows == o.rows && batch == o.batch && stride_h == o.stride_h &&
               stride_s == o.stride_s && stride_b == o.stride_b && dtype == o.dtype && head_dim == o.head_dim &&
               box_heads == o.box_heads && box_rows == o.box_rows;
    }
};
struct TmapCache {
    static constexpr int kEntries = 64;
    TmapKey key[kEntries];
    CUtensorMap map[kEntries];
    bool valid[kEntries] = {};
};
thread_local TmapCache g_tmap_cache;

int make_tmap_uncached(CUtensorMap* tm, int dtype, const void* ptr, int head_dim, int64_t heads, int64_t rows,
                       int64_t batch, int64_t stride_h, int64_t stride_s, int64_t stride_b, const char* name,
                       int box_heads, int box_rows);

int make_tmap(CUtensorMap* tm, int dtype, const void* ptr, int head_dim, int64_t heads, int64_t rows,
              int64_t batch, int64_t stride_h, int64_t stride_s, int64_t stride_b, const char* name,
              int box_heads = 1, int box_rows = 128) {
    const TmapKey k{ptr, heads, rows, batch, stride_h, stride_s, stride_b, dtype, head_dim, box_heads, box_rows};
    uint64_t h = reinterpret_cast<uintptr_t>(ptr) >> 4;
    h = (h ^ (uint64_t)rows * 0x9E3779B97F4A7C15ull ^ (uint64_t)stride_s * 0xC2B2AE3D27D4EB4Full ^ (uint64_t)box_rows) * 0xFF51AFD7ED558CCDull;
    const int slot = (int)(h >> 58);  // 64 entries, direct mapped
    TmapCache& c = g_tmap_cache;
    if (c.valid[slot] && c.key[slot] == k) {
        memcpy(tm, &c.map[slot], sizeof(CUtensorMap));
        return 0;
    }
    if (int rc = make_tmap_uncached(tm, dtype, ptr, head_dim, heads, rows, batch, stride_h, stride_s, stride_b, name, box_heads, box_rows))
        return rc;
    c.key[slot] = k;
    memcpy(&c.map[slot], tm, sizeof(CUtensorMap));
    c.valid[slot] = true;
    return 0;
}

int make_tmap_uncached(CUtensorMap* tm, int dtype, const void* ptr, int head_dim, int64_t heads, int64_t rows,
                       int64_t batch, int64_t stride_h, int64_t stride_s, int64_t stride_b, const char* name,
                       int box_heads, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(FA_B200_EARCH, "cuTensorMapEncodeTiled is not available from this driver");
    CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "%s must be 16-byte aligned", name);
    cuuint64_t dims[4] = {(cuuint64_t)head_dim, (cuuint64_t)(heads > 0 ? heads : 1), (cuuint64_t)(rows > 0 ? rows : 1),
                          (cuuint64_t)(batch > 0 ? batch : 1)};
    // A dimension that is stepped over (extent > 1) needs a positive stride that is a multiple of 16 bytes: an
    // expanded (stride-0) view cannot be described to TMA and must be materialised by the caller (the reference
    // calls .contiguous()). A dimension of extent 1 is never stepped over; it gets a dummy 16-byte stride.
    const int64_t in_strides[3] = {stride_h, stride_s, stride_b};
    cuuint64_t strides[3];
    for (int i = 0; i < 3; ++i) {
        if (dims[i + 1] > 1) {
            CHECK_ARG(in_strides[i] > 0, "%s has a non-positive stride (%lld) over a dimension of extent %llu: expanded / "
                      "broadcast views are not supported, pass a materialised tensor", name, (long long)in_strides[i],
                      (unsigned long long)dims[i + 1]);
            CHECK_ARG(in_strides[i] % 8 == 0, "%s strides must be multiples of 8 elements (16 bytes)", name);
            strides[i] = (cuuint64_t)in_strides[i] * 2;
        } else {
            strides[i] = 16;
        }
    }
    cuuint32_t box[4] = {64, (cuuint32_t)box_heads, (cuuint32_t)box_rows, 1};  // box_heads * box_rows == 128 tile rows
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, dtype == FA_B200_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                     4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FA_B200_EINVAL, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", name, (int)r);
    return 0;
}

int check_device(int device) {
    static std::atomic<int> cached_major[64];  // 0 = not queried yet; threads may race to fill it with the same value
    if (device < 0 || device >= 64) return fail(FA_B200_EINVAL, "bad device ordinal %d", device);
    int major = cached_major[device].load(std::memory_order_relaxed);
    if (major == 0) {
        cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
        if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
        cached_major[device].store(major, std::memory_order_relaxed);
    }
    if (major != 10)
        return fail(FA_B200_EARCH, "this library only runs on sm_100 (B200); device %d is sm_%dx", device, major);
    return 0;
}

// Tile width the kernels are built for: the real head_dim (any multiple of 8) is rounded up to it and TMA
// zero-fills the columns in between (the tensor maps carry the real head_dim as their innermost extent).
int tile_dim(int head_dim) { return head_dim <= 64 ? 64 : head_dim <= 128 ? 128 : 256; }

// `pdl`: launch with programmatic stream serialisation -- the kernel may start while its predecessor on the stream
// (the kv-cache preparation kernel) is still running and waits for it with griddepcontrol.wait (decode path only).
template <int D, bool BF16, bool FEAT, bool DECODE = false, bool DROPOUT = false>
int launch_fwd_t(const fa::FwdKernelParams& kp, dim3 grid, cudaStream_t stream, bool pdl = false) {
    using Cfg = fa::FwdConfig<D>;
    auto kern = fa::fa_fwd_sm100_kernel<D, BF16, FEAT, DECODE, DROPOUT>;
    static std::once_flag once[64];
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t attr_err = cudaSuccess;
    std::call_once(once[dev & 63], [&] {
        attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    });
    if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(smem)");
    cudaError_t e;
    if (pdl) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = grid;
        cfg.blockDim = dim3(512, 1, 1);
        cfg.dynamicSmemBytes = Cfg::kSmemBytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, kp);
    } else {
        kern<<<grid, 512, Cfg::kSmemBytes, stream>>>(kp);
        e = cudaSuccess;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "fa_fwd_sm100_kernel launch");
    return 0;
}

// Forward v2 (fwd2_sm100.cuh: one 128-row query tile per CTA, three score buffers, two alternating softmax groups,
// optionally CTA pairs sharing their K/V stream through TMA multicast) can serve head dims 65..128 without score
// modifiers or dropout. It is parity-clean (the whole GPU suite passes on it) but measured slower than the round-1
// kernel on this hardware, so it is opt-in: FA_B200_FWD_KERNEL=2s (single CTAs) or =2p (CTA pairs) in the
// environment, read once per process. Everything else runs the round-1 kernel.
int fwd2_mode() {  // 0 = off, 1 = single CTAs, 2 = CTA pairs
    static const int mode = [] {
        const char* e = getenv("FA_B200_FWD_KERNEL");
        if (e && e[0] == '2' && e[1] == 's') return 1;
        if (e && e[0] == '2') return 2;
        return 0;  // default: the round-1 kernel, which measures faster on every BASELINE shape (forward v2)
    }();
    return mode;
}
int fwd2_cluster(const fa_b200_params_t* p) {  // 0: not a v2 shape; else CTAs per cluster
    if (fwd2_mode() == 0 || tile_dim(p->head_dim) != 128 || p->alibi_slopes != nullptr || p->softcap != 0.f ||
        p->p_dropout != 0.f || p->dmask != nullptr)
        return 0;
    return fwd2_mode();
}

template <bool BF16, int CL>
int launch_fwd2_t(const fa::FwdKernelParams& kp, dim3 grid, cudaStream_t stream) {
    using Cfg = fa::Fwd2Config;
    auto kern = fa::fa_fwd2_sm100_kernel<BF16, CL>;
    static std::once_flag once[64];
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t attr_err = cudaSuccess;
    std::call_once(once[dev & 63], [&] {
        attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    });
    if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(smem)");
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(512, 1, 1);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, kp);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "fa_fwd2_sm100_kernel launch");
    return 0;
}

int launch_fwd(const fa::FwdKernelParams& kp, int real_head_dim, int dtype, bool feat, dim3 grid, cudaStream_t stream, int v2 = 0) {
    const bool bf16 = dtype == FA_B200_DTYPE_BF16;
    const int head_dim = tile_dim(real_head_dim);
    if (v2 == 2) return bf16 ? launch_fwd2_t<true, 2>(kp, grid, stream) : launch_fwd2_t<false, 2>(kp, grid, stream);
    if (v2 == 1) return bf16 ? launch_fwd2_t<true, 1>(kp, grid, stream) : launch_fwd2_t<false, 1>(kp, grid, stream);
    if (kp.drop_thr != 0xffffffffu || kp.dmask != nullptr) {  // dropout on (a clamped threshold keeps everything but still writes dmask): one variant per (D, dtype), with the score-modifier path compiled in
        if (head_dim == 128 && bf16) return launch_fwd_t<128, true, true, false, true>(kp, grid, stream);
        if (head_dim == 128) return launch_fwd_t<128, false, true, false, true>(kp, grid, stream);
        if (head_dim == 64 && bf16) return launch_fwd_t<64, true, true, false, true>(kp, grid, stream);
        if (head_dim == 64) return launch_fwd_t<64, false, true, false, true>(kp, grid, stream);
        if (head_dim == 256 && bf16) return launch_fwd_t<256, true, true, false, true>(kp, grid, stream);
        if (head_dim == 256) return launch_fwd_t<256, false, true, false, true>(kp, grid, stream);
    }
#define FA_CASE(DD, BB, FF) \
    if (head_dim == DD && bf16 == BB && feat == FF) return launch_fwd_t<DD, BB, FF>(kp, grid, stream);
    FA_CASE(128, true, false)
    FA_CASE(128, true, true)
    FA_CASE(128, false, false)
    FA_CASE(128, false, true)
    FA_CASE(64, true, false)
    FA_CASE(64, true, true)
    FA_CASE(64, false, false)
    FA_CASE(64, false, true)
    FA_CASE(256, true, false)
    FA_CASE(256, true, true)
    FA_CASE(256, false, false)
    FA_CASE(256, false, true)
#undef FA_CASE
    return fail(FA_B200_EUNSUPPORTED, "head_dim %d is not built", head_dim);
}

// Decode path: packed-GQA split-KV launch + combine.
int launch_decode(const fa::FwdKernelParams& kp, int real_head_dim, int dtype, dim3 grid, cudaStream_t stream, bool pdl) {
    const bool bf16 = dtype == FA_B200_DTYPE_BF16;
    const int head_dim = tile_dim(real_head_dim);
    if (head_dim == 128 && bf16) return launch_fwd_t<128, true, false, true>(kp, grid, stream, pdl);
    if (head_dim == 128 && !bf16) return launch_fwd_t<128, false, false, true>(kp, grid, stream, pdl);
    if (head_dim == 64 && bf16) return launch_fwd_t<64, true, false, true>(kp, grid, stream, pdl);
    if (head_dim == 64 && !bf16) return launch_fwd_t<64, false, false, true>(kp, grid, stream, pdl);
    if (head_dim == 256 && bf16) return launch_fwd_t<256, true, false, true>(kp, grid, stream, pdl);
    if (head_dim == 256 && !bf16) return launch_fwd_t<256, false, false, true>(kp, grid, stream, pdl);
    return fail(FA_B200_EUNSUPPORTED, "head_dim %d is not built", head_dim);
}

int launch_combine(const fa::FwdKernelParams& kp, const fa_b200_params_t* p, cudaStream_t stream) {
    const int64_t rows = (int64_t)p->batch * p->num_heads * p->seqlen_q;
    const int warps = 4;
    const dim3 grid((unsigned)((rows + warps - 1) / warps));
    uint16_t* out = static_cast<uint16_t*>(p->out);
    const bool bf16 = p->dtype == FA_B200_DTYPE_BF16;
    // launched with programmatic stream serialisation: it starts while the decode kernel drains and waits for it
    // with griddepcontrol.wait
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(warps * 32, 1, 1);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaSuccess;
#define FA_COMBINE(DD, BB)                                                                                   \
    le = cudaLaunchKernelEx(&cfg, fa::fa_combine_kernel<DD, BB>, (const float*)kp.o_partial, (const float*)kp.lse_partial, out, p->lse, \
        kp.num_splits, p->batch, p->num_heads, p->seqlen_q, p->o_stride_b, p->o_stride_s, p->o_stride_h, p->head_dim)
    if (tile_dim(p->head_dim) == 256 && bf16) FA_COMBINE(256, true);
    else if (tile_dim(p->head_dim) == 256) FA_COMBINE(256, false);
    else if (tile_dim(p->head_dim) == 128 && bf16) FA_COMBINE(128, true);
    else if (tile_dim(p->head_dim) == 128) FA_COMBINE(128, false);
    else if (bf16) FA_COMBINE(64, true);
    else FA_COMBINE(64, false);
#undef FA_COMBINE
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = le != cudaSuccess ? le : cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "fa_combine_kernel launch");
    return 0;
}

// Decode applies when a whole GQA group's query rows fit one 128-row tile and no score modifier is on.
bool decode_applies(const fa_b200_params_t* p) {
    const int G = p->num_heads / p->num_heads_k;
    if (G > 128 || (G & (G - 1)) != 0) return false;
    if ((int64_t)p->seqlen_q * G > 128) return false;
    return p->alibi_slopes == nullptr && p->softcap == 0.f;
}

// KV splits per (batch, kv head): fill the 148 SMs for several waves but keep >= 4 tiles per split.
int decode_num_splits(const fa_b200_params_t* p) {
    if (p->num_splits == 1) return 1;
    static const int forced = [] {  // tuning aid (tools/gpu_call_r02_23.sh): FA_B200_DECODE_SPLITS=<n> overrides the heuristic
        const char* e = getenv("FA_B200_DECODE_SPLITS");
        return e ? atoi(e) : 0;
    }();
    if (forced > 0) {
        const int tiles_f = (p->seqlen_k + 127) / 128;
        return forced < tiles_f ? (forced < 64 ? forced : 64) : (tiles_f < 64 ? tiles_f : 64);
    }
    const int tiles = (p->seqlen_k + 127) / 128;
    const int64_t base = (int64_t)p->batch * p->num_heads_k;
    int best = 1;
    double best_score = 0.0;
    for (int s = 1; s <= 64 && s * 4 <= (tiles > 4 ? tiles : 4); ++s) {
        const int64_t ctas = base * s;
        const double waves = (double)ctas / 148.0;
        const double eff = waves / (double)((ctas + 147) / 148);      // wave quantisation
        const double per = (double)((tiles + s - 1) / s);
        const double score = eff * per / (per + 1.5);                  // ~1.5 tiles of fixed cost per CTA
        if (score > best_score * 1.02) {
            best_score = score;
            best = s;
        }
    }
    return best;
}

int64_t decode_workspace_bytes(const fa_b200_params_t* p) {
    if (!decode_applies(p)) return 0;
    const int s = decode_num_splits(p);
    if (s <= 1) return 0;
    const int64_t rows = (int64_t)s * p->batch * p->num_heads * p->seqlen_q;
    return ((rows * tile_dim(p->head_dim) * 4 + 255) & ~(int64_t)255) + ((rows * 4 + 255) & ~(int64_t)255);
}

// Debug counters of the forward kernel (tests only): see FwdKernelParams::dbg_counters.
std::atomic<unsigned long long*> g_dbg_counters{nullptr};

// Tile-scheduler counters (FwdKernelParams::sched): every launch with more work items than CTAs gets one 8-byte
// counter of its own, zeroed by a cudaMemsetAsync on the launch stream right before the kernel, so a launch never
// depends on what an earlier user of the slot left behind and the kernel has nothing to re-arm.
//  * eager launches take slots round-robin from a ring of kEagerSlots: two launches could only share a live counter
//    if kEagerSlots launches were enqueued while one kernel is still running;
//  * launches recorded into a CUDA graph (cudaStreamIsCapturing) take slots from a separate pool that is never
//    recycled: the slot pointer is baked into the graph for good, so it must never be handed to anybody else.
//    Replays of one graph instance serialise on the device, and every replay re-runs its own memset node.
// The pools (192 KB per device) are the only memory this library ever allocates; fa_b200_init(device) creates them
// ahead of time (it must run outside stream capture; the first launch on a device calls it otherwise).
constexpr unsigned kEagerSlots = 16384, kCaptureSlots = 8192;
struct SchedPools {
    int* eager = nullptr;
    int* capture = nullptr;
    std::atomic<unsigned> next_eager{0}, next_capture{0};
};
SchedPools g_pools[64];
std::once_flag g_pools_once[64];
cudaError_t g_pools_err[64];

int init_pools(int device) {
    std::call_once(g_pools_once[device & 63], [&] {
        int prev = -1;
        cudaGetDevice(&prev);
        cudaError_t e = cudaSetDevice(device);
        int* ptr = nullptr;
        if (e == cudaSuccess) e = cudaMalloc(&ptr, (size_t)(kEagerSlots + kCaptureSlots) * 2 * sizeof(int));
        if (e == cudaSuccess) {
            g_pools[device & 63].eager = ptr;
            g_pools[device & 63].capture = ptr + 2 * kEagerSlots;
        }
        if (prev >= 0 && prev != device) cudaSetDevice(prev);
        g_pools_err[device & 63] = e;
        if (e != cudaSuccess) cudaGetLastError();  // clear the sticky-free error state
    });
    if (g_pools_err[device & 63] != cudaSuccess)
        return fail((int)g_pools_err[device & 63],
                    "fa_b200_init(%d): could not allocate the tile-scheduler counters (%s); if the first call on this "
                    "device happens under CUDA-graph capture, call fa_b200_init(device) once beforehand",
                    device, cudaGetErrorString(g_pools_err[device & 63]));
    return 0;
}

// Counter for one launch, zeroed on `stream`. *slot = nullptr when the launch needs none.
int take_sched_slot(int device, cudaStream_t stream, int** slot) {
    *slot = nullptr;
    if (int rc = init_pools(device)) return rc;
    SchedPools& pl = g_pools[device & 63];
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) {
        cudaGetLastError();
        st = cudaStreamCaptureStatusNone;
    }
    int* s;
    if (st == cudaStreamCaptureStatusActive) {
        const unsigned i = pl.next_capture.fetch_add(1, std::memory_order_relaxed);
        if (i >= kCaptureSlots)
            return fail(FA_B200_EUNSUPPORTED, "more than %u attention launches were recorded into CUDA graphs on device %d: "
                        "the pool of per-launch scheduler counters for captured launches is exhausted", kCaptureSlots, device);
        s = pl.capture + 2 * i;
    } else {
        s = pl.eager + 2 * (pl.next_eager.fetch_add(1, std::memory_order_relaxed) % kEagerSlots);
    }
    cudaError_t e = cudaMemsetAsync(s, 0, 2 * sizeof(int), stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(scheduler counter)");
    *slot = s;
    return 0;
}

int sm_count(int device) {
    static std::atomic<int> cached[64];
    int n = cached[device & 63].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
        cached[device & 63].store(n, std::memory_order_relaxed);
    }
    return n;
}

// Persistent 1-D grid; work items in sectioned longest-first order (see FwdKernelParams::section_bh).
// `v2` = CTAs per work item of forward v2 (0: the round-1 kernel). A v2 pair takes a 256-row item, like the round-1 CTA.
dim3 set_launch_order(fa::FwdKernelParams& kp, int device, int batch, int heads, int heads_k, int max_seqlen_q, int max_seqlen_k, int head_dim, int v2 = 0) {
    const int item_rows = (v2 == 1 || tile_dim(head_dim) == 256) ? 128 : 256;  // FwdConfig<D>::kItemRows; v2 single CTA: one 128-row tile
    kp.num_m_blocks = (max_seqlen_q + item_rows - 1) / item_rows;
    kp.num_bh = batch * heads;
    // K+V bytes one kv head streams; keep a section's K/V within ~32 MB of the 126 MB L2
    const int64_t kv_bytes = 2ll * max_seqlen_k * head_dim * 2;
    const int group = heads / heads_k;
    int64_t sec = (32ll << 20) / (kv_bytes > 0 ? kv_bytes : 1) * group;
    if (sec < group) sec = group;
    if (sec > kp.num_bh) sec = kp.num_bh;
    kp.section_bh = (int)sec;
    kp.dbg_counters = g_dbg_counters.load(std::memory_order_relaxed);
    kp.sched = nullptr;
    const int64_t total = (int64_t)kp.num_m_blocks * kp.num_bh;
#if FA_SCHED_CLC
    (void)device;
    return dim3((unsigned)total, 1, 1);
#else
    const int cl = v2 == 2 ? 2 : 1;
    const int units = sm_count(device) / cl;  // CTAs, or CTA pairs, that fit the device at once
    return dim3((unsigned)((total < units ? total : units) * cl), 1, 1);
#endif
}

// A launch with more work items than CTAs needs a scheduler counter (not under FA_SCHED_CLC).
int arm_scheduler(fa::FwdKernelParams& kp, dim3 grid, int device, cudaStream_t stream, int cl = 1) {
#if !FA_SCHED_CLC
    if ((int64_t)kp.num_m_blocks * kp.num_bh > (int64_t)grid.x / cl) return take_sched_slot(device, stream, &kp.sched);
#endif
    (void)grid; (void)device; (void)stream; (void)cl;
    return 0;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    // cudaSetDevice is called even when `dev` is already the current device: on a thread that has not touched CUDA
    // yet (PyTorch's autograd thread calling the backward) it is what binds the primary context to the thread, and
    // cuTensorMapEncodeTiled is a driver call that fails with CUDA_ERROR_INVALID_CONTEXT without one.
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && cudaSetDevice(dev) != cudaSuccess) ok = false;
        if (ok) target = dev;
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != target) cudaSetDevice(prev);
    }
    int target = -1;
};

int check_common(const fa_b200_params_t* p) {
    if (!p) return fail(FA_B200_EINVAL, "params is NULL");
    CHECK_ARG(p->struct_bytes == (int32_t)sizeof(fa_b200_params_t),
              "fa_b200_params_t size mismatch: caller %d, library %d (ABI %d)", p->struct_bytes,
              (int)sizeof(fa_b200_params_t), FA_B200_ABI_VERSION);
    CHECK_ARG(p->dtype == FA_B200_DTYPE_FP16 || p->dtype == FA_B200_DTYPE_BF16, "dtype must be fp16 (0) or bf16 (1)");
    CHECK_ARG(p->batch > 0, "batch size must be positive");
    CHECK_ARG(p->num_heads > 0 && p->num_heads_k > 0, "head counts must be positive");
    CHECK_ARG(p->num_heads % p->num_heads_k == 0, "H_Q must be divisible by H_K for GQA/MQA");
    CHECK_ARG(p->head_dim % 8 == 0, "head dimension must be multiple of 8");
    CHECK_ARG(p->head_dim > 0 && p->head_dim <= 256, "head dimension must be <= 256");
    CHECK_ARG(p->q && p->k && p->v && p->out && p->lse, "q, k, v, out and lse must be non-NULL");
    CHECK_ARG(p->softcap >= 0.f, "softcap must be >= 0");
    CHECK_ARG((reinterpret_cast<uintptr_t>(p->out) & 15) == 0 && p->o_stride_b % 8 == 0 && p->o_stride_s % 8 == 0 &&
                  p->o_stride_h % 8 == 0,
              "out must be 16-byte aligned with strides that are multiples of 8 elements");
    return check_device(p->device);
}

void fill_common(fa::FwdKernelParams& kp, const fa_b200_params_t* p, bool causal, int wl, int wr) {
    memset(&kp, 0, sizeof(kp));
    kp.out = p->out;
    kp.lse = p->lse;
    kp.o_stride_b = p->o_stride_b;
    kp.o_stride_s = p->o_stride_s;
    kp.o_stride_h = p->o_stride_h;
    kp.alibi = p->alibi_slopes;
    kp.alibi_stride_b = p->alibi_stride_b;
    kp.num_heads = p->num_heads;
    kp.heads_per_kv = p->num_heads / p->num_heads_k;
    kp.head_dim = p->head_dim;
    kp.scale = p->softmax_scale;
    kp.scale_log2 = p->softmax_scale * fa::kLog2e;
    kp.softcap = p->softcap;
    kp.window_left = wl;
    kp.window_right = causal ? 0 : wr;  // the causal mask is the right window of width 0
    kp.reverse_m = (kp.window_right >= 0 || kp.window_left >= 0) ? 1 : 0;
    kp.rp_dropout = 1.0f;
    kp.drop_thr = 0xffffffffu;  // keep everything
}

// Keep threshold of the dropout test `Philox word <= thr` (reference include/softmax.h:50-51 computes
// (uint32_t)((1 - p) * 4294967295.0f) in float). The float product is 2^32 when 1 - p rounds to 1 (p < 2^-25),
// which does not fit a uint32_t (undefined behaviour in the reference's expression): clamp it to 2^32 - 1.
uint32_t drop_threshold(float p_dropout) {
    const float t = (1.0f - p_dropout) * 4294967295.0f;
    return t >= 4294967296.0f ? 0xffffffffu : static_cast<uint32_t>(t);
}

// Dropout state (reference include/softmax.h:50-51). `row_elems` = elements of one dmask row group.
int fill_dropout(fa::FwdKernelParams& kp, const fa_b200_params_t* p, bool varlen) {
    CHECK_ARG(p->p_dropout >= 0.f && p->p_dropout < 1.f, "p_dropout must be in [0, 1)");
    if (p->p_dropout == 0.f) {
        CHECK_ARG(p->dmask == nullptr, "return_softmax requires p_dropout > 0");
        return 0;
    }
    CHECK_ARG(p->softcap == 0.f, "Softcapping does not support dropout");
    kp.rp_dropout = 1.0f / (1.0f - p->p_dropout);
    kp.drop_thr = drop_threshold(p->p_dropout);
    kp.drop_seed = p->dropout_seed;
    kp.drop_offset = p->dropout_offset;
    kp.dmask = static_cast<uint16_t*>(p->dmask);
    if (varlen) {  // (total_q, H, max_seqlen_k)   reference ..._varlen.cu:532
        kp.dmask_stride_b = 0;
        kp.dmask_stride_h = p->seqlen_k;
        kp.dmask_stride_row = (int64_t)p->num_heads * p->seqlen_k;
    } else {       // (B, H, Sq, Sk)               reference fused_mha_forward.cu:403
        kp.dmask_stride_row = p->seqlen_k;
        kp.dmask_stride_h = (int64_t)p->seqlen_q * p->seqlen_k;
        kp.dmask_stride_b = (int64_t)p->num_heads * kp.dmask_stride_h;
    }
    return 0;
}

}  // namespace

extern "C" {

#ifdef FA_TRACE
// debug builds only: device buffer of 16 warps x 4096 int64 for the event trace
__attribute__((visibility("default"))) int fa_b200_debug_set_trace(void* dev_ptr) {
    long long* p = static_cast<long long*>(dev_ptr);
    return (int)cudaMemcpyToSymbol(fa::g_fa_trace, &p, sizeof(p));
}
#endif

#ifdef FA_WAIT_LOG
// debug builds only: zero-copy pinned host buffer of 256 blocks x 16 warps x uint64 (ptx_sm100.cuh: fa_wait_log)
__attribute__((visibility("default"))) int fa_b200_debug_set_wait_log(void* host_mapped_ptr) {
    unsigned long long* p = static_cast<unsigned long long*>(host_mapped_ptr);
    return (int)cudaMemcpyToSymbol(fa::g_fa_wait_log, &p, sizeof(p));
}
#endif

// Tests only: device pointer to two zero-initialised 64-bit counters ([0] softmax rows that crossed the lazy-rescale
// threshold, [1] accumulator rescales executed), or NULL to switch counting off. Not part of the drop-in surface.
__attribute__((visibility("default"))) int fa_b200_debug_set_counters(void* dev_ptr) {
    g_dbg_counters.store(static_cast<unsigned long long*>(dev_ptr), std::memory_order_relaxed);
    return 0;
}

FA_B200_API int fa_b200_init(int device) {
    g_err[0] = 0;
    if (int rc = check_device(device)) return rc;
    return init_pools(device);
}

FA_B200_API int fa_b200_abi_version(void) { return FA_B200_ABI_VERSION; }
FA_B200_API const char* fa_b200_last_error(void) { return g_err; }
FA_B200_API int64_t fa_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

FA_B200_API int64_t fa_b200_workspace_bytes(const fa_b200_params_t* p, int kind) {
    if (!p || kind != FA_B200_KIND_KVCACHE) return 0;
    int64_t bytes = 0;
    if (p->rotary_dim > 0)  // rotated copy of q
        bytes += (int64_t)p->batch * p->seqlen_q * p->num_heads * p->head_dim * 2;
    bytes = (bytes + 255) & ~(int64_t)255;
    bytes += decode_workspace_bytes(p);
    return bytes;
}

// ------------------------------------------------------------------------------------------ dense
FA_B200_API int fa_b200_fwd(const fa_b200_params_t* p, void* stream_v) {
    g_err[0] = 0;
    if (int rc = check_common(p)) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    CHECK_ARG(p->seqlen_q > 0, "seqlen_q must be positive");
    CHECK_ARG(p->seqlen_k > 0, "seqlen_k must be positive (the caller handles the empty-KV case)");
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(FA_B200_EINVAL, "cannot select device %d", p->device);

    // reference kernel/fused_mha_forward.cu:343,351-352
    bool causal = p->is_causal != 0;
    if (p->seqlen_q == 1 && !p->alibi_slopes) causal = false;
    int wl = p->window_left, wr = p->window_right;
    if (wl >= p->seqlen_k) wl = -1;
    if (wr >= p->seqlen_k) wr = -1;

    fa::FwdKernelParams kp;
    fill_common(kp, p, causal, wl, wr);
    kp.seqlen_q = p->seqlen_q;
    kp.seqlen_k = p->seqlen_k;
    kp.lse_stride_b = (int64_t)p->num_heads * p->seqlen_q;
    kp.lse_stride_h = p->seqlen_q;
    if (int rc = make_tmap(&kp.tm_q, p->dtype, p->q, p->head_dim, p->num_heads, p->seqlen_q, p->batch, p->q_stride_h, p->q_stride_s, p->q_stride_b, "q")) return rc;
    if (int rc = make_tmap(&kp.tm_k, p->dtype, p->k, p->head_dim, p->num_heads_k, p->seqlen_k, p->batch, p->k_stride_h, p->k_stride_s, p->k_stride_b, "k")) return rc;
    if (int rc = make_tmap(&kp.tm_v, p->dtype, p->v, p->head_dim, p->num_heads_k, p->seqlen_k, p->batch, p->v_stride_h, p->v_stride_s, p->v_stride_b, "v")) return rc;

    if (int rc = fill_dropout(kp, p, false)) return rc;
    const bool feat = p->alibi_slopes != nullptr || p->softcap > 0.f;
    const int v2 = fwd2_cluster(p);
    dim3 grid = set_launch_order(kp, p->device, p->batch, p->num_heads, p->num_heads_k, p->seqlen_q, p->seqlen_k, p->head_dim, v2);
    if (int rc = arm_scheduler(kp, grid, p->device, stream, v2 == 2 ? 2 : 1)) return rc;
    return launch_fwd(kp, p->head_dim, p->dtype, feat, grid, stream, v2);
}

// ------------------------------------------------------------------------------------------ varlen
FA_B200_API int fa_b200_varlen_fwd(const fa_b200_params_t* p, void* stream_v) {
    g_err[0] = 0;
    if (int rc = check_common(p)) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    CHECK_ARG(p->cu_seqlens_q && p->cu_seqlens_k, "cu_seqlens_q and cu_seqlens_k are required");
    CHECK_ARG(p->total_q > 0, "total_q must be positive");
    CHECK_ARG(p->seqlen_q > 0 && p->seqlen_k > 0, "max_seqlen_q / max_seqlen_k must be positive");
    CHECK_ARG(p->num_splits <= 1, "num_splits > 1 is not supported");  // reference ..._varlen.cu:422
    const bool paged = p->block_table != nullptr;
    if (paged) {
        CHECK_ARG(p->page_size > 0 && p->page_size % 128 == 0, "page_block_size must be a multiple of 128");
        CHECK_ARG(p->num_pages > 0, "num_pages must be positive for paged KV");
    } else {
        CHECK_ARG(p->total_k > 0, "total_k must be positive");
    }
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(FA_B200_EINVAL, "cannot select device %d", p->device);

    // reference kernel/fused_mha_forward_varlen.cu:425, 481-482
    bool causal = p->is_causal != 0;
    if (p->seqlen_q == 1 && !p->alibi_slopes) causal = false;
    int wl = p->window_left, wr = p->window_right;
    if (wl >= p->seqlen_k) wl = -1;
    if (wr >= p->seqlen_k) wr = -1;

    fa::FwdKernelParams kp;
    fill_common(kp, p, causal, wl, wr);
    kp.seqlen_q = p->seqlen_q;
    kp.seqlen_k = p->seqlen_k;
    kp.cu_seqlens_q = p->cu_seqlens_q;
    kp.cu_seqlens_k = p->cu_seqlens_k;
    kp.seqused_k = p->seqused_k;
    kp.block_table = p->block_table;
    kp.block_table_stride = p->block_table_stride;
    kp.page_size = p->page_size;
    kp.lse_stride_b = 0;  // lse is [H, total_q] (reference ..._varlen.cu:519)
    kp.lse_stride_h = p->total_q;
    if (int rc = make_tmap(&kp.tm_q, p->dtype, p->q, p->head_dim, p->num_heads, p->total_q, 1, p->q_stride_h, p->q_stride_s, 0, "q")) return rc;
    if (paged) {
        if (int rc = make_tmap(&kp.tm_k, p->dtype, p->k, p->head_dim, p->num_heads_k, p->page_size, p->num_pages, p->k_stride_h, p->k_stride_s, p->k_stride_b, "k")) return rc;
        if (int rc = make_tmap(&kp.tm_v, p->dtype, p->v, p->head_dim, p->num_heads_k, p->page_size, p->num_pages, p->v_stride_h, p->v_stride_s, p->v_stride_b, "v")) return rc;
    } else {
        if (int rc = make_tmap(&kp.tm_k, p->dtype, p->k, p->head_dim, p->num_heads_k, p->total_k, 1, p->k_stride_h, p->k_stride_s, 0, "k")) return rc;
        if (int rc = make_tmap(&kp.tm_v, p->dtype, p->v, p->head_dim, p->num_heads_k, p->total_k, 1, p->v_stride_h, p->v_stride_s, 0, "v")) return rc;
    }
    if (int rc = fill_dropout(kp, p, true)) return rc;
    const bool feat = p->alibi_slopes != nullptr || p->softcap > 0.f;
    const int v2 = fwd2_cluster(p);
    dim3 grid = set_launch_order(kp, p->device, p->batch, p->num_heads, p->num_heads_k, p->seqlen_q, p->seqlen_k, p->head_dim, v2);
    if (int rc = arm_scheduler(kp, grid, p->device, stream, v2 == 2 ? 2 : 1)) return rc;
    return launch_fwd(kp, p->head_dim, p->dtype, feat, grid, stream, v2);
}

// ------------------------------------------------------------------------------------------ kv-cache
FA_B200_API int fa_b200_kvcache_fwd(const fa_b200_params_t* p, void* stream_v) {
    g_err[0] = 0;
    if (int rc = check_common(p)) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    CHECK_ARG(p->seqlen_q > 0, "seqlen_q must be positive");
    CHECK_ARG(p->seqlen_k > 0, "cache capacity (seqlen_k) must be positive");
    const bool paged = p->block_table != nullptr;
    const bool has_new = p->k_new != nullptr;
    const bool rotary = p->rotary_dim > 0;
    // reference kernel/fused_mha_forward_kvcache.cu:469-472, 480, 506-509, 556, 569-594
    CHECK_ARG((p->k_new != nullptr) == (p->v_new != nullptr), "k and v must be given together");
    CHECK_ARG(!has_new || p->cache_seqlens, "cache_seqlens is required when k/v are appended");
    CHECK_ARG(!has_new || p->seqlen_new > 0, "seqlen_new must be positive when k/v are appended");
    if (p->softcap > 0.f) {
        CHECK_ARG(p->window_left < 0 && p->window_right < 0, "softcap does not support a sliding window in the kv-cache path");
        CHECK_ARG(!p->alibi_slopes, "softcap does not support ALiBi in the kv-cache path");
    }
    if (paged) {
        CHECK_ARG(!p->cache_batch_idx, "paged KV does not support cache_batch_idx");
        CHECK_ARG(!p->cache_leftpad, "paged KV does not support cache_leftpad");
        CHECK_ARG(p->page_size > 0 && p->page_size % 128 == 0, "page_block_size must be a multiple of 128");
        CHECK_ARG(p->num_pages > 0, "num_pages must be positive for paged KV");
    } else {
        CHECK_ARG(p->batch_k > 0, "batch_k (cache batch size) must be positive");
    }
    if (rotary) {
        CHECK_ARG(has_new, "rotary embedding requires k/v to append");
        CHECK_ARG(p->rotary_cos && p->rotary_sin, "rotary_cos and rotary_sin must both be given");
        CHECK_ARG(p->rotary_dim <= p->head_dim, "rotary_dim must be <= head_dim");
        CHECK_ARG(p->rotary_dim % 16 == 0, "rotary_dim must be a multiple of 16");
        CHECK_ARG(p->rotary_seqlen >= p->seqlen_k, "rotary cos/sin must cover the cache length");
        CHECK_ARG(p->workspace, "workspace is required when rotary is used (see fa_b200_workspace_bytes)");
    }
    CHECK_ARG(p->num_splits >= 0, "num_splits must be >= 0");
    CHECK_ARG(p->p_dropout == 0.f && p->dmask == nullptr, "the kv-cache path has no dropout");
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(FA_B200_EINVAL, "cannot select device %d", p->device);

    // reference ..._kvcache.cu:465-466, 597-598
    bool causal = p->is_causal != 0;
    if (p->seqlen_q == 1 && !p->alibi_slopes) causal = false;
    int wl = p->window_left, wr = p->window_right;
    if (causal) wr = 0;
    if (wl >= p->seqlen_k) wl = -1;
    if (wr >= p->seqlen_k) wr = -1;
    const bool bf16 = p->dtype == FA_B200_DTYPE_BF16;

    // 1. append (+ RoPE on K), once per kv head; 2. RoPE on Q into the workspace -- ONE launch when both are needed
    fa::AppendParams ap;
    memset(&ap, 0, sizeof(ap));
    int64_t prep_total = 0;
    if (has_new) {
        ap.k_new = static_cast<const uint16_t*>(p->k_new);
        ap.v_new = static_cast<const uint16_t*>(p->v_new);
        ap.k_cache = static_cast<uint16_t*>(const_cast<void*>(p->k));
        ap.v_cache = static_cast<uint16_t*>(const_cast<void*>(p->v));
        ap.cos = static_cast<const uint16_t*>(p->rotary_cos);
        ap.sin = static_cast<const uint16_t*>(p->rotary_sin);
        ap.cache_seqlens = p->cache_seqlens;
        ap.cache_batch_idx = p->cache_batch_idx;
        ap.leftpad = p->cache_leftpad;
        ap.block_table = p->block_table;
        ap.knew_sb = p->knew_stride_b; ap.knew_ss = p->knew_stride_s; ap.knew_sh = p->knew_stride_h;
        ap.vnew_sb = p->vnew_stride_b; ap.vnew_ss = p->vnew_stride_s; ap.vnew_sh = p->vnew_stride_h;
        ap.kc_sb = p->k_stride_b; ap.kc_ss = p->k_stride_s; ap.kc_sh = p->k_stride_h;
        ap.vc_sb = p->v_stride_b; ap.vc_ss = p->v_stride_s; ap.vc_sh = p->v_stride_h;
        ap.batch = p->batch; ap.seqlen_new = p->seqlen_new; ap.heads_k = p->num_heads_k; ap.head_dim = p->head_dim;
        ap.rotary_dim = rotary ? p->rotary_dim : 0;
        ap.interleaved = p->rotary_interleaved;
        ap.block_table_stride = p->block_table_stride;
        ap.page_size = p->page_size;
        prep_total = (int64_t)p->batch * p->seqlen_new * p->num_heads_k * (p->head_dim / 8);
    }
    const void* q_ptr = p->q;
    int64_t q_sb = p->q_stride_b, q_ss = p->q_stride_s, q_sh = p->q_stride_h;
    char* ws = static_cast<char*>(p->workspace);
    int64_t ws_left = p->workspace_bytes;
    if (rotary) {
        const int64_t qbytes = (((int64_t)p->batch * p->seqlen_q * p->num_heads * p->head_dim * 2) + 255) & ~(int64_t)255;
        CHECK_ARG(ws_left >= qbytes, "workspace too small for the rotated q (%lld < %lld)", (long long)ws_left, (long long)qbytes);
        fa::QRotaryParams rp;
        memset(&rp, 0, sizeof(rp));
        rp.q = static_cast<const uint16_t*>(p->q);
        rp.q_out = reinterpret_cast<uint16_t*>(ws);
        rp.cos = static_cast<const uint16_t*>(p->rotary_cos);
        rp.sin = static_cast<const uint16_t*>(p->rotary_sin);
        rp.cache_seqlens = p->cache_seqlens;
        rp.leftpad = p->cache_leftpad;
        rp.q_sb = q_sb; rp.q_ss = q_ss; rp.q_sh = q_sh;
        rp.batch = p->batch; rp.seqlen_q = p->seqlen_q; rp.heads = p->num_heads; rp.head_dim = p->head_dim;
        rp.rotary_dim = p->rotary_dim;
        rp.interleaved = p->rotary_interleaved;
        rp.per_row_pos = (causal || wl >= 0 || wr >= 0) ? 1 : 0;
        const int64_t total = (int64_t)p->batch * p->seqlen_q * p->num_heads * (p->head_dim / 8);
        auto blocks_for = [](int64_t n) { return (int)((n + 127) / 128 < 148 * 4 ? (n + 127) / 128 : 148 * 4); };
        const int append_blocks = blocks_for(prep_total), q_blocks = blocks_for(total);
        // rotary implies has_new (checked above): append + RoPE(K) + RoPE(Q) in one launch, on disjoint CTAs
        if (bf16) fa::kv_prep_kernel<true><<<append_blocks + q_blocks, 128, 0, stream>>>(ap, rp, append_blocks);
        else fa::kv_prep_kernel<false><<<append_blocks + q_blocks, 128, 0, stream>>>(ap, rp, append_blocks);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "kv_prep_kernel launch");
        q_ptr = ws;
        q_sh = p->head_dim;
        q_ss = (int64_t)p->num_heads * p->head_dim;
        q_sb = (int64_t)p->seqlen_q * q_ss;
        ws += qbytes;
        ws_left -= qbytes;
    } else if (has_new) {
        const int blocks = (int)((prep_total + 255) / 256 < 148 * 8 ? (prep_total + 255) / 256 : 148 * 8);
        if (bf16) fa::kv_append_kernel<true><<<blocks, 256, 0, stream>>>(ap);
        else fa::kv_append_kernel<false><<<blocks, 256, 0, stream>>>(ap);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "kv_append_kernel launch");
    }

    // 3. attention over the cache: the tcgen05 kernel reading the (paged) cache through TMA
    fa::FwdKernelParams kp;
    fill_common(kp, p, causal, wl, wr);
    const bool decode = decode_applies(p);
    kp.seqlen_q = p->seqlen_q;
    kp.seqlen_k = p->seqlen_k;
    kp.cache_seqlens = p->cache_seqlens;
    kp.seqlen_k_add = has_new ? p->seqlen_new : 0;
    kp.cache_batch_idx = p->cache_batch_idx;
    kp.leftpad_k = p->cache_leftpad;
    kp.block_table = p->block_table;
    kp.block_table_stride = p->block_table_stride;
    kp.page_size = p->page_size;
    kp.lse_stride_b = (int64_t)p->num_heads * p->seqlen_q;
    kp.lse_stride_h = p->seqlen_q;
    const int G = p->num_heads / p->num_heads_k;
    if (int rc = make_tmap(&kp.tm_q, p->dtype, q_ptr, p->head_dim, p->num_heads, p->seqlen_q, p->batch, q_sh, q_ss, q_sb, "q",
                           decode ? G : 1, decode ? 128 / G : 128)) return rc;
    const int64_t rows = paged ? p->page_size : p->seqlen_k;
    const int64_t nb = paged ? p->num_pages : p->batch_k;
    if (int rc = make_tmap(&kp.tm_k, p->dtype, p->k, p->head_dim, p->num_heads_k, rows, nb, p->k_stride_h, p->k_stride_s, p->k_stride_b, "k_cache")) return rc;
    if (int rc = make_tmap(&kp.tm_v, p->dtype, p->v, p->head_dim, p->num_heads_k, rows, nb, p->v_stride_h, p->v_stride_s, p->v_stride_b, "v_cache")) return rc;
    if (decode) {
        // few query rows: pack the GQA group into the tile rows and split the KV length over CTAs
        kp.gqa_pack = G;
        kp.num_splits = decode_num_splits(p);
        if (kp.num_splits > 1) {
            const int64_t prow = (int64_t)kp.num_splits * p->batch * p->num_heads * p->seqlen_q;
            const int64_t obytes = (prow * tile_dim(p->head_dim) * 4 + 255) & ~(int64_t)255;
            const int64_t lbytes = (prow * 4 + 255) & ~(int64_t)255;
            CHECK_ARG(ws != nullptr && ws_left >= obytes + lbytes, "workspace too small for %d KV splits (see fa_b200_workspace_bytes)", kp.num_splits);
            kp.o_partial = reinterpret_cast<float*>(ws);
            kp.lse_partial = reinterpret_cast<float*>(ws + obytes);
        }
        dim3 grid(kp.num_splits, p->num_heads_k, p->batch);
        // programmatic dependent launch behind the preparation kernel (when there is one)
        if (int rc = launch_decode(kp, p->head_dim, p->dtype, grid, stream, has_new)) return rc;
        if (kp.num_splits > 1) return launch_combine(kp, p, stream);
        return 0;
    }
    const bool feat = p->alibi_slopes != nullptr || p->softcap > 0.f;
    const int v2 = fwd2_cluster(p);
    dim3 grid = set_launch_order(kp, p->device, p->batch, p->num_heads, p->num_heads_k, p->seqlen_q, p->seqlen_k, p->head_dim, v2);
    if (int rc = arm_scheduler(kp, grid, p->device, stream, v2 == 2 ? 2 : 1)) return rc;
    return launch_fwd(kp, p->head_dim, p->dtype, feat, grid, stream, v2);
}

// ------------------------------------------------------------------------------------------ backward
}  // extern "C"

namespace {

template <int D, bool BF16, bool FEAT, bool KV_STAT, bool DROPOUT>
int launch_bwd_t(const fa::BwdKernelParams& kp, dim3 grid, cudaStream_t stream) {
    using Cfg = fa::BwdConfig<D>;
    auto kern = fa::fa_bwd_sm100_kernel<D, BF16, FEAT, KV_STAT, DROPOUT>;
    static std::once_flag once[64];
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t attr_err = cudaSuccess;
    std::call_once(once[dev & 63], [&] {
        attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    });
    if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(smem)");
    kern<<<grid, 384, Cfg::kSmemBytes, stream>>>(kp);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "fa_bwd_sm100_kernel launch");
    return 0;
}

template <bool KV_STAT>
int launch_bwd(const fa::BwdKernelParams& kp, int real_head_dim, bool bf16, bool feat, bool dropout, dim3 grid, cudaStream_t stream) {
    const int head_dim = tile_dim(real_head_dim);
    // dropout variants are built with the score-modifier path compiled in (one variant per D and dtype)
#define FA_BCASE(DD, BB)                                                                                       \
    if (head_dim == DD && bf16 == BB) {                                                                        \
        if (dropout) return launch_bwd_t<DD, BB, true, KV_STAT, true>(kp, grid, stream);                       \
        if (feat) return launch_bwd_t<DD, BB, true, KV_STAT, false>(kp, grid, stream);                         \
        return launch_bwd_t<DD, BB, false, KV_STAT, false>(kp, grid, stream);                                  \
    }
    FA_BCASE(128, true)
    FA_BCASE(128, false)
    FA_BCASE(64, true)
    FA_BCASE(64, false)
    FA_BCASE(256, true)
    FA_BCASE(256, false)
#undef FA_BCASE
    return fail(FA_B200_EUNSUPPORTED, "head_dim %d is not built", head_dim);
}

int bwd_common(const fa_b200_params_t* p, void* stream_v, bool varlen) {
    g_err[0] = 0;
    if (int rc = check_common(p)) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    // reference kernel/fused_mha_backward.cu:604-700
    CHECK_ARG(p->dout && p->dq && p->dk && p->dv && p->softmax_d, "dout, dq, dk, dv and softmax_d must be non-NULL");
    CHECK_ARG(p->seqlen_q > 0 && p->seqlen_k > 0, "seqlen_q / seqlen_k must be positive (the caller handles empty inputs)");
    CHECK_ARG(p->p_dropout >= 0.f && p->p_dropout < 1.f, "p_dropout must be in [0, 1)");
    CHECK_ARG(p->softcap == 0.f || p->p_dropout == 0.f, "Softcapping does not support dropout");
    CHECK_ARG(p->block_table == nullptr, "the backward has no paged-KV form");
    auto aligned16 = [](const void* ptr, int64_t a, int64_t b, int64_t c) {
        return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && a % 8 == 0 && b % 8 == 0 && c % 8 == 0;
    };
    CHECK_ARG(aligned16(p->dq, p->dq_stride_b, p->dq_stride_s, p->dq_stride_h) &&
                  aligned16(p->dk, p->dk_stride_b, p->dk_stride_s, p->dk_stride_h) &&
                  aligned16(p->dv, p->dv_stride_b, p->dv_stride_s, p->dv_stride_h),
              "dq, dk, dv must be 16-byte aligned with strides that are multiples of 8 elements");
    CHECK_ARG(aligned16(p->out, p->o_stride_b, p->o_stride_s, p->o_stride_h) &&
                  aligned16(p->dout, p->do_stride_b, p->do_stride_s, p->do_stride_h),
              "out and dout must be 16-byte aligned with strides that are multiples of 8 elements");
    if (varlen) {
        CHECK_ARG(p->cu_seqlens_q && p->cu_seqlens_k, "cu_seqlens_q and cu_seqlens_k are required");
        CHECK_ARG(p->total_q > 0 && p->total_k > 0, "total_q and total_k must be positive");
    }
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(FA_B200_EINVAL, "cannot select device %d", p->device);

    // the same normalisations as the forward (reference kernel/fused_mha_backward.cu:636-641)
    bool causal = p->is_causal != 0;
    if (p->seqlen_q == 1 && !p->alibi_slopes) causal = false;
    int wl = p->window_left, wr = p->window_right;
    if (wl >= p->seqlen_k) wl = -1;
    if (wr >= p->seqlen_k) wr = -1;
    if (causal) wr = 0;
    const bool bf16 = p->dtype == FA_B200_DTYPE_BF16;

    fa::BwdKernelParams kp;
    memset(&kp, 0, sizeof(kp));
    const int64_t rows_q = varlen ? p->total_q : p->seqlen_q;
    const int64_t rows_k = varlen ? p->total_k : p->seqlen_k;
    const int64_t nb = varlen ? 1 : p->batch;
    // head_dim 256 streams 64-row tiles (BwdConfig<256>::kBTS); the stationary side always uses 128-row boxes
    const bool split_d = tile_dim(p->head_dim) == 256;
    const int stream_rows = split_d ? 64 : 128;
    auto map_q = [&](CUtensorMap* tm, int box_rows) { return make_tmap(tm, p->dtype, p->q, p->head_dim, p->num_heads, rows_q, nb, p->q_stride_h, p->q_stride_s, varlen ? 0 : p->q_stride_b, "q", 1, box_rows); };
    auto map_do = [&](CUtensorMap* tm, int box_rows) { return make_tmap(tm, p->dtype, p->dout, p->head_dim, p->num_heads, rows_q, nb, p->do_stride_h, p->do_stride_s, varlen ? 0 : p->do_stride_b, "dout", 1, box_rows); };
    auto map_k = [&](CUtensorMap* tm, int box_rows) { return make_tmap(tm, p->dtype, p->k, p->head_dim, p->num_heads_k, rows_k, nb, p->k_stride_h, p->k_stride_s, varlen ? 0 : p->k_stride_b, "k", 1, box_rows); };
    auto map_v = [&](CUtensorMap* tm, int box_rows) { return make_tmap(tm, p->dtype, p->v, p->head_dim, p->num_heads_k, rows_k, nb, p->v_stride_h, p->v_stride_s, varlen ? 0 : p->v_stride_b, "v", 1, box_rows); };
    // dK/dV pass: K, V stationary; Q, dO streamed
    if (int rc = map_q(&kp.tm_q, stream_rows)) return rc;
    if (int rc = map_do(&kp.tm_do, stream_rows)) return rc;
    if (int rc = map_k(&kp.tm_k, 128)) return rc;
    if (int rc = map_v(&kp.tm_v, 128)) return rc;
    kp.dq = p->dq; kp.dk = p->dk; kp.dv = p->dv;
    kp.dq_stride_b = p->dq_stride_b; kp.dq_stride_s = p->dq_stride_s; kp.dq_stride_h = p->dq_stride_h;
    kp.dk_stride_b = p->dk_stride_b; kp.dk_stride_s = p->dk_stride_s; kp.dk_stride_h = p->dk_stride_h;
    kp.dv_stride_b = p->dv_stride_b; kp.dv_stride_s = p->dv_stride_s; kp.dv_stride_h = p->dv_stride_h;
    kp.lse = p->lse;
    kp.delta = p->softmax_d;
    kp.lse_stride_b = varlen ? 0 : (int64_t)p->num_heads * p->seqlen_q;
    kp.lse_stride_h = varlen ? p->total_q : p->seqlen_q;
    kp.cu_seqlens_q = varlen ? p->cu_seqlens_q : nullptr;
    kp.cu_seqlens_k = varlen ? p->cu_seqlens_k : nullptr;
    kp.alibi = p->alibi_slopes;
    kp.alibi_stride_b = p->alibi_stride_b;
    kp.seqlen_q = p->seqlen_q;
    kp.seqlen_k = p->seqlen_k;
    kp.num_heads = p->num_heads;
    kp.heads_per_kv = p->num_heads / p->num_heads_k;
    kp.head_dim = p->head_dim;
    kp.scale = p->softmax_scale;
    kp.scale_log2 = p->softmax_scale * fa::kLog2e;
    kp.softcap = p->softcap;
    kp.window_left = wl;
    kp.window_right = wr;
    kp.rp_dropout = 1.0f;
    kp.drop_thr = 0xffffffffu;
    const bool dropout = p->p_dropout > 0.f;
    if (dropout) {
        kp.rp_dropout = 1.0f / (1.0f - p->p_dropout);
        kp.drop_thr = drop_threshold(p->p_dropout);
        kp.drop_seed = p->dropout_seed;
        kp.drop_offset = p->dropout_offset;
    }
    const bool feat = p->alibi_slopes != nullptr || p->softcap > 0.f;

    // 1. delta = rowsum(dO * O)   (reference include/product.h; returned as softmax_d)
    {
        const int64_t rows = (int64_t)nb * rows_q * p->num_heads;
        int lpr = 1;  // lanes per row: head_dim / 8 rounded up to a power of two
        while (lpr * 8 < p->head_dim) lpr <<= 1;
        const int warps = 8;
        const int64_t rows_per_block = (int64_t)warps * (32 / lpr) * fa::kDotUnroll;
        const dim3 grid((unsigned)((rows + rows_per_block - 1) / rows_per_block));
        const uint16_t* o = static_cast<const uint16_t*>(p->out);
        const uint16_t* d = static_cast<const uint16_t*>(p->dout);
        if (bf16)
            fa::fa_bwd_dot_kernel<true><<<grid, warps * 32, 0, stream>>>(o, d, p->softmax_d, p->head_dim, rows, (int)rows_q, p->num_heads,
                varlen ? 0 : p->o_stride_b, p->o_stride_s, p->o_stride_h, varlen ? 0 : p->do_stride_b, p->do_stride_s, p->do_stride_h,
                kp.lse_stride_b, kp.lse_stride_h, lpr);
        else
            fa::fa_bwd_dot_kernel<false><<<grid, warps * 32, 0, stream>>>(o, d, p->softmax_d, p->head_dim, rows, (int)rows_q, p->num_heads,
                varlen ? 0 : p->o_stride_b, p->o_stride_s, p->o_stride_h, varlen ? 0 : p->do_stride_b, p->do_stride_s, p->do_stride_h,
                kp.lse_stride_b, kp.lse_stride_h, lpr);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "fa_bwd_dot_kernel launch");
    }
    // 1-D grid in sectioned longest-first order (BwdKernelParams::section_bh): a section is the (head, batch) pairs
    // whose two streamed tensors (`stream_rows` rows per grid head) fit ~32 MB of L2.
    // Sections wider than one pair only pay where block lengths differ a lot: the dK/dV pass with GQA, whose blocks
    // stream the whole group (measured: +4.5 % there, -1.5..-3.5 % on uniform work from the lost L2 sharing between
    // the blocks of one head). `wide = false` is the plain order: all blocks of a pair, then the next pair.
    auto set_order = [&](int blocks, int grid_heads, int64_t stream_rows_per_head, bool wide) {
        kp.num_blocks = blocks;
        kp.grid_heads = grid_heads;
        kp.num_bh = grid_heads * p->batch;
        const int64_t bytes = 2ll * stream_rows_per_head * p->head_dim * 2;
        int64_t sec = (32ll << 20) / (bytes > 0 ? bytes : 1);
        if (sec < 1 || !wide) sec = 1;
        if (sec > kp.num_bh) sec = kp.num_bh;
        kp.section_bh = (int)sec;
        return dim3((unsigned)((int64_t)blocks * kp.num_bh * (split_d ? 2 : 1)), 1, 1);
    };
    CHECK_ARG(((int64_t)(p->seqlen_k + 127) / 128 + (p->seqlen_q + 127) / 128) * p->batch * p->num_heads * 2 < (1ll << 31),
              "too many (block, head, batch) work items for one launch");
    // 2. dK / dV pass: one CTA per (128-key block, KV head, batch)
    {
        kp.reverse = 0;
        dim3 grid = set_order((p->seqlen_k + 127) / 128, p->num_heads_k, (int64_t)p->seqlen_q * kp.heads_per_kv,
                              kp.heads_per_kv > 1 && (wl >= 0 || wr >= 0));
        if (int rc = launch_bwd<true>(kp, p->head_dim, bf16, feat, dropout, grid, stream)) return rc;
    }
    // 3. dQ pass: one CTA per (128-query block, head, batch)
    {
        kp.reverse = wr >= 0 ? 1 : 0;
        if (split_d) {  // dQ pass: Q, dO stationary (128-row boxes); K, V streamed in 64-row tiles
            if (int rc = map_q(&kp.tm_q_stat, 128)) return rc;
            if (int rc = map_do(&kp.tm_do_stat, 128)) return rc;
            if (int rc = map_k(&kp.tm_k, stream_rows)) return rc;
            if (int rc = map_v(&kp.tm_v, stream_rows)) return rc;
        }
        dim3 grid = set_order((p->seqlen_q + 127) / 128, p->num_heads, p->seqlen_k, false);
        if (int rc = launch_bwd<false>(kp, p->head_dim, bf16, feat, dropout, grid, stream)) return rc;
    }
    return 0;
}

}  // namespace

extern "C" {

FA_B200_API int fa_b200_bwd(const fa_b200_params_t* p, void* stream_v) { return bwd_common(p, stream_v, false); }
FA_B200_API int fa_b200_varlen_bwd(const fa_b200_params_t* p, void* stream_v) { return bwd_common(p, stream_v, true); }

}  // extern "C"
