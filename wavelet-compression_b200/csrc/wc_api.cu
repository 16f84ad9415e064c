// libwcgpu C ABI (include/wcgpu.h): contexts, plans, buffers and the launch sequences.
// No CPU fallback anywhere in this file: every numeric result comes from a CUDA kernel.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "wc_common.cuh"
#include "wc_fused.h"
#include "wc_kernels.h"

using namespace wc;

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
namespace {

struct DevBuf {
    void*  p   = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p   = nullptr;
        cap = 0;
        size_t want = (bytes + 255) & ~size_t(255);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p   = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

struct PinBuf {
    void*  p   = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p   = nullptr;
        cap = 0;
        size_t want = (bytes + 4095) & ~size_t(4095);
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p   = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline size_t dtype_size(int dt) { return dt == WC_F64 ? 8 : 4; }

struct CopyRange {
    char*       dst;
    const char* src;
    size_t      bytes;
};

// merges copies whose source and destination are both contiguous with the previous one
struct CopyList {
    std::vector<CopyRange> r;
    void add(void* dst, const void* src, size_t bytes) {
        if (bytes == 0) return;
        if (!r.empty()) {
            CopyRange& b = r.back();
            if (b.dst + b.bytes == (char*)dst && b.src + b.bytes == (const char*)src) {
                b.bytes += bytes;
                return;
            }
        }
        r.push_back({ (char*)dst, (const char*)src, bytes });
    }
};

} // namespace

struct wc_ctx {
    int          device     = 0;
    cudaStream_t stream     = nullptr;
    bool         own_stream = false;
    std::string  last_error;
    LaunchStats  ls;
    uint64_t     h2d = 0, d2h = 0;
    int          opt_path = 0;
    int          opt_overlap = 1;
    int          opt_seg_index = 0;   // 0 = streamed k_seg_index3 (TMA ring), 1 = k_seg_index (direct loads)
    int          opt_copy_only = 0;   // probe: wc_plan_compress_to_host moves the bytes but skips the kernels
    int          opt_ingest_stats = 0; // compress also records per-unit min / max of the narrowed inputs
    int          opt_decode_pipe = 2;  // 32^3 cubes decode from the TMA-fed staging area (2: also when a table came along)
    int          sm_count = 0;
    wc_plan*     batch_plan = nullptr; // owner of the memory handed out by wc_compress_batch
    // workspace of the blocking decompress / rmse / primitive calls (grow-only)
    DevBuf ws_pairs, ws_coef, ws_boxes, ws_tbl0, ws_tbl1, ws_tbl2, ws_tiles0, ws_tiles1, ws_sum,
        ws_misc, ws_a, ws_b;
    PinBuf ws_pin;
    // second stream for running the single-CTA and cluster kernels of one step concurrently
    cudaStream_t s_aux = nullptr;
    cudaEvent_t  ev_fork = nullptr, ev_join = nullptr;
    DevBuf       d_counter;
    unsigned     counter_next = 0;
};

// Device tables of a plan's last wc_plan_decompress call (see run_decompress / relaunch_decompress).
struct DecCache {
    uint64_t h1 = 0, h2 = 0;      // 128-bit hash of (output space, pointers, dtypes) of the cached call
    bool     valid = false;
    size_t   fl_n[16] = {};
    // copy of the cached call's output descriptors (device outputs only): an identical array re-launches without the
    // per-unit validation and hashing, which cost ~5 ns per unit and made the round trip of a million 8^3 boxes host-bound
    std::vector<wc_box_out> last_out;
};
// Fused compress classes (wc_fused.h: fused_class) in launch order: cluster kernels first (they go on the ctx
// stream), then the single-CTA kernels (second stream when both kinds are present).
enum { FL_N = 15 };
static const int FL_CLASS[FL_N] = {FUSED_CLS_CUBE64, FUSED_CLS_R8, FUSED_CLS_R4, FUSED_CLS_R2, FUSED_CLS_CUBE32,
                                    FUSED_CLS_R1, FUSED_CLS_CUBE16, FUSED_CLS_R1S, FUSED_CLS_CUBE8,
                                    FUSED_CLS_RBIG /* decompress only; compress plans never fill it */,
                                    FUSED_CLS_XS1, FUSED_CLS_XS2, FUSED_CLS_XS4, FUSED_CLS_XS8, FUSED_CLS_XS1S /* any shape: wc_xslab.cu */};
static_assert(sizeof(DecCache::fl_n) / sizeof(size_t) >= FL_N, "DecCache::fl_n");
enum { FL_BIG = 9 };
// One launch takes units of one launch key (slab count, 128^3 cube or not: big_run_key): the FUSED_CLS_RBIG list is kept
// sorted by it (then by unit id) and launched
// run by run; every other class is a single run.  f(first, count, slabs)
template <class Dims, class F>
static int for_each_run(int cls, const std::vector<int>& list, Dims dims, F f) {
    if (cls != FUSED_CLS_RBIG) return list.empty() ? WC_OK : f((size_t)0, list.size(), 0);
    size_t a = 0;
    while (a < list.size()) {
        int nx, ny, nz;
        dims(list[a], nx, ny, nz);
        const int S = big_run_key(nx, ny, nz);
        size_t b = a + 1;
        for (; b < list.size(); ++b) {
            dims(list[b], nx, ny, nz);
            if (big_run_key(nx, ny, nz) != S) break;
        }
        int rc = f(a, b - a, S);
        if (rc != WC_OK) return rc;
        a = b;
    }
    return WC_OK;
}
template <class Dims>
static void sort_big_list(std::vector<int>& list, Dims dims) {
    std::stable_sort(list.begin(), list.end(), [&](int x, int y) {
        int ax, ay, az, bx, by, bz;
        dims(x, ax, ay, az); dims(y, bx, by, bz);
        return big_run_key(ax, ay, az) < big_run_key(bx, by, bz);
    });
}
static inline bool fl_is_cluster(int k) { return k < 4; }

struct wc_plan {
    wc_ctx* ctx      = nullptr;
    int     n_units  = 0;
    int     in_space = WC_DEVICE;
    std::vector<wc_box_desc> units;
    std::vector<UnitDev>     h_units;
    // unit ids per path: fl[k] = the fused class FL_CLASS[k] (ascending ids), generic = the rest
    std::vector<int>         fl[FL_N], generic;
    std::vector<int>         bigfwd;     // generic units whose forward transform runs by y-slabs (k_big_forward), sorted by slab count
    DevBuf                   d_bigfwd;
    int                      n_pk_items = 0;   // (unit, chunk) items of the one-pass packing kernel, every generic unit
    DevBuf                   d_pk_items, d_pk_status;
    // quantile thresholds (extension): radix-select state and histograms, one row per unit or one for the batch
    DevBuf                   d_q, d_qhist, d_qrank;
    int                      q_global = -1;    // -1: no quantile run in progress; 0 / 1: mode of the run begun
    double                   q_keep = 0.0;
    std::vector<char>        has_segtab;                     // per unit: UnitDev::coef is a segment table
    long long total_n = 0;     // sum of ncoef
    size_t    in_bytes = 0;    // sum of input bytes
    // device memory
    DevBuf d_units, d_states, d_in, d_out, d_coef, d_xtiles, d_ctiles, d_tile_i, d_gkey,
        d_offsets, d_dense, d_fl[FL_N], d_dec_units, d_inv_units, d_inv_tiles, d_ptiles, d_psum,
        d_err, d_rmse_units, d_rmse_sum, d_rmse, d_stage_out;
    PinBuf h_states, h_dense, h_misc;
    int    n_xtiles = 0, n_ctiles = 0;
    bool   compressed = false;
    bool   transformed = false;
    bool   has_stats = false;    // the last compress ran with WC_OPT_INGEST_STATS: UnitState::vmin / vmax are valid
    std::vector<size_t> in_dev_off; // per unit offset in d_in (host inputs)
    // pipelined host path (wc_plan_compress_to_host)
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_aux = nullptr;
    cudaEvent_t  ev_fork = nullptr, ev_join = nullptr;
    DevBuf d_counter;
    unsigned counter_next = 0;
    DevBuf d_rmse_tiles;
    long long rmse_tiles = -1;   // tiles resident in d_rmse_tiles (-1: not built yet)
    DevBuf d_dec_list;           // fused work lists of the last wc_plan_decompress (re-used on a cache hit)
    DecCache* dec_cache = nullptr;
    std::vector<cudaEvent_t> ev, ev_d2h;
    DevBuf d_running;
    // wc_plan_set_inputs: ring of pinned pointer tables (host) + one device table, patched by k_patch_inputs
    enum { IN_RING = 4 };
    PinBuf      h_inptr[IN_RING];
    cudaEvent_t ev_inptr[IN_RING] = {}, ev_patch[IN_RING] = {};
    DevBuf      d_inptr[IN_RING];
    cudaStream_t s_inptr = nullptr;      // side stream of the pointer-table copies
    unsigned    inptr_next = 0;
};

#define CTX_CUDA(ctx, call)                                                                        \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            (ctx)->last_error = std::string(#call) + ": " + cudaGetErrorString(e__);               \
            cudaGetLastError();                                                                    \
            return e__ == cudaErrorMemoryAllocation ? WC_ERR_OOM : WC_ERR_CUDA;                    \
        }                                                                                          \
    } while (0)

static int check_dims(int nx, int ny, int nz) {
    if (nx < 0 || ny < 0 || nz < 0) return WC_ERR_BAD_DIMS;
    long long n = (long long)nx * ny * nz;
    if (n >= (1ll << 31) - 1) return WC_ERR_BAD_DIMS;
    return WC_OK;
}

// ---------------------------------------------------------------------------------------------
// library / context
// ---------------------------------------------------------------------------------------------
extern "C" {

int wc_version(void) { return WCGPU_VERSION; }

const char* wc_strerror(int status) {
    switch (status) {
    case WC_OK: return "ok";
    case WC_ERR_INVALID_ARG: return "invalid argument";
    case WC_ERR_BAD_DIMS: return "bad box dimensions";
    case WC_ERR_NO_DEVICE: return "no usable CUDA device (libwcgpu has no CPU fallback)";
    case WC_ERR_CUDA: return "CUDA error (see wc_last_error)";
    case WC_ERR_OOM: return "out of device or pinned host memory";
    case WC_ERR_CAPACITY: return "output buffer too small";
    case WC_ERR_CORRUPT: return "corrupt packed stream";
    case WC_ERR_STATE: return "call order violated";
    default: return "unknown status";
    }
}

const char* wc_box_kernel_class(int nx, int ny, int nz, int dtype, int decompress) {
    if (nx < 0 || ny < 0 || nz < 0 || (dtype != WC_F32 && dtype != WC_F64)) return "invalid";
    if ((long long)nx * ny * nz == 0) return "empty";
    const void* aligned = reinterpret_cast<const void*>(static_cast<uintptr_t>(256));
    const int cls = decompress ? fused_decode_class(nx, ny, nz, dtype, aligned) : fused_class(nx, ny, nz, dtype, aligned);
    switch (cls) {
    case FUSED_CLS_CUBE8: return "cube8";
    case FUSED_CLS_CUBE16: return "cube16";
    case FUSED_CLS_CUBE32: return "cube32";
    case FUSED_CLS_CUBE64: return "cube64";
    case FUSED_CLS_R1S: return "r1s";
    case FUSED_CLS_R1: return "r1";
    case FUSED_CLS_R2: return "r2";
    case FUSED_CLS_R4: return "r4";
    case FUSED_CLS_R8: return "r8";
    case FUSED_CLS_RBIG: return "yslab";
    case FUSED_CLS_XS1S: return "xs1s";
    case FUSED_CLS_XS1: return "xs1";
    case FUSED_CLS_XS2: return "xs2";
    case FUSED_CLS_XS4: return "xs4";
    case FUSED_CLS_XS8: return "xs8";
    }
    // compress: boxes no cluster holds run their forward transform by y-slabs into the coefficient scratch
    if (!decompress && big_forward_slabs(nx, ny, nz, dtype, aligned)) return "yslab";
    return "generic";
}

int wc_device_count(int* count) {
    if (!count) return WC_ERR_INVALID_ARG;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return WC_ERR_NO_DEVICE;
    }
    *count = n;
    return WC_OK;
}

static int create_impl(wc_ctx** out, int device_id, void* stream, bool use_given) {
    if (!out) return WC_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return WC_ERR_NO_DEVICE;
    }
    if (device_id < 0 || device_id >= n) return WC_ERR_INVALID_ARG;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) {
        cudaGetLastError();
        return WC_ERR_NO_DEVICE;
    }
    if (prop.major != 10) return WC_ERR_NO_DEVICE; // the only code in this library is sm_100a SASS
    wc_ctx* c = new (std::nothrow) wc_ctx();
    if (!c) return WC_ERR_OOM;
    c->device   = device_id;
    c->sm_count = prop.multiProcessorCount;
    if (cudaSetDevice(device_id) != cudaSuccess) {
        cudaGetLastError();
        delete c;
        return WC_ERR_NO_DEVICE;
    }
    if (use_given) {
        c->stream     = static_cast<cudaStream_t>(stream);
        c->own_stream = false;
    } else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            delete c;
            return WC_ERR_CUDA;
        }
        c->own_stream = true;
    }
    *out = c;
    return WC_OK;
}

int wc_create(wc_ctx** ctx, int device_id) { return create_impl(ctx, device_id, nullptr, false); }
int wc_create_on_stream(wc_ctx** ctx, int device_id, void* cuda_stream) {
    return create_impl(ctx, device_id, cuda_stream, true);
}

int wc_plan_destroy(wc_plan* plan);
int wc_minmax_batch(wc_ctx* ctx, const wc_box_desc* boxes, int n_units, int space, float* mins, float* maxs);
int wc_plan_compress_to_host(wc_plan* p, double keep, wc_packed* out);
int wc_plan_compress_to_host_chunked(wc_plan* p, double keep, wc_packed* out, wc_chunk_fn on_chunk, void* user);

int wc_destroy(wc_ctx* ctx) {
    if (!ctx) return WC_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->batch_plan) wc_plan_destroy(ctx->batch_plan);
    DevBuf* bufs[] = { &ctx->ws_pairs, &ctx->ws_coef, &ctx->ws_boxes, &ctx->ws_tbl0, &ctx->ws_tbl1,
                       &ctx->ws_tbl2, &ctx->ws_tiles0, &ctx->ws_tiles1, &ctx->ws_sum, &ctx->ws_misc,
                       &ctx->ws_a, &ctx->ws_b };
    for (DevBuf* b : bufs) b->release();
    ctx->ws_pin.release();
    ctx->ls.destroy();
    ctx->d_counter.release();
    if (ctx->s_aux) cudaStreamDestroy(ctx->s_aux);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    cudaGetLastError();
    delete ctx;
    return WC_OK;
}

int wc_sync(wc_ctx* ctx) {
    if (!ctx) return WC_ERR_INVALID_ARG;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    CTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return WC_OK;
}

const char* wc_last_error(const wc_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int wc_set_option(wc_ctx* ctx, int option, int64_t value) {
    if (!ctx) return WC_ERR_INVALID_ARG;
    switch (option) {
    case WC_OPT_PATH:
        if (value < 0 || value > 2) return WC_ERR_INVALID_ARG;
        ctx->opt_path = (int)value;
        return WC_OK;
    case WC_OPT_OVERLAP:
        ctx->opt_overlap = value != 0;
        return WC_OK;
    case WC_OPT_SEG_INDEX:
        if (value < 0 || value > 1) return WC_ERR_INVALID_ARG;
        ctx->opt_seg_index = (int)value;
        return WC_OK;
    case WC_OPT_COPY_ONLY:
        ctx->opt_copy_only = value != 0;
        return WC_OK;
    case WC_OPT_INGEST_STATS:
        ctx->opt_ingest_stats = value != 0;
        return WC_OK;
    case WC_OPT_DECODE_PIPE:
        if (value < 0 || value > 2) return WC_ERR_INVALID_ARG;
        ctx->opt_decode_pipe = (int)value;
        return WC_OK;
    case WC_OPT_PROFILE:
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        ctx->ls.collect();
        ctx->ls.profile = value != 0;
        return WC_OK;
    default: return WC_ERR_INVALID_ARG;
    }
}

int wc_get_counter(const wc_ctx* ctx, int counter, uint64_t* value) {
    if (!ctx || !value) return WC_ERR_INVALID_ARG;
    switch (counter) {
    case WC_CTR_KERNEL_LAUNCHES: *value = ctx->ls.launches; return WC_OK;
    case WC_CTR_H2D_BYTES: *value = ctx->h2d; return WC_OK;
    case WC_CTR_D2H_BYTES: *value = ctx->d2h; return WC_OK;
    default: return WC_ERR_INVALID_ARG;
    }
}

int wc_kernel_stats(wc_ctx* ctx, int index, const char** name, double* total_ms, uint64_t* launches) {
    if (!ctx || index < 0) return WC_ERR_INVALID_ARG;
    if (index >= KID_N) return WC_ERR_INVALID_ARG;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    CTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->ls.collect();
    if (name) *name = kernel_name(index);
    if (total_ms) *total_ms = ctx->ls.ms[index];
    if (launches) *launches = ctx->ls.count[index];
    return WC_OK;
}

#ifdef WC_PHASE_PROFILE
// Debug hook of PHASE_PROFILE builds only (not part of include/wcgpu.h, absent from the product build): per-phase SM cycles of the fused compress
// kernel summed over CTAs since the last reset: A, B, C1, scan, C2, units.
__attribute__((visibility("default"))) int wc_debug_phase_cycles(wc_ctx* ctx, unsigned long long* out6, int reset) {
    if (!ctx || !out6) return WC_ERR_INVALID_ARG;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    CTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CTX_CUDA(ctx, debug_phase_cycles(out6, reset != 0));
    return WC_OK;
}
#endif

int wc_reset_counters(wc_ctx* ctx) {
    if (!ctx) return WC_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->ls.collect();
    ctx->ls.reset();
    ctx->h2d = ctx->d2h = 0;
    return WC_OK;
}

int wc_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return WC_ERR_INVALID_ARG;
    *ptr = nullptr;
    cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? WC_ERR_OOM : WC_ERR_NO_DEVICE;
    }
    return WC_OK;
}

int wc_host_free(void* ptr) {
    if (ptr && cudaFreeHost(ptr) != cudaSuccess) {
        cudaGetLastError();
        return WC_ERR_CUDA;
    }
    return WC_OK;
}

int wc_device_alloc(wc_ctx* ctx, void** ptr, size_t bytes) {
    if (!ctx || !ptr) return WC_ERR_INVALID_ARG;
    *ptr = nullptr;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    CTX_CUDA(ctx, cudaMalloc(ptr, bytes ? bytes : 1));
    return WC_OK;
}

int wc_device_free(wc_ctx* ctx, void* ptr) {
    if (!ctx) return WC_ERR_INVALID_ARG;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ptr) CTX_CUDA(ctx, cudaFree(ptr));
    return WC_OK;
}

int wc_memcpy(wc_ctx* ctx, void* dst, const void* src, size_t bytes, int kind) {
    if (!ctx || (bytes && (!dst || !src)) || kind < 0 || kind > 2) return WC_ERR_INVALID_ARG;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice
                                 : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    CTX_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, k, ctx->stream));
    CTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (kind == 0) ctx->h2d += bytes;
    if (kind == 1) ctx->d2h += bytes;
    return WC_OK;
}

int wc_serialize_header(const wc_packed* unit, uint8_t header_out[20]) {
    if (!unit || !header_out) return WC_ERR_INVALID_ARG;
    int32_t h[5] = { unit->shape[0], unit->shape[1], unit->shape[2], unit->ncoef, unit->npairs };
    std::memcpy(header_out, h, 20); // native byte order, as serialize_int (src/compressor.cpp:47-51)
    return WC_OK;
}

// ---------------------------------------------------------------------------------------------
// plans
// ---------------------------------------------------------------------------------------------
static int plan_upload_units(wc_plan* p) {
    wc_ctx* ctx = p->ctx;
    if (p->n_units == 0) return WC_OK;
    CTX_CUDA(ctx, cudaMemcpyAsync(p->d_units.p, p->h_units.data(), sizeof(UnitDev) * p->n_units,
                                  cudaMemcpyHostToDevice, ctx->stream));
    // h_units is pageable: the copy above is staged synchronously by the runtime, safe to reuse
    return WC_OK;
}

int wc_plan_create(wc_ctx* ctx, const wc_box_desc* units, int n_units, int in_space,
                   wc_plan** out) {
    if (!ctx || !out || n_units < 0 || (n_units > 0 && !units) ||
        (in_space != WC_HOST && in_space != WC_DEVICE))
        return WC_ERR_INVALID_ARG;
    *out = nullptr;
    for (int i = 0; i < n_units; ++i) {
        int rc = check_dims(units[i].nx, units[i].ny, units[i].nz);
        if (rc != WC_OK) return rc;
        if (units[i].dtype != WC_F32 && units[i].dtype != WC_F64) return WC_ERR_INVALID_ARG;
        long long n = (long long)units[i].nx * units[i].ny * units[i].nz;
        if (n > 0 && !units[i].data) return WC_ERR_INVALID_ARG;
    }
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    wc_plan* p = new (std::nothrow) wc_plan();
    if (!p) return WC_ERR_OOM;
    p->ctx      = ctx;
    p->n_units  = n_units;
    p->in_space = in_space;
    p->units.assign(units, units + n_units);
    p->h_units.resize(n_units);
    p->in_dev_off.resize(n_units);
    p->has_segtab.assign(n_units, 0);
    std::vector<char> is_generic(n_units, 0);

    // classify + lay out
    size_t slot_pairs = 0, coef_floats = 0, in_off = 0;
    std::vector<size_t> slot_off(n_units), coef_off(n_units);
    std::vector<int2> xtiles, ctiles, pk_items;
    for (int i = 0; i < n_units; ++i) {
        const wc_box_desc& b = units[i];
        long long n = (long long)b.nx * b.ny * b.nz;
        UnitDev& u  = p->h_units[i];
        std::memset(&u, 0, sizeof(u));
        u.nx = b.nx; u.ny = b.ny; u.nz = b.nz; u.n = (int32_t)n; u.dtype = b.dtype;
        p->total_n += n;
        size_t bytes = (size_t)n * dtype_size(b.dtype);
        p->in_bytes += bytes;
        p->in_dev_off[i] = in_off;
        in_off += (bytes % 16 == 0) ? bytes : align_up(bytes, 256);
        slot_off[i] = slot_pairs;
        slot_pairs += align_up((size_t)n, 2); // keep every slot 16-byte aligned

        const void* cls_ptr = in_space == WC_HOST ? reinterpret_cast<const void*>(p->in_dev_off[i]) : b.data;
        int cls = fused_class(b.nx, b.ny, b.nz, b.dtype, cls_ptr);
        if (ctx->opt_path == 1) cls = 0;
        if (ctx->opt_path == 2 && cls == 0 && n > 0) {
            delete p;
            return WC_ERR_BAD_DIMS;
        }
        if (n == 0) cls = -1; // nothing to do: K = 0
        int fk = -1;
        for (int k = 0; k < FL_N; ++k)
            if (cls == FL_CLASS[k]) fk = k;
        if (fk >= 0) {
            p->fl[fk].push_back(i);
            // cluster classes: room for the decode-side segment table the compress kernel fills for free
            const size_t te = fused_decode_table_entries(cls, b.nx, b.ny, b.nz);
            if (te) {
                coef_off[i] = coef_floats;
                coef_floats += align_up(2 * te, 4);
                p->has_segtab[i] = 1;
            }
        } else if (cls == 0) {
            p->generic.push_back(i);
            is_generic[i] = 1;
            coef_off[i] = coef_floats;
            coef_floats += align_up((size_t)n, 4);
            if (ctx->opt_path != 1 && big_forward_slabs(b.nx, b.ny, b.nz, b.dtype, cls_ptr)) {
                p->bigfwd.push_back(i);
            } else {
                int nt = xtile_count(b.nx, b.ny, b.nz);
                for (int t = 0; t < nt; ++t) xtiles.push_back(make_int2(i, t));
            }
            u.ctile0  = (int32_t)ctiles.size();
            u.nctiles = ctile_count(n);
            for (int t = 0; t < u.nctiles; ++t) ctiles.push_back(make_int2(i, t));
            for (int c = 0, nc = big_pack_chunks(n); c < nc; ++c) pk_items.push_back(make_int2(i, c));
        }
    }
    p->n_pk_items = (int)pk_items.size();
    p->n_xtiles = (int)xtiles.size();
    p->n_ctiles = (int)ctiles.size();

    auto fail = [&](cudaError_t e, const char* what) {
        ctx->last_error = std::string(what) + ": " + cudaGetErrorString(e);
        cudaGetLastError();
        wc_plan_destroy(p);
        return e == cudaErrorMemoryAllocation ? WC_ERR_OOM : WC_ERR_CUDA;
    };
    cudaError_t e;
#define PLAN_RESERVE(buf, bytes)                                      \
    if ((e = (buf).reserve(bytes)) != cudaSuccess) return fail(e, "plan alloc " #buf)
    PLAN_RESERVE(p->d_units, sizeof(UnitDev) * std::max(n_units, 1));
    PLAN_RESERVE(p->d_states, sizeof(UnitState) * std::max(n_units, 1));
    PLAN_RESERVE(p->d_out, sizeof(wc_pair) * std::max<size_t>(slot_pairs, 2));
    PLAN_RESERVE(p->d_gkey, 64);
    if (coef_floats) PLAN_RESERVE(p->d_coef, sizeof(float) * coef_floats);
    if (p->n_xtiles) PLAN_RESERVE(p->d_xtiles, sizeof(int2) * p->n_xtiles);
    if (p->n_ctiles) {
        PLAN_RESERVE(p->d_ctiles, sizeof(int2) * p->n_ctiles);
        PLAN_RESERVE(p->d_tile_i, sizeof(int) * 4 * (size_t)p->n_ctiles);
    }
    if (in_space == WC_HOST && in_off) PLAN_RESERVE(p->d_in, in_off);
    if ((e = p->h_states.reserve(sizeof(UnitState) * std::max(n_units, 1) + 64)) != cudaSuccess)
        return fail(e, "plan pinned alloc");

    for (int i = 0; i < n_units; ++i) {
        UnitDev& u = p->h_units[i];
        u.in   = in_space == WC_HOST ? (const void*)(p->d_in.as<char>() + p->in_dev_off[i]) : units[i].data;
        u.out  = p->d_out.as<wc_pair>() + slot_off[i];
        // generic units: coefficient scratch; cluster-class fused units: segment table; other fused units: none
        const bool uses = p->has_segtab[i] || is_generic[i];
        u.coef = (uses && p->d_coef.p) ? p->d_coef.as<float>() + coef_off[i] : nullptr;
    }
    int rc = plan_upload_units(p);
    if (rc != WC_OK) { wc_plan_destroy(p); return rc; }
    if (p->n_xtiles)
        if ((e = cudaMemcpyAsync(p->d_xtiles.p, xtiles.data(), sizeof(int2) * p->n_xtiles,
                                 cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
            return fail(e, "xtiles upload");
    if (p->n_ctiles)
        if ((e = cudaMemcpyAsync(p->d_ctiles.p, ctiles.data(), sizeof(int2) * p->n_ctiles,
                                 cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
            return fail(e, "ctiles upload");
    // fused work lists
    for (int k = 0; k < FL_N; ++k) {
        if (p->fl[k].empty()) continue;
        PLAN_RESERVE(p->d_fl[k], sizeof(int) * p->fl[k].size());
        if ((e = cudaMemcpyAsync(p->d_fl[k].p, p->fl[k].data(), sizeof(int) * p->fl[k].size(),
                                 cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
            return fail(e, "fused list upload");
    }
    if (p->n_pk_items) {
        PLAN_RESERVE(p->d_pk_items, sizeof(int2) * (size_t)p->n_pk_items);
        PLAN_RESERVE(p->d_pk_status, sizeof(u64) * (size_t)p->n_pk_items);
        if ((e = cudaMemcpyAsync(p->d_pk_items.p, pk_items.data(), sizeof(int2) * (size_t)p->n_pk_items, cudaMemcpyHostToDevice,
                                 ctx->stream)) != cudaSuccess)
            return fail(e, "pack item upload");
    }
    if (!p->bigfwd.empty()) {
        sort_big_list(p->bigfwd, [&](int i, int& nx, int& ny, int& nz) { nx = p->h_units[i].nx; ny = p->h_units[i].ny; nz = p->h_units[i].nz; });
        PLAN_RESERVE(p->d_bigfwd, sizeof(int) * p->bigfwd.size());
        if ((e = cudaMemcpyAsync(p->d_bigfwd.p, p->bigfwd.data(), sizeof(int) * p->bigfwd.size(), cudaMemcpyHostToDevice,
                                 ctx->stream)) != cudaSuccess)
            return fail(e, "big-box list upload");
    }
#undef PLAN_RESERVE
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return fail(e, "plan sync");
    *out = p;
    return WC_OK;
}

int wc_plan_destroy(wc_plan* p) {
    if (!p) return WC_OK;
    wc_ctx* ctx = p->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = { &p->d_units, &p->d_states, &p->d_in, &p->d_out, &p->d_coef, &p->d_xtiles,
                       &p->d_ctiles, &p->d_tile_i, &p->d_gkey, &p->d_offsets, &p->d_dense, &p->d_dec_units, &p->d_inv_units, &p->d_inv_tiles, &p->d_ptiles,
                       &p->d_psum, &p->d_err, &p->d_rmse_units, &p->d_rmse_sum, &p->d_rmse,
                       &p->d_stage_out };
    for (DevBuf* b : bufs) b->release();
    for (int k = 0; k < FL_N; ++k) p->d_fl[k].release();
    p->d_bigfwd.release();
    p->d_pk_items.release();
    p->d_pk_status.release();
    p->d_q.release();
    p->d_qhist.release();
    p->d_qrank.release();
    p->d_running.release();
    p->d_counter.release();
    p->d_rmse_tiles.release();
    p->d_dec_list.release();
    if (p->s_inptr) { cudaStreamSynchronize(p->s_inptr); cudaStreamDestroy(p->s_inptr); }
    for (int i = 0; i < wc_plan::IN_RING; ++i) {
        p->h_inptr[i].release();
        p->d_inptr[i].release();
        if (p->ev_patch[i]) cudaEventDestroy(p->ev_patch[i]);
        if (p->ev_inptr[i]) cudaEventDestroy(p->ev_inptr[i]);
    }
    delete p->dec_cache;
    if (p->s_aux) cudaStreamDestroy(p->s_aux);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_join) cudaEventDestroy(p->ev_join);
    for (cudaEvent_t e : p->ev) cudaEventDestroy(e);
    for (cudaEvent_t e : p->ev_d2h) cudaEventDestroy(e);
    if (p->s_h2d) cudaStreamDestroy(p->s_h2d);
    if (p->s_d2h) cudaStreamDestroy(p->s_d2h);
    p->h_states.release();
    p->h_dense.release();
    p->h_misc.release();
    cudaGetLastError();
    if (ctx->batch_plan == p) ctx->batch_plan = nullptr;
    delete p;
    return WC_OK;
}

int wc_plan_set_inputs(wc_plan* p, const wc_box_desc* units) {
    if (!p || (p->n_units > 0 && !units)) return WC_ERR_INVALID_ARG;
    wc_ctx* ctx = p->ctx;
    for (int i = 0; i < p->n_units; ++i) {
        const wc_box_desc& a = p->units[i];
        if (units[i].nx != a.nx || units[i].ny != a.ny || units[i].nz != a.nz ||
            units[i].dtype != a.dtype)
            return WC_ERR_INVALID_ARG;
        // the fused kernels read their input with 16-byte TMA bulk copies: the kernel class chosen at
        // plan creation must still fit the new address
        if (p->in_space == WC_DEVICE && ((reinterpret_cast<uintptr_t>(a.data) & 15u) == 0) &&
            (reinterpret_cast<uintptr_t>(units[i].data) & 15u) != 0)
            return WC_ERR_INVALID_ARG;
    }
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < p->n_units; ++i) {
        p->units[i].data = units[i].data;
        if (p->in_space == WC_DEVICE) p->h_units[i].in = units[i].data;
    }
    if (p->in_space == WC_DEVICE && p->n_units > 0) {
        // Stream-ordered and without a host synchronisation: the new addresses go through a ring of pinned
        // tables and device tables (one H2D of 8 bytes per unit, on a side stream so that it overlaps whatever the
        // ctx stream is still doing — in-stream it cost ~30 us per step, ~80 us with eight ranks sharing the host) and
        // a kernel on the ctx stream patches UnitDev::in, so a timestep series can call this between two
        // wc_plan_compress without draining the GPU.  A pinned slot is reused only after the copy that read it has
        // completed, a device slot only after the patch kernel that read it (events).
        const int slot = (int)(p->inptr_next++ % wc_plan::IN_RING);
        const size_t bytes = sizeof(void*) * (size_t)p->n_units;
        CTX_CUDA(ctx, p->h_inptr[slot].reserve(bytes));
        CTX_CUDA(ctx, p->d_inptr[slot].reserve(bytes));
        if (!p->s_inptr) CTX_CUDA(ctx, cudaStreamCreateWithFlags(&p->s_inptr, cudaStreamNonBlocking));
        if (!p->ev_inptr[slot]) {
            CTX_CUDA(ctx, cudaEventCreateWithFlags(&p->ev_inptr[slot], cudaEventDisableTiming));
            CTX_CUDA(ctx, cudaEventCreateWithFlags(&p->ev_patch[slot], cudaEventDisableTiming));
        } else {
            CTX_CUDA(ctx, cudaEventSynchronize(p->ev_inptr[slot]));
            CTX_CUDA(ctx, cudaStreamWaitEvent(p->s_inptr, p->ev_patch[slot], 0));
        }
        const void** hp = p->h_inptr[slot].as<const void*>();
        for (int i = 0; i < p->n_units; ++i) hp[i] = units[i].data;
        CTX_CUDA(ctx, cudaMemcpyAsync(p->d_inptr[slot].p, hp, bytes, cudaMemcpyHostToDevice, p->s_inptr));
        CTX_CUDA(ctx, cudaEventRecord(p->ev_inptr[slot], p->s_inptr));
        ctx->h2d += bytes;
        CTX_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, p->ev_inptr[slot], 0));
        CTX_CUDA(ctx, launch_patch_inputs(p->d_units.as<UnitDev>(), p->d_inptr[slot].as<const void*>(), p->n_units,
                                          ctx->stream, &ctx->ls));
        CTX_CUDA(ctx, cudaEventRecord(p->ev_patch[slot], ctx->stream));
    }
    return WC_OK;
}

static int plan_stage_inputs(wc_plan* p) {
    wc_ctx* ctx = p->ctx;
    if (p->in_space != WC_HOST) return WC_OK;
    CopyList cl;
    for (int i = 0; i < p->n_units; ++i) {
        size_t bytes = (size_t)p->h_units[i].n * dtype_size(p->units[i].dtype);
        cl.add(p->d_in.as<char>() + p->in_dev_off[i], p->units[i].data, bytes);
    }
    for (const CopyRange& r : cl.r) {
        CTX_CUDA(ctx, cudaMemcpyAsync(r.dst, r.src, r.bytes, cudaMemcpyHostToDevice, ctx->stream));
        ctx->h2d += r.bytes;
    }
    return WC_OK;
}

// A zeroed work counter for one single-CTA-class launch on `st` (dynamic unit hand-out: evens out the
// ~10 % speed spread between SMs and lets late CTAs take less).  Ring of 64: launches that could still be
// running 64 launches later are ordered before the reuse by the streams' own ordering / the fork-join events.
static int plan_counter(wc_plan* p, cudaStream_t st, int** out) {
    wc_ctx* ctx = p->ctx;
    if (!p->d_counter.p) CTX_CUDA(ctx, p->d_counter.reserve(64 * sizeof(int)));
    int* c = p->d_counter.as<int>() + (p->counter_next++ & 63);
    CTX_CUDA(ctx, cudaMemsetAsync(c, 0, sizeof(int), st));
    *out = c;
    return WC_OK;
}

// forward transform (+ arg-max keys) of every unit; the fused classes only run here when the
// threshold is batch-wide (they otherwise do everything in one kernel in plan_pack)
static int plan_forward(wc_plan* p, bool global_mode) {
    wc_ctx* ctx = p->ctx;
    CTX_CUDA(ctx, cudaMemsetAsync(p->d_states.p, 0, sizeof(UnitState) * std::max(p->n_units, 1),
                                  ctx->stream));
    CTX_CUDA(ctx, launch_forward_generic(p->d_units.as<UnitDev>(), p->d_states.as<UnitState>(),
                                         p->d_xtiles.as<int2>(), p->n_xtiles, ctx->stream,
                                         &ctx->ls));
    if (!p->bigfwd.empty()) {
        auto dims = [&](int i, int& nx, int& ny, int& nz) { nx = p->h_units[i].nx; ny = p->h_units[i].ny; nz = p->h_units[i].nz; };
        int rc = for_each_run(FUSED_CLS_RBIG, p->bigfwd, dims, [&](size_t a, size_t cnt, int s_rt) -> int {
            int* counter = nullptr;
            int rc2 = plan_counter(p, ctx->stream, &counter);
            if (rc2 != WC_OK) return rc2;
            CTX_CUDA(ctx, launch_big_forward(p->d_units.as<UnitDev>(), p->d_states.as<UnitState>(), p->d_bigfwd.as<int>() + a,
                                             (int)cnt, s_rt, counter, ctx->sm_count, ctx->stream, &ctx->ls));
            return WC_OK;
        });
        if (rc != WC_OK) return rc;
    }
    if (global_mode) {
        for (int k = 0; k < FL_N; ++k)
            if (!p->fl[k].empty()) {
                int* counter = nullptr;
                if (!fl_is_cluster(k)) {
                    int rc = plan_counter(p, ctx->stream, &counter);
                    if (rc != WC_OK) return rc;
                }
                CTX_CUDA(ctx, launch_fused_compress(FL_CLASS[k], FUSED_KEYS_ONLY | (ctx->opt_ingest_stats ? FUSED_MINMAX : 0),
                                                    p->d_units.as<UnitDev>(),
                                                    p->d_states.as<UnitState>(), p->d_fl[k].as<int>(),
                                                    (int)p->fl[k].size(), 0.0, nullptr, ctx->sm_count,
                                                    ctx->stream, &ctx->ls, counter));
            }
    }
    return WC_OK;
}

static int plan_pack(wc_plan* p, double keep, const u64* global_key_dev) {
    wc_ctx* ctx = p->ctx;
    // 1 - keep evaluated in double exactly as `(1 - keep)` at src/compressor.cpp:216
    volatile double one = 1.0;
    double omk = one - keep;
    if (!p->generic.empty() || global_key_dev) {
        CTX_CUDA(ctx, launch_finalize_thresh(p->d_states.as<UnitState>(), p->n_units, omk,
                                             global_key_dev, ctx->stream, &ctx->ls));
    }
    if (!p->generic.empty()) {
        if (ctx->opt_path == 1 && ctx->opt_seg_index == 1) {
            // the three-kernel packing (count, scan, emit), kept selectable for comparison
            int* ti = p->d_tile_i.as<int>();
            size_t nt = (size_t)p->n_ctiles;
            CTX_CUDA(ctx, launch_pack_generic(p->d_units.as<UnitDev>(), p->d_states.as<UnitState>(),
                                              p->n_units, p->d_ctiles.as<int2>(), p->n_ctiles, ti,
                                              ti + nt, ti + 2 * nt, ti + 3 * nt, ctx->stream,
                                              &ctx->ls));
        } else {
            int* counter = nullptr;
            int rc = plan_counter(p, ctx->stream, &counter);
            if (rc != WC_OK) return rc;
            CTX_CUDA(ctx, launch_big_pack(p->d_units.as<UnitDev>(), p->d_states.as<UnitState>(), p->d_pk_items.as<int2>(),
                                          p->n_pk_items, p->d_pk_status.as<u64>(), counter, ctx->sm_count, ctx->stream, &ctx->ls));
        }
    }
    int mode = global_key_dev ? FUSED_GIVEN_THRESH : (FUSED_FULL | (ctx->opt_ingest_stats ? FUSED_MINMAX : 0));
    // The cluster kernel cannot use every SM (clusters of 8 must fit inside a GPC: 15 clusters = 120 of 148
    // SMs on B200).  When a step has both classes, the single-CTA kernel runs concurrently on a second
    // stream with dynamic unit hand-out and back-fills the idle SMs.  (Serialised while per-kernel event
    // profiling is on, so each kernel is timed alone.)
    bool any_cluster = false, any_single = false;
    for (int k = 0; k < FL_N; ++k)
        if (!p->fl[k].empty()) (fl_is_cluster(k) ? any_cluster : any_single) = true;
    const bool overlap = any_cluster && any_single && !ctx->ls.profile && ctx->opt_overlap;
    if (overlap) {
        if (!p->s_aux) {
            CTX_CUDA(ctx, cudaStreamCreateWithFlags(&p->s_aux, cudaStreamNonBlocking));
            CTX_CUDA(ctx, cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
            CTX_CUDA(ctx, cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
        }
        CTX_CUDA(ctx, cudaEventRecord(p->ev_fork, ctx->stream));
        CTX_CUDA(ctx, cudaStreamWaitEvent(p->s_aux, p->ev_fork, 0));
    }
    for (int k = 0; k < FL_N; ++k) {
        if (p->fl[k].empty()) continue;
        cudaStream_t st = (overlap && !fl_is_cluster(k)) ? p->s_aux : ctx->stream;
        int* counter = nullptr;
        if (!fl_is_cluster(k)) {
            int rc = plan_counter(p, st, &counter);
            if (rc != WC_OK) return rc;
        }
        CTX_CUDA(ctx, launch_fused_compress(FL_CLASS[k], mode, p->d_units.as<UnitDev>(),
                                            p->d_states.as<UnitState>(), p->d_fl[k].as<int>(),
                                            (int)p->fl[k].size(), omk, global_key_dev, ctx->sm_count,
                                            st, &ctx->ls, counter));
    }
    if (overlap) {
        CTX_CUDA(ctx, cudaEventRecord(p->ev_join, p->s_aux));
        CTX_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, p->ev_join, 0));
    }
    p->compressed = true;
    return WC_OK;
}

// ---- EXTENSION: quantile thresholds (radix select over the coefficient scratch, see k_q_hist) -----------------
int wc_plan_quantile_begin(wc_plan* p, double keep, int global, uint64_t n_total) {
    if (!p || !(keep >= 0.0 && keep <= 1.0)) return WC_ERR_INVALID_ARG;
    wc_ctx* ctx = p->ctx;
    // every unit needs its coefficients in HBM: plans whose units all take the scratch path (WC_OPT_PATH = 1 at creation)
    for (int k = 0; k < FL_N; ++k)
        if (!p->fl[k].empty()) {
            ctx->last_error = "quantile thresholds need a plan created under WC_OPT_PATH = 1 (a coefficient scratch for every unit)";
            return WC_ERR_STATE;
        }
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = plan_stage_inputs(p);
    if (rc != WC_OK) return rc;
    p->has_stats = false;
    rc = plan_forward(p, false);
    if (rc != WC_OK) return rc;
    const int nq = global ? 1 : p->n_units;
    std::vector<unsigned long long> rank((size_t)std::max(nq, 1));
    auto kept_target = [&](unsigned long long n) {            // Kt = n - floor(keep * n): rank of the threshold element
        const unsigned long long drop = (unsigned long long)std::floor(keep * (double)n);
        return n - std::min(drop, n);
    };
    if (global) {
        unsigned long long n = n_total;
        if (n == 0) n = (unsigned long long)p->total_n;
        rank[0] = kept_target(n);
    } else {
        for (int i = 0; i < p->n_units; ++i) rank[i] = kept_target((unsigned long long)p->h_units[i].n);
    }
    CTX_CUDA(ctx, p->d_q.reserve(sizeof(QState) * (size_t)std::max(nq, 1)));
    CTX_CUDA(ctx, p->d_qhist.reserve(sizeof(unsigned long long) * Q_BINS * (size_t)std::max(nq, 1)));
    CTX_CUDA(ctx, p->d_qrank.reserve(sizeof(unsigned long long) * (size_t)std::max(nq, 1)));
    CTX_CUDA(ctx, cudaMemcpyAsync(p->d_qrank.p, rank.data(), sizeof(unsigned long long) * nq, cudaMemcpyHostToDevice, ctx->stream));
    CTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // `rank` is a local
    CTX_CUDA(ctx, cudaMemsetAsync(p->d_qhist.p, 0, sizeof(unsigned long long) * Q_BINS * (size_t)nq, ctx->stream));
    CTX_CUDA(ctx, launch_q_init(p->d_q.as<QState>(), p->d_qrank.as<unsigned long long>(), nq, ctx->stream));
    p->q_global = global ? 1 : 0;
    p->q_keep = keep;
    return WC_OK;
}

int wc_plan_quantile_hist(wc_plan* p, int pass, uint64_t** hist_dev) {
    if (!p || pass < 0 || pass > 2) return WC_ERR_INVALID_ARG;
    if (p->q_global < 0) return WC_ERR_STATE;
    wc_ctx* ctx = p->ctx;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    CTX_CUDA(ctx, launch_q_hist(p->d_units.as<UnitDev>(), p->d_ctiles.as<int2>(), p->n_ctiles, p->d_q.as<QState>(),
                                p->d_qhist.as<unsigned long long>(), pass, p->q_global == 1, ctx->stream, &ctx->ls));
    if (hist_dev) *hist_dev = reinterpret_cast<uint64_t*>(p->d_qhist.p);
    return WC_OK;
}

int wc_plan_quantile_pick(wc_plan* p, int pass) {
    if (!p || pass < 0 || pass > 2) return WC_ERR_INVALID_ARG;
    if (p->q_global < 0) return WC_ERR_STATE;
    wc_ctx* ctx = p->ctx;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    CTX_CUDA(ctx, launch_q_pick(p->d_q.as<QState>(), p->d_qhist.as<unsigned long long>(), p->q_global ? 1 : p->n_units, pass,
                                ctx->stream, &ctx->ls));
    return WC_OK;
}

int wc_plan_quantile_pack(wc_plan* p) {
    if (!p) return WC_ERR_INVALID_ARG;
    if (p->q_global < 0) return WC_ERR_STATE;
    wc_ctx* ctx = p->ctx;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    CTX_CUDA(ctx, launch_q_apply(p->d_states.as<UnitState>(), p->d_q.as<QState>(), p->n_units, p->q_global == 1, ctx->stream));
    if (!p->generic.empty()) {
        int* counter = nullptr;
        int rc = plan_counter(p, ctx->stream, &counter);
        if (rc != WC_OK) return rc;
        CTX_CUDA(ctx, launch_big_pack(p->d_units.as<UnitDev>(), p->d_states.as<UnitState>(), p->d_pk_items.as<int2>(),
                                      p->n_pk_items, p->d_pk_status.as<u64>(), counter, ctx->sm_count, ctx->stream, &ctx->ls));
    }
    p->q_global = -1;
    p->compressed = true;
    return WC_OK;
}

int wc_plan_compress(wc_plan* p, double keep, int thresh_mode) {
    if (p && (thresh_mode == WC_THRESH_QUANTILE || thresh_mode == WC_THRESH_QUANTILE_GLOBAL)) {
        int rc = wc_plan_quantile_begin(p, keep, thresh_mode == WC_THRESH_QUANTILE_GLOBAL, 0);
        for (int pass = 0; pass < 3 && rc == WC_OK; ++pass) {
            rc = wc_plan_quantile_hist(p, pass, nullptr);
            if (rc == WC_OK) rc = wc_plan_quantile_pick(p, pass);
        }
        return rc == WC_OK ? wc_plan_quantile_pack(p) : rc;
    }
    if (!p || (thresh_mode != WC_THRESH_PER_UNIT && thresh_mode != WC_THRESH_GLOBAL))
        return WC_ERR_INVALID_ARG;
    wc_ctx* ctx = p->ctx;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = plan_stage_inputs(p);
    if (rc != WC_OK) return rc;
    p->has_stats = ctx->opt_ingest_stats != 0;
    bool global_mode = thresh_mode == WC_THRESH_GLOBAL;
    rc = plan_forward(p, global_mode);
    if (rc != WC_OK) return rc;
    const u64* gk = nullptr;
    if (global_mode) {
        CTX_CUDA(ctx, launch_global_key(p->d_units.as<UnitDev>(), p->d_states.as<UnitState>(), p->n_units, p->d_gkey.as<u64>(),
                                        ctx->stream, &ctx->ls));
        gk = p->d_gkey.as<u64>();
    }
    return plan_pack(p, keep, gk);
}

int wc_plan_transform(wc_plan* p, uint64_t** key_dev) {
    if (!p || !key_dev) return WC_ERR_INVALID_ARG;
    wc_ctx* ctx = p->ctx;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = plan_stage_inputs(p);
    if (rc != WC_OK) return rc;
    p->has_stats = ctx->opt_ingest_stats != 0;
    rc = plan_forward(p, true);
    if (rc != WC_OK) return rc;
    CTX_CUDA(ctx, launch_global_key(p->d_units.as<UnitDev>(), p->d_states.as<UnitState>(), p->n_units, p->d_gkey.as<u64>(),
                                    ctx->stream, &ctx->ls));
    *key_dev       = reinterpret_cast<uint64_t*>(p->d_gkey.p);
    p->transformed = true;
    return WC_OK;
}

int wc_plan_pack_with_key(wc_plan* p, double keep, const uint64_t* key_dev) {
    if (!p || !key_dev) return WC_ERR_INVALID_ARG;
    if (!p->transformed) return WC_ERR_STATE;
    wc_ctx* ctx = p->ctx;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    return plan_pack(p, keep, reinterpret_cast<const u64*>(key_dev));
}

static int plan_read_states(wc_plan* p) {
    wc_ctx* ctx = p->ctx;
    if (p->n_units == 0) return WC_OK;
    CTX_CUDA(ctx, cudaMemcpyAsync(p->h_states.p, p->d_states.p, sizeof(UnitState) * p->n_units,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    ctx->d2h += sizeof(UnitState) * p->n_units;
    CTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return WC_OK;
}

int wc_plan_total_pairs(wc_plan* p, int64_t* total) {
    if (!p || !total) return WC_ERR_INVALID_ARG;
    if (!p->compressed) return WC_ERR_STATE;
    CTX_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
    int rc = plan_read_states(p);
    if (rc != WC_OK) return rc;
    int64_t t = 0;
    const UnitState* hs = p->h_states.as<UnitState>();
    for (int i = 0; i < p->n_units; ++i) t += hs[i].npairs;
    *total = t;
    return WC_OK;
}

int wc_plan_fetch(wc_plan* p, wc_packed* out, int out_space) {
    if (!p || (p->n_units > 0 && !out) || (out_space != WC_HOST && out_space != WC_DEVICE))
        return WC_ERR_INVALID_ARG;
    if (!p->compressed) return WC_ERR_STATE;
    wc_ctx* ctx = p->ctx;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    if (out_space == WC_HOST && p->n_units > 0) {
        cudaError_t e = p->d_offsets.reserve(sizeof(long long) * (p->n_units + 1));
        CTX_CUDA(ctx, e);
        CTX_CUDA(ctx, launch_gather_dense(p->d_units.as<UnitDev>(), p->d_states.as<UnitState>(),
                                          p->n_units, p->d_offsets.as<long long>(), nullptr, true,
                                          ctx->stream, &ctx->ls));
    }
    int rc = plan_read_states(p);
    if (rc != WC_OK) return rc;
    const UnitState* hs = p->h_states.as<UnitState>();
    size_t total = 0;
    for (int i = 0; i < p->n_units; ++i) {
        const UnitDev& u = p->h_units[i];
        out[i].shape[0] = u.nx; out[i].shape[1] = u.ny; out[i].shape[2] = u.nz;
        out[i].ncoef    = u.n;
        out[i].npairs   = hs[i].npairs;
        out[i].flags    = (hs[i].flags & UNIT_FLAG_NEED32) ? WC_PACKED_NEED32 : 0;
        out[i].pairs    = u.out;
        total += (size_t)hs[i].npairs;
    }
    if (out_space == WC_DEVICE || p->n_units == 0) return WC_OK;
    CTX_CUDA(ctx, p->d_dense.reserve(sizeof(wc_pair) * std::max<size_t>(total, 1)));
    CTX_CUDA(ctx, p->h_dense.reserve(sizeof(wc_pair) * std::max<size_t>(total, 1)));
    CTX_CUDA(ctx, launch_gather_dense(p->d_units.as<UnitDev>(), p->d_states.as<UnitState>(),
                                      p->n_units, p->d_offsets.as<long long>(),
                                      p->d_dense.as<wc_pair>(), false, ctx->stream, &ctx->ls));
    if (total) {
        CTX_CUDA(ctx, cudaMemcpyAsync(p->h_dense.p, p->d_dense.p, sizeof(wc_pair) * total,
                                      cudaMemcpyDeviceToHost, ctx->stream));
        ctx->d2h += sizeof(wc_pair) * total;
    }
    CTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    size_t off = 0;
    for (int i = 0; i < p->n_units; ++i) {
        out[i].pairs = p->h_dense.as<wc_pair>() + off;
        off += (size_t)hs[i].npairs;
    }
    return WC_OK;
}


// Pipelined host path: the unit list is cut into chunks; chunk c's boxes are copied H2D on one stream
// while chunk c-1 is compressed on the ctx stream and chunk c-2's pairs are gathered densely and copied
// D2H on a third stream (PCIe is full duplex), so the call costs about max(H2D, D2H) instead of their
// sum.  Same result as wc_plan_compress + wc_plan_fetch(WC_HOST).
int wc_plan_compress_to_host(wc_plan* p, double keep, wc_packed* out) {
    return wc_plan_compress_to_host_chunked(p, keep, out, nullptr, nullptr);
}

int wc_plan_compress_to_host_chunked(wc_plan* p, double keep, wc_packed* out, wc_chunk_fn on_chunk, void* user) {
    if (!p || (p->n_units > 0 && !out)) return WC_ERR_INVALID_ARG;
    wc_ctx* ctx = p->ctx;
    if (p->in_space != WC_HOST || !p->generic.empty() || p->n_units == 0) {
        int rc = wc_plan_compress(p, keep, WC_THRESH_PER_UNIT);
        if (rc != WC_OK) return rc;
        rc = wc_plan_fetch(p, out, WC_HOST);
        if (rc == WC_OK && on_chunk && p->n_units > 0) on_chunk(user, 0, p->n_units, out);
        return rc;
    }
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    p->has_stats = ctx->opt_ingest_stats != 0;
    const int NCH = 16;
    if (!p->s_h2d) {
        CTX_CUDA(ctx, cudaStreamCreateWithFlags(&p->s_h2d, cudaStreamNonBlocking));
        CTX_CUDA(ctx, cudaStreamCreateWithFlags(&p->s_d2h, cudaStreamNonBlocking));
        p->ev.resize(3 * NCH + 1);
        for (cudaEvent_t& e : p->ev) CTX_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CTX_CUDA(ctx, p->d_running.reserve(64));
        CTX_CUDA(ctx, p->h_misc.reserve(sizeof(long long) * (NCH + 1)));
        CTX_CUDA(ctx, p->d_offsets.reserve(sizeof(long long) * (p->n_units + NCH + 1)));
    }
    // worst case: every coefficient kept.  The pinned destination is sized for the worst case up front
    // as well: the chunk callback hands out pointers into it while later chunks are still in flight, so it
    // must never move.
    CTX_CUDA(ctx, p->d_dense.reserve(sizeof(wc_pair) * std::max<size_t>((size_t)p->total_n, 1)));
    if (on_chunk) CTX_CUDA(ctx, p->h_dense.reserve(sizeof(wc_pair) * std::max<size_t>((size_t)p->total_n, 1)));
    const bool copy_only = ctx->opt_copy_only != 0;      // probe: same copies, no kernels (previous counts stand)
    volatile double one = 1.0;
    const double omk = one - keep;
    // chunk boundaries by input bytes
    std::vector<int> cut(NCH + 1, p->n_units);
    cut[0] = 0;
    {
        size_t acc = 0, per = p->in_bytes / NCH + 1;
        int c = 1;
        for (int i = 0; i < p->n_units && c < NCH; ++i) {
            acc += (size_t)p->h_units[i].n * dtype_size(p->units[i].dtype);
            if (acc >= per * c) cut[c++] = i + 1;
        }
    }
    long long* h_run = p->h_misc.as<long long>();
    CTX_CUDA(ctx, cudaMemsetAsync(p->d_running.p, 0, 8, ctx->stream));
    if (!copy_only) CTX_CUDA(ctx, cudaMemsetAsync(p->d_states.p, 0, sizeof(UnitState) * p->n_units, ctx->stream));
    CTX_CUDA(ctx, cudaEventRecord(p->ev[3 * NCH], ctx->stream));
    CTX_CUDA(ctx, cudaStreamWaitEvent(p->s_h2d, p->ev[3 * NCH], 0));
    size_t fi[FL_N] = {};
    long long done_total = 0;
    UnitState* const hs_all = p->h_states.as<UnitState>();
    std::vector<long long> chunk_off(NCH + 1, 0);     // running pair total in front of each chunk
    std::vector<cudaEvent_t>& ev_done = p->ev_d2h;    // D2H of a chunk's pairs has landed (callback mode)
    if (on_chunk && ev_done.empty()) {
        ev_done.resize(NCH);
        for (cudaEvent_t& e : ev_done) CTX_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    int delivered = 0;                                // chunks already handed to the callback
    auto fill_out = [&](int c) {
        size_t off = (size_t)chunk_off[c];
        for (int i = cut[c]; i < cut[c + 1]; ++i) {
            const UnitDev& u = p->h_units[i];
            out[i].shape[0] = u.nx; out[i].shape[1] = u.ny; out[i].shape[2] = u.nz;
            out[i].ncoef    = u.n;
            out[i].npairs   = hs_all[i].npairs;
            out[i].flags    = (hs_all[i].flags & UNIT_FLAG_NEED32) ? WC_PACKED_NEED32 : 0;
            out[i].pairs    = p->h_dense.as<wc_pair>() + off;
            off += (size_t)hs_all[i].npairs;
        }
    };
    // hands every chunk whose D2H has completed to the callback, in order; `block`: wait for all up to `upto`
    auto deliver = [&](int upto, bool block) -> int {
        while (on_chunk && delivered < upto) {
            if (block) CTX_CUDA(ctx, cudaEventSynchronize(ev_done[delivered]));
            else if (cudaEventQuery(ev_done[delivered]) != cudaSuccess) { cudaGetLastError(); break; }
            fill_out(delivered);
            if (cut[delivered + 1] > cut[delivered])
                on_chunk(user, cut[delivered], cut[delivered + 1] - cut[delivered], out + cut[delivered]);
            ++delivered;
        }
        return WC_OK;
    };
    auto finish_chunk = [&](int c, long long& prev_total) -> int {
        // host learns the running total of chunk c, then enqueues its D2H
        CTX_CUDA(ctx, cudaEventSynchronize(p->ev[NCH + c]));
        long long tot = h_run[c];
        size_t need = sizeof(wc_pair) * (size_t)std::max<long long>(tot, 1);
        if (need > p->h_dense.cap) {
            // grow the pinned buffer; earlier chunks' copies must land first, then move them over
            CTX_CUDA(ctx, cudaStreamSynchronize(p->s_d2h));
            PinBuf bigger;
            CTX_CUDA(ctx, bigger.reserve(std::max(need, sizeof(wc_pair) * (size_t)p->total_n / 2 + 4096)));
            if (p->h_dense.p && prev_total > 0) std::memcpy(bigger.p, p->h_dense.p, sizeof(wc_pair) * (size_t)prev_total);
            p->h_dense.release();
            p->h_dense = bigger;
        }
        if (tot > prev_total) {
            CTX_CUDA(ctx, cudaStreamWaitEvent(p->s_d2h, p->ev[2 * NCH + c], 0));
            CTX_CUDA(ctx, cudaMemcpyAsync(p->h_dense.as<wc_pair>() + prev_total, p->d_dense.as<wc_pair>() + prev_total,
                                          sizeof(wc_pair) * (size_t)(tot - prev_total), cudaMemcpyDeviceToHost, p->s_d2h));
            ctx->d2h += sizeof(wc_pair) * (size_t)(tot - prev_total);
        }
        chunk_off[c]     = prev_total;
        chunk_off[c + 1] = tot;
        if (on_chunk) CTX_CUDA(ctx, cudaEventRecord(ev_done[c], p->s_d2h));
        prev_total = tot;
        return deliver(c, false);
    };
    for (int c = 0; c < NCH; ++c) {
        const int u0 = cut[c], u1 = cut[c + 1];
        // H2D of the chunk
        CopyList cl;
        for (int i = u0; i < u1; ++i)
            cl.add(p->d_in.as<char>() + p->in_dev_off[i], p->units[i].data, (size_t)p->h_units[i].n * dtype_size(p->units[i].dtype));
        for (const CopyRange& r : cl.r) {
            CTX_CUDA(ctx, cudaMemcpyAsync(r.dst, r.src, r.bytes, cudaMemcpyHostToDevice, p->s_h2d));
            ctx->h2d += r.bytes;
        }
        CTX_CUDA(ctx, cudaEventRecord(p->ev[c], p->s_h2d));
        // compute on the ctx stream
        CTX_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, p->ev[c], 0));
        for (int k = 0; k < FL_N; ++k) {
            size_t j = fi[k];
            while (j < p->fl[k].size() && p->fl[k][j] < u1) ++j;
            if (j > fi[k] && !copy_only) {
                int* counter = nullptr;
                if (!fl_is_cluster(k)) {
                    int rc = plan_counter(p, ctx->stream, &counter);
                    if (rc != WC_OK) return rc;
                }
                CTX_CUDA(ctx, launch_fused_compress(FL_CLASS[k], FUSED_FULL | (ctx->opt_ingest_stats ? FUSED_MINMAX : 0),
                                                    p->d_units.as<UnitDev>(),
                                                    p->d_states.as<UnitState>(), p->d_fl[k].as<int>() + fi[k],
                                                    (int)(j - fi[k]), omk, nullptr, ctx->sm_count, ctx->stream,
                                                    &ctx->ls, counter));
            }
            fi[k] = j;
        }
        if (u1 > u0) {
            CTX_CUDA(ctx, launch_gather_dense(p->d_units.as<UnitDev>() + u0, p->d_states.as<UnitState>() + u0, u1 - u0,
                                              p->d_offsets.as<long long>() + u0 + c, nullptr, true, ctx->stream, &ctx->ls,
                                              p->d_running.as<long long>()));
        }
        CTX_CUDA(ctx, cudaMemcpyAsync(&h_run[c], p->d_running.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (u1 > u0) {   // the chunk's unit records travel with its running total (callback mode needs them early)
            CTX_CUDA(ctx, cudaMemcpyAsync(hs_all + u0, p->d_states.as<UnitState>() + u0, sizeof(UnitState) * (size_t)(u1 - u0),
                                          cudaMemcpyDeviceToHost, ctx->stream));
            ctx->d2h += sizeof(UnitState) * (size_t)(u1 - u0);
        }
        CTX_CUDA(ctx, cudaEventRecord(p->ev[NCH + c], ctx->stream));
        if (u1 > u0 && !copy_only)
            CTX_CUDA(ctx, launch_gather_dense(p->d_units.as<UnitDev>() + u0, p->d_states.as<UnitState>() + u0, u1 - u0,
                                              p->d_offsets.as<long long>() + u0 + c, p->d_dense.as<wc_pair>(), false,
                                              ctx->stream, &ctx->ls));
        CTX_CUDA(ctx, cudaEventRecord(p->ev[2 * NCH + c], ctx->stream));
        // while chunk c computes, hand chunk c-1's pairs to the D2H stream
        if (c > 0) {
            int rc = finish_chunk(c - 1, done_total);
            if (rc != WC_OK) return rc;
        }
    }
    {
        int rc = finish_chunk(NCH - 1, done_total);
        if (rc != WC_OK) return rc;
    }
    p->compressed = true;
    if (on_chunk) {
        int rc = deliver(NCH, true);
        if (rc != WC_OK) return rc;
    }
    CTX_CUDA(ctx, cudaStreamSynchronize(p->s_d2h));
    CTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int c = delivered; c < NCH; ++c) fill_out(c);
    return WC_OK;
}

// ---------------------------------------------------------------------------------------------
// decompression
// ---------------------------------------------------------------------------------------------
// Common engine: units described by (pairs on device, K or a device pointer to K, dims, output).
struct DecJob {
    const wc_pair* pairs_dev;
    const int32_t* npairs_dev; // optional: K read on the device (plan round trip)
    int32_t        npairs;     // K (host known) or capacity bound when npairs_dev is set
    int32_t        nx, ny, nz;
    void*          out_dev;
    int32_t        out_dtype;
    const int2*    segtab = nullptr; // optional: segment table already on the device (plan round trip)
};

static inline void dec_hash(uint64_t& h1, uint64_t& h2, uint64_t v) {
    h1 = (h1 ^ v) * 0x9E3779B97F4A7C15ull; h1 ^= h1 >> 29;
    h2 = (h2 + v) * 0xC2B2AE3D27D4EB4Full; h2 ^= h2 >> 31;
}

// the decompress kernel of one fused class list (WC_OPT_DECODE_PIPE selects the staged variant of the 32^3 kernel)
static cudaError_t launch_decode(wc_ctx* ctx, int fused_cls, const DecUnitDev* dec, const InvUnitDev* inv, const int* list,
                                 int n, int* err, int* counter, bool v1_tables, int s_rt = 0) {
    return launch_fused_decompress(fused_cls, dec, inv, list, n, err, ctx->sm_count, ctx->stream, &ctx->ls, counter, v1_tables,
                                   ctx->opt_decode_pipe, s_rt);
}

// A plan that decodes into the same boxes again (keep sweeps of the estimate mode) re-launches from the
// device tables of the previous call; only all-fused batches without scratch are cached.
static int relaunch_decompress(wc_ctx* ctx, const DecCache* cache, DevBuf& d_dec_units, DevBuf& d_inv_units,
                               DevBuf& d_err, DevBuf& d_fused_list) {
    CTX_CUDA(ctx, cudaMemsetAsync(d_err.p, 0, 64, ctx->stream));
    CTX_CUDA(ctx, ctx->d_counter.reserve(64 * sizeof(int)));
    int* dl = d_fused_list.as<int>();
    size_t o = 0;
    for (int k = 0; k < FL_N; ++k) {
        if (!cache->fl_n[k]) continue;
        int* counter = ctx->d_counter.as<int>() + (ctx->counter_next++ & 63);
        CTX_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(int), ctx->stream));
        CTX_CUDA(ctx, launch_decode(ctx, FL_CLASS[k], d_dec_units.as<DecUnitDev>(), d_inv_units.as<InvUnitDev>(), dl + o,
                                    (int)cache->fl_n[k], d_err.as<int>(), counter, false));
        o += cache->fl_n[k];
    }
    return WC_OK;
}

// Segment tables of one slab-decoded class list with the streamed index kernel (k_seg_index3).
static int build_tables_streamed(wc_ctx* ctx, int fused_cls, const int* d_list, int n, const DecUnitDev* d_dec,
                                 const InvUnitDev* d_inv, int* d_err, int s_rt = 0) {
    CTX_CUDA(ctx, ctx->d_counter.reserve(64 * sizeof(int)));
    int* counter = ctx->d_counter.as<int>() + (ctx->counter_next++ & 63);
    CTX_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(int), ctx->stream));
    CTX_CUDA(ctx, launch_seg_index3(fused_cls, d_dec, d_inv, d_list, n, counter, d_err, ctx->sm_count, ctx->stream, &ctx->ls, s_rt));
    return WC_OK;
}

static int run_decompress(wc_ctx* ctx, const std::vector<DecJob>& jobs, DevBuf& d_coef,
                          DevBuf& d_dec_units, DevBuf& d_inv_units, DevBuf& d_inv_tiles,
                          DevBuf& d_ptiles, DevBuf& d_psum, DevBuf& d_err, DevBuf& d_fused_list,
                          DecCache* cache = nullptr) {
    int n = (int)jobs.size();
    if (n == 0) return WC_OK;
    if (cache) cache->valid = false;
    std::vector<DecUnitDev> du(n);
    std::vector<InvUnitDev> iu(n);
    std::vector<int2> ptiles, xtiles;
    std::vector<int> fl[FL_N];   // fused classes, FL_CLASS order (cluster kernels first)
    size_t coef_floats = 0;
    std::vector<size_t> coef_off(n);
    std::vector<char> slab_tab(n, 0);      // slab-decoded unit whose segment table came with the job
    bool build_tables[FL_N] = {};
    for (int i = 0; i < n; ++i) {
        long long total = (long long)jobs[i].nx * jobs[i].ny * jobs[i].nz;
        int cls = total > 0 ? fused_decode_class(jobs[i].nx, jobs[i].ny, jobs[i].nz, jobs[i].out_dtype, jobs[i].out_dev) : -1;
        if (ctx->opt_path == 1 && cls > 0) cls = 0;
        if (ctx->opt_path == 2 && cls == 0) return WC_ERR_BAD_DIMS;
        coef_off[i] = coef_floats;
        if (cls == 0) coef_floats += align_up((size_t)total, 4);
        else if (cls > 0 && !jobs[i].segtab && fused_decode_needs_table(cls))
            coef_floats += align_up(2 * fused_decode_table_entries(cls, jobs[i].nx, jobs[i].ny, jobs[i].nz), 4);   // int2 segment table
        du[i].pairs      = jobs[i].pairs_dev;
        du[i].npairs_dev = jobs[i].npairs_dev;
        du[i].npairs     = jobs[i].npairs;
        du[i].total      = (int32_t)total;
        du[i].ptile0     = (int32_t)ptiles.size();
        du[i].nptiles    = 0;
        iu[i].nx = jobs[i].nx; iu[i].ny = jobs[i].ny; iu[i].nz = jobs[i].nz;
        iu[i].out = jobs[i].out_dev;
        iu[i].dtype = jobs[i].out_dtype;
        if (cls == 0) {
            du[i].nptiles = ptile_count(jobs[i].npairs);
            for (int t = 0; t < du[i].nptiles; ++t) ptiles.push_back(make_int2(i, t));
            int nt = xtile_count(jobs[i].nx, jobs[i].ny, jobs[i].nz);
            for (int t = 0; t < nt; ++t) xtiles.push_back(make_int2(i, t));
        } else {
            for (int k = 0; k < FL_N; ++k)
                if (cls == FL_CLASS[k]) {
                    fl[k].push_back(i);
                    if (jobs[i].segtab) slab_tab[i] = 1;                       // table came with the job
                    else if (fused_decode_needs_table(cls)) build_tables[k] = true;
                    else slab_tab[i] = 2;                                      // block-wide scan: no table
                }
        }
    }
    auto job_dims = [&](int i, int& nx, int& ny, int& nz) { nx = jobs[i].nx; ny = jobs[i].ny; nz = jobs[i].nz; };
    sort_big_list(fl[FL_BIG], job_dims);
    CTX_CUDA(ctx, d_coef.reserve(sizeof(float) * std::max<size_t>(coef_floats, 4)));
    for (int i = 0; i < n; ++i) {
        du[i].coef = d_coef.as<float>() + coef_off[i];
        if (slab_tab[i]) du[i].coef = reinterpret_cast<float*>(const_cast<int2*>(jobs[i].segtab));   // nullptr for 2
        iu[i].coef = du[i].coef;
    }
    CTX_CUDA(ctx, d_dec_units.reserve(sizeof(DecUnitDev) * n));
    CTX_CUDA(ctx, d_inv_units.reserve(sizeof(InvUnitDev) * n));
    CTX_CUDA(ctx, d_inv_tiles.reserve(sizeof(int2) * std::max<size_t>(xtiles.size(), 1)));
    CTX_CUDA(ctx, d_ptiles.reserve(sizeof(int2) * std::max<size_t>(ptiles.size(), 1)));
    CTX_CUDA(ctx, d_psum.reserve(sizeof(long long) * std::max<size_t>(ptiles.size(), 1)));
    CTX_CUDA(ctx, d_err.reserve(64));
    CTX_CUDA(ctx, cudaMemcpyAsync(d_dec_units.p, du.data(), sizeof(DecUnitDev) * n,
                                  cudaMemcpyHostToDevice, ctx->stream));
    CTX_CUDA(ctx, cudaMemcpyAsync(d_inv_units.p, iu.data(), sizeof(InvUnitDev) * n,
                                  cudaMemcpyHostToDevice, ctx->stream));
    if (!xtiles.empty())
        CTX_CUDA(ctx, cudaMemcpyAsync(d_inv_tiles.p, xtiles.data(), sizeof(int2) * xtiles.size(),
                                      cudaMemcpyHostToDevice, ctx->stream));
    if (!ptiles.empty())
        CTX_CUDA(ctx, cudaMemcpyAsync(d_ptiles.p, ptiles.data(), sizeof(int2) * ptiles.size(),
                                      cudaMemcpyHostToDevice, ctx->stream));
    CTX_CUDA(ctx, cudaMemsetAsync(d_err.p, 0, 64, ctx->stream));
    if (coef_floats) CTX_CUDA(ctx, cudaMemsetAsync(d_coef.p, 0, sizeof(float) * coef_floats, ctx->stream));
    CTX_CUDA(ctx, launch_rle_decode_generic(d_dec_units.as<DecUnitDev>(), n, d_ptiles.as<int2>(),
                                            (int)ptiles.size(), d_psum.as<long long>(),
                                            d_err.as<int>(), ctx->stream, &ctx->ls));
    CTX_CUDA(ctx, launch_inverse_generic(d_inv_units.as<InvUnitDev>(), d_inv_tiles.as<int2>(),
                                         (int)xtiles.size(), ctx->stream, &ctx->ls));
    size_t n_fused = 0;
    for (int k = 0; k < FL_N; ++k) n_fused += fl[k].size();
    if (n_fused) {
        CTX_CUDA(ctx, d_fused_list.reserve(sizeof(int) * n_fused));
        CTX_CUDA(ctx, ctx->d_counter.reserve(64 * sizeof(int)));
        int* dl = d_fused_list.as<int>();
        size_t o = 0;
        // every fused decode kernel is a persistent one-CTA-per-SM kernel with dynamic item hand-out
        for (int k = 0; k < FL_N; ++k) {
            if (fl[k].empty()) continue;
            CTX_CUDA(ctx, cudaMemcpyAsync(dl + o, fl[k].data(), sizeof(int) * fl[k].size(),
                                          cudaMemcpyHostToDevice, ctx->stream));
            // every unit of a slab-decoded class arrives either with or without its table; the streamed
            // index only handles lists where all do without (mixed lists keep the one-CTA-per-unit kernel)
            bool all_without = true;
            for (int i : fl[k]) all_without = all_without && !slab_tab[i];
            int rc = for_each_run(FL_CLASS[k], fl[k], job_dims, [&](size_t a, size_t cnt, int s_rt) -> int {
                bool v1_tables = build_tables[k];
                if (build_tables[k] && ctx->opt_seg_index == 0 && all_without) {
                    int rc2 = build_tables_streamed(ctx, FL_CLASS[k], dl + o + a, (int)cnt, d_dec_units.as<DecUnitDev>(),
                                                    d_inv_units.as<InvUnitDev>(), d_err.as<int>(), s_rt);
                    if (rc2 != WC_OK) return rc2;
                    v1_tables = false;
                }
                int* counter = ctx->d_counter.as<int>() + (ctx->counter_next++ & 63);
                CTX_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(int), ctx->stream));
                CTX_CUDA(ctx, launch_decode(ctx, FL_CLASS[k], d_dec_units.as<DecUnitDev>(), d_inv_units.as<InvUnitDev>(), dl + o + a,
                                            (int)cnt, d_err.as<int>(), counter, v1_tables, s_rt));
                return WC_OK;
            });
            if (rc != WC_OK) return rc;
            o += fl[k].size();
        }
    }
    if (cache) {
        bool ok = coef_floats == 0 && ptiles.empty() && xtiles.empty() && fl[FL_BIG].empty();
        for (int k = 0; k < FL_N; ++k) {
            cache->fl_n[k] = fl[k].size();
            ok = ok && !build_tables[k];
        }
        cache->valid = ok;
    }
    return WC_OK;
}

int wc_plan_decompress(wc_plan* p, const wc_box_out* out, int out_space) {
    if (!p || (p->n_units > 0 && !out) || (out_space != WC_HOST && out_space != WC_DEVICE))
        return WC_ERR_INVALID_ARG;
    if (!p->compressed) return WC_ERR_STATE;
    wc_ctx* ctx = p->ctx;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    int n = p->n_units;
    if (!p->dec_cache) p->dec_cache = new DecCache();
    if (out_space == WC_DEVICE && n > 0 && p->dec_cache->valid && p->dec_cache->last_out.size() == (size_t)n &&
        std::memcmp(p->dec_cache->last_out.data(), out, sizeof(wc_box_out) * (size_t)n) == 0)
        return relaunch_decompress(ctx, p->dec_cache, p->d_dec_units, p->d_inv_units, p->d_err, p->d_dec_list);
    size_t stage = 0;
    std::vector<size_t> stage_off(n);
    uint64_t h1 = 0x243F6A8885A308D3ull ^ (uint64_t)out_space, h2 = 0x13198A2E03707344ull + (uint64_t)n;
    for (int i = 0; i < n; ++i) {
        const UnitDev& u = p->h_units[i];
        if (out[i].nx != u.nx || out[i].ny != u.ny || out[i].nz != u.nz) return WC_ERR_INVALID_ARG;
        if (out[i].dtype != WC_F32 && out[i].dtype != WC_F64) return WC_ERR_INVALID_ARG;
        if (u.n > 0 && !out[i].data) return WC_ERR_INVALID_ARG;
        stage_off[i] = stage;
        stage += align_up((size_t)u.n * dtype_size(out[i].dtype), 256);
        dec_hash(h1, h2, out_space == WC_DEVICE ? (uint64_t)(uintptr_t)out[i].data : 0);
        dec_hash(h1, h2, (uint64_t)out[i].dtype);
    }
    if (out_space == WC_HOST) CTX_CUDA(ctx, p->d_stage_out.reserve(std::max<size_t>(stage, 256)));
    dec_hash(h1, h2, (uint64_t)(uintptr_t)p->d_stage_out.p);
    const bool hit = p->dec_cache->valid && p->dec_cache->h1 == h1 && p->dec_cache->h2 == h2;
    std::vector<DecJob> jobs(hit ? 0 : n);
    if (!hit)
    for (int i = 0; i < n; ++i) {
        const UnitDev& u = p->h_units[i];
        jobs[i].pairs_dev  = u.out;
        jobs[i].npairs_dev = &p->d_states.as<UnitState>()[i].npairs;
        jobs[i].npairs     = u.n; // bound
        jobs[i].nx = u.nx; jobs[i].ny = u.ny; jobs[i].nz = u.nz;
        jobs[i].out_dev   = out_space == WC_HOST ? (void*)(p->d_stage_out.as<char>() + stage_off[i]) : out[i].data;
        jobs[i].out_dtype = out[i].dtype;
        if (p->has_segtab[i] && u.coef) jobs[i].segtab = reinterpret_cast<const int2*>(u.coef);
    }
    int rc = hit ? relaunch_decompress(ctx, p->dec_cache, p->d_dec_units, p->d_inv_units, p->d_err, p->d_dec_list)
                 : run_decompress(ctx, jobs, ctx->ws_coef, p->d_dec_units, p->d_inv_units, p->d_inv_tiles,
                                  p->d_ptiles, p->d_psum, p->d_err, p->d_dec_list, p->dec_cache);
    if (rc != WC_OK) return rc;
    p->dec_cache->h1 = h1; p->dec_cache->h2 = h2;
    if (out_space == WC_DEVICE) p->dec_cache->last_out.assign(out, out + n);
    else                        p->dec_cache->last_out.clear();
    if (out_space == WC_HOST) {
        CopyList cl;
        for (int i = 0; i < n; ++i)
            cl.add(out[i].data, p->d_stage_out.as<char>() + stage_off[i],
                   (size_t)p->h_units[i].n * dtype_size(out[i].dtype));
        for (const CopyRange& r : cl.r) {
            CTX_CUDA(ctx, cudaMemcpyAsync(r.dst, r.src, r.bytes, cudaMemcpyDeviceToHost, ctx->stream));
            ctx->d2h += r.bytes;
        }
    }
    return WC_OK;
}

int wc_decompress_batch(wc_ctx* ctx, const wc_packed* in, int n_units, int in_space,
                        const wc_box_out* out, int out_space) {
    if (!ctx || n_units < 0 || (n_units > 0 && (!in || !out)) ||
        (in_space != WC_HOST && in_space != WC_DEVICE) ||
        (out_space != WC_HOST && out_space != WC_DEVICE))
        return WC_ERR_INVALID_ARG;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<DecJob> jobs(n_units);
    size_t pair_total = 0, stage = 0;
    std::vector<size_t> pair_off(n_units), stage_off(n_units);
    for (int i = 0; i < n_units; ++i) {
        int rc = check_dims(in[i].shape[0], in[i].shape[1], in[i].shape[2]);
        if (rc != WC_OK) return rc;
        long long n = (long long)in[i].shape[0] * in[i].shape[1] * in[i].shape[2];
        if (in[i].npairs < 0 || in[i].ncoef < 0) return WC_ERR_CORRUPT;
        // more pairs than coefficients cannot come from the encoder (every pair is one kept coefficient):
        // reject it here, before its size drives an allocation or a copy (hostile / truncated headers)
        if (in[i].npairs > in[i].ncoef) return WC_ERR_CORRUPT;
        // The reference decodes into coeff_shape[0] floats and then reads shape[0]*shape[1]*shape[2]
        // of them (src/decompressor.cpp:245-251); a mismatch is out-of-bounds there, an error here.
        if ((long long)in[i].ncoef != n) return WC_ERR_CORRUPT;
        if (in[i].npairs > 0 && !in[i].pairs) return WC_ERR_INVALID_ARG;
        if (out[i].nx != in[i].shape[0] || out[i].ny != in[i].shape[1] || out[i].nz != in[i].shape[2])
            return WC_ERR_INVALID_ARG;
        if (out[i].dtype != WC_F32 && out[i].dtype != WC_F64) return WC_ERR_INVALID_ARG;
        if (n > 0 && !out[i].data) return WC_ERR_INVALID_ARG;
        pair_off[i] = pair_total;
        pair_total += (size_t)in[i].npairs;
        stage_off[i] = stage;
        stage += align_up((size_t)n * dtype_size(out[i].dtype), 256);
    }
    if (in_space == WC_HOST) {
        CTX_CUDA(ctx, ctx->ws_pairs.reserve(sizeof(wc_pair) * std::max<size_t>(pair_total, 1)));
        CopyList cl;
        for (int i = 0; i < n_units; ++i)
            cl.add(ctx->ws_pairs.as<wc_pair>() + pair_off[i], in[i].pairs, sizeof(wc_pair) * (size_t)in[i].npairs);
        for (const CopyRange& r : cl.r) {
            CTX_CUDA(ctx, cudaMemcpyAsync(r.dst, r.src, r.bytes, cudaMemcpyHostToDevice, ctx->stream));
            ctx->h2d += r.bytes;
        }
    }
    if (out_space == WC_HOST) CTX_CUDA(ctx, ctx->ws_boxes.reserve(std::max<size_t>(stage, 256)));
    for (int i = 0; i < n_units; ++i) {
        jobs[i].pairs_dev  = in_space == WC_HOST ? ctx->ws_pairs.as<wc_pair>() + pair_off[i] : in[i].pairs;
        jobs[i].npairs_dev = nullptr;
        jobs[i].npairs     = in[i].npairs;
        jobs[i].nx = in[i].shape[0]; jobs[i].ny = in[i].shape[1]; jobs[i].nz = in[i].shape[2];
        jobs[i].out_dev   = out_space == WC_HOST ? (void*)(ctx->ws_boxes.as<char>() + stage_off[i]) : out[i].data;
        jobs[i].out_dtype = out[i].dtype;
    }
    int rc = run_decompress(ctx, jobs, ctx->ws_coef, ctx->ws_tbl0, ctx->ws_tbl1, ctx->ws_tiles0,
                            ctx->ws_tiles1, ctx->ws_sum, ctx->ws_tbl2, ctx->ws_misc);
    if (rc != WC_OK) return rc;
    if (out_space == WC_HOST) {
        CopyList cl;
        for (int i = 0; i < n_units; ++i) {
            long long n = (long long)in[i].shape[0] * in[i].shape[1] * in[i].shape[2];
            cl.add(out[i].data, ctx->ws_boxes.as<char>() + stage_off[i], (size_t)n * dtype_size(out[i].dtype));
        }
        for (const CopyRange& r : cl.r) {
            CTX_CUDA(ctx, cudaMemcpyAsync(r.dst, r.src, r.bytes, cudaMemcpyDeviceToHost, ctx->stream));
            ctx->d2h += r.bytes;
        }
    }
    int h_err = 0;
    CTX_CUDA(ctx, cudaMemcpyAsync(&h_err, ctx->ws_tbl2.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return h_err ? WC_ERR_CORRUPT : WC_OK;
}

// ---------------------------------------------------------------------------------------------
// decode plans: the `-d` path (src/decompressor.cpp:238-255) for a whole batch of packed units that arrive
// as ONE dense pair stream — the concatenation of the files' bytes [20, 20+8K), which is also what
// wc_plan_fetch(WC_HOST) / wc_plan_compress_to_host deliver — plus the per-unit pair counts.
// ---------------------------------------------------------------------------------------------
struct wc_dplan {
    wc_ctx* ctx       = nullptr;
    int     n         = 0;
    int     out_space = WC_DEVICE;
    std::vector<wc_box_out> outs;
    std::vector<size_t>     stage_off;
    size_t    stage_bytes = 0;
    long long total_n     = 0;
    std::vector<int> fl[FL_N];            // unit ids per fused class (ascending)
    bool      has_generic = false;        // some unit needs the generic kernels: decode falls back to run_decompress
    size_t    fl_off[FL_N] = {};          // offset of each class list inside d_lists
    size_t    tab_floats = 0;
    DevBuf d_dec, d_inv, d_lists, d_tab, d_err, d_stage_out, d_pairs, d_npairs, d_counter, d_chain;
    PinBuf h_err;
    unsigned counter_next = 0;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    std::vector<cudaEvent_t> ev;
    bool decoded = false;
    // generic fallback workspace
    DevBuf g_coef, g_dec, g_inv, g_tiles, g_ptiles, g_psum, g_list;
    std::vector<int32_t> h_npairs;
};

int wc_dplan_destroy(wc_dplan* dp);

int wc_dplan_create(wc_ctx* ctx, const wc_box_out* outs, int n_units, int out_space, wc_dplan** out) {
    if (!ctx || !out || n_units < 0 || (n_units > 0 && !outs) || (out_space != WC_HOST && out_space != WC_DEVICE))
        return WC_ERR_INVALID_ARG;
    *out = nullptr;
    for (int i = 0; i < n_units; ++i) {
        int rc = check_dims(outs[i].nx, outs[i].ny, outs[i].nz);
        if (rc != WC_OK) return rc;
        if (outs[i].dtype != WC_F32 && outs[i].dtype != WC_F64) return WC_ERR_INVALID_ARG;
        if ((long long)outs[i].nx * outs[i].ny * outs[i].nz > 0 && !outs[i].data) return WC_ERR_INVALID_ARG;
    }
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    wc_dplan* dp = new (std::nothrow) wc_dplan();
    if (!dp) return WC_ERR_OOM;
    dp->ctx = ctx; dp->n = n_units; dp->out_space = out_space;
    dp->outs.assign(outs, outs + n_units);
    dp->stage_off.resize(n_units);
    for (int i = 0; i < n_units; ++i) {
        const long long n = (long long)outs[i].nx * outs[i].ny * outs[i].nz;
        dp->total_n += n;
        dp->stage_off[i] = dp->stage_bytes;
        dp->stage_bytes += align_up((size_t)n * dtype_size(outs[i].dtype), 256);
    }
    auto fail = [&](cudaError_t e, const char* what) {
        ctx->last_error = std::string(what) + ": " + cudaGetErrorString(e);
        cudaGetLastError();
        wc_dplan_destroy(dp);
        return e == cudaErrorMemoryAllocation ? WC_ERR_OOM : WC_ERR_CUDA;
    };
    cudaError_t e;
#define DP_RESERVE(buf, bytes) if ((e = (buf).reserve(bytes)) != cudaSuccess) return fail(e, "dplan alloc " #buf)
    if (out_space == WC_HOST) DP_RESERVE(dp->d_stage_out, std::max<size_t>(dp->stage_bytes, 256));
    std::vector<DecUnitDev> du(std::max(n_units, 1));
    std::vector<InvUnitDev> iu(std::max(n_units, 1));
    std::vector<size_t> tab_off(n_units, 0);
    std::vector<int> cls_of(n_units, -1);
    for (int i = 0; i < n_units; ++i) {
        const wc_box_out& o = outs[i];
        const long long n = (long long)o.nx * o.ny * o.nz;
        void* dev_out = out_space == WC_HOST ? (void*)(dp->d_stage_out.as<char>() + dp->stage_off[i]) : o.data;
        int cls = n > 0 ? fused_decode_class(o.nx, o.ny, o.nz, o.dtype, dev_out) : -1;
        if (ctx->opt_path == 1 && cls > 0) cls = 0;
        if (ctx->opt_path == 2 && cls == 0) { wc_dplan_destroy(dp); return WC_ERR_BAD_DIMS; }
        cls_of[i] = cls;
        std::memset(&du[i], 0, sizeof(DecUnitDev));
        du[i].total = (int32_t)n;
        iu[i].coef = nullptr; iu[i].out = dev_out;
        iu[i].nx = o.nx; iu[i].ny = o.ny; iu[i].nz = o.nz; iu[i].dtype = o.dtype;
        if (cls == 0) dp->has_generic = true;
        for (int k = 0; k < FL_N; ++k)
            if (cls == FL_CLASS[k]) {
                dp->fl[k].push_back(i);
                if (fused_decode_needs_table(cls)) {
                    tab_off[i] = dp->tab_floats;
                    dp->tab_floats += align_up(2 * fused_decode_table_entries(cls, o.nx, o.ny, o.nz), 4);
                }
            }
    }
    DP_RESERVE(dp->d_dec, sizeof(DecUnitDev) * std::max(n_units, 1));
    DP_RESERVE(dp->d_inv, sizeof(InvUnitDev) * std::max(n_units, 1));
    DP_RESERVE(dp->d_tab, sizeof(float) * std::max<size_t>(dp->tab_floats, 4));
    DP_RESERVE(dp->d_err, 64);
    DP_RESERVE(dp->d_npairs, sizeof(int32_t) * std::max(n_units, 1));
    DP_RESERVE(dp->d_counter, 64 * sizeof(int));
    DP_RESERVE(dp->d_chain, sizeof(unsigned long long) * ((size_t)(n_units + 1023) / 1024 + 2));
    if ((e = dp->h_err.reserve(64)) != cudaSuccess) return fail(e, "dplan pinned alloc");
    size_t n_fused = 0;
    std::vector<int> lists;
    sort_big_list(dp->fl[FL_BIG], [&](int i, int& nx, int& ny, int& nz) { nx = outs[i].nx; ny = outs[i].ny; nz = outs[i].nz; });
    for (int k = 0; k < FL_N; ++k) {
        dp->fl_off[k] = n_fused;
        n_fused += dp->fl[k].size();
        lists.insert(lists.end(), dp->fl[k].begin(), dp->fl[k].end());
    }
    for (int i = 0; i < n_units; ++i)
        if (cls_of[i] > 0 && fused_decode_needs_table(cls_of[i])) {
            du[i].coef = dp->d_tab.as<float>() + tab_off[i];
            iu[i].coef = du[i].coef;
        }
    DP_RESERVE(dp->d_lists, sizeof(int) * std::max<size_t>(n_fused, 1));
#undef DP_RESERVE
    if (n_units) {
        if ((e = cudaMemcpyAsync(dp->d_dec.p, du.data(), sizeof(DecUnitDev) * n_units, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return fail(e, "dplan upload");
        if ((e = cudaMemcpyAsync(dp->d_inv.p, iu.data(), sizeof(InvUnitDev) * n_units, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return fail(e, "dplan upload");
    }
    if (n_fused)
        if ((e = cudaMemcpyAsync(dp->d_lists.p, lists.data(), sizeof(int) * n_fused, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return fail(e, "dplan upload");
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return fail(e, "dplan sync");
    *out = dp;
    return WC_OK;
}

int wc_dplan_destroy(wc_dplan* dp) {
    if (!dp) return WC_OK;
    wc_ctx* ctx = dp->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (dp->s_h2d) { cudaStreamSynchronize(dp->s_h2d); cudaStreamDestroy(dp->s_h2d); }
    if (dp->s_d2h) { cudaStreamSynchronize(dp->s_d2h); cudaStreamDestroy(dp->s_d2h); }
    for (cudaEvent_t e : dp->ev) cudaEventDestroy(e);
    DevBuf* bufs[] = { &dp->d_dec, &dp->d_inv, &dp->d_lists, &dp->d_tab, &dp->d_err,
                       &dp->d_stage_out, &dp->d_pairs, &dp->d_npairs, &dp->d_counter, &dp->d_chain, &dp->g_coef,
                       &dp->g_dec, &dp->g_inv, &dp->g_tiles, &dp->g_ptiles, &dp->g_psum, &dp->g_list };
    for (DevBuf* b : bufs) b->release();
    dp->h_err.release();
    cudaGetLastError();
    delete dp;
    return WC_OK;
}

// kernels of the units [u0, u1) (all of them: 0, n) on the ctx stream; the dense stream and the counts are on
// the device, k_dec_prepare has run
static int dplan_launch_range(wc_dplan* dp, int u0, int u1, size_t fi[FL_N]) {
    wc_ctx* ctx = dp->ctx;
    int tl = 0;
    for (int k = 0; k < FL_N; ++k) {
        if (dp->fl[k].empty()) continue;
        const bool tab = fused_decode_needs_table(FL_CLASS[k]);
        size_t j = fi[k];
        while (j < dp->fl[k].size() && dp->fl[k][j] < u1) ++j;
        if (j > fi[k]) {
            const std::vector<int> sub(dp->fl[k].begin() + fi[k], dp->fl[k].begin() + j);
            auto dims = [&](int i, int& nx, int& ny, int& nz) { nx = dp->outs[i].nx; ny = dp->outs[i].ny; nz = dp->outs[i].nz; };
            int rc = for_each_run(FL_CLASS[k], sub, dims, [&](size_t a, size_t cnt, int s_rt) -> int {
                const int* list = dp->d_lists.as<int>() + dp->fl_off[k] + fi[k] + a;
                const int  nl   = (int)cnt;
                bool v1 = false;
                if (tab) {
                    if (ctx->opt_seg_index == 0) {
                        int* counter = dp->d_counter.as<int>() + (dp->counter_next++ & 63);
                        CTX_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(int), ctx->stream));
                        CTX_CUDA(ctx, launch_seg_index3(FL_CLASS[k], dp->d_dec.as<DecUnitDev>(), dp->d_inv.as<InvUnitDev>(), list, nl,
                                                        counter, dp->d_err.as<int>(), ctx->sm_count, ctx->stream, &ctx->ls, s_rt));
                    } else {
                        v1 = true;
                    }
                }
                int* counter = dp->d_counter.as<int>() + (dp->counter_next++ & 63);
                CTX_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(int), ctx->stream));
                CTX_CUDA(ctx, launch_decode(ctx, FL_CLASS[k], dp->d_dec.as<DecUnitDev>(), dp->d_inv.as<InvUnitDev>(), list, nl,
                                            dp->d_err.as<int>(), counter, v1, s_rt));
                return WC_OK;
            });
            if (rc != WC_OK) return rc;
        }
        fi[k] = j;
        if (tab) ++tl;
    }
    (void)tl;
    return WC_OK;
}

static int dplan_generic_fallback(wc_dplan* dp, const wc_pair* dense_dev, const int32_t* npairs_host) {
    wc_ctx* ctx = dp->ctx;
    std::vector<DecJob> jobs(dp->n);
    size_t off = 0;
    for (int i = 0; i < dp->n; ++i) {
        const wc_box_out& o = dp->outs[i];
        const long long n = (long long)o.nx * o.ny * o.nz;
        if (npairs_host[i] < 0 || npairs_host[i] > n) return WC_ERR_CORRUPT;
        jobs[i].pairs_dev = dense_dev + off;
        jobs[i].npairs_dev = nullptr;
        jobs[i].npairs = npairs_host[i];
        jobs[i].nx = o.nx; jobs[i].ny = o.ny; jobs[i].nz = o.nz;
        jobs[i].out_dev = dp->out_space == WC_HOST ? (void*)(dp->d_stage_out.as<char>() + dp->stage_off[i]) : o.data;
        jobs[i].out_dtype = o.dtype;
        off += (size_t)npairs_host[i];
    }
    return run_decompress(ctx, jobs, dp->g_coef, dp->g_dec, dp->g_inv, dp->g_tiles, dp->g_ptiles, dp->g_psum, dp->d_err, dp->g_list);
}

int wc_dplan_decode(wc_dplan* dp, const wc_pair* pairs, const int32_t* npairs, int in_space) {
    if (!dp || (in_space != WC_HOST && in_space != WC_DEVICE) || (dp->n > 0 && !npairs)) return WC_ERR_INVALID_ARG;
    wc_ctx* ctx = dp->ctx;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    dp->decoded = false;
    const int n = dp->n;
    if (n == 0) { dp->decoded = true; return WC_OK; }
    const int NCH = 8;
    if (!dp->s_h2d && (in_space == WC_HOST || dp->out_space == WC_HOST)) {
        CTX_CUDA(ctx, cudaStreamCreateWithFlags(&dp->s_h2d, cudaStreamNonBlocking));
        CTX_CUDA(ctx, cudaStreamCreateWithFlags(&dp->s_d2h, cudaStreamNonBlocking));
        dp->ev.resize(2 * NCH + 2);
        for (cudaEvent_t& e : dp->ev) CTX_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    CTX_CUDA(ctx, cudaMemsetAsync(dp->d_err.p, 0, 64, ctx->stream));
    // host-side counts: needed to size the H2D of the stream (host input) and by the generic fallback
    const int32_t* np_host = nullptr;
    if (in_space == WC_HOST) np_host = npairs;
    else if (dp->has_generic) {
        dp->h_npairs.resize(n);
        CTX_CUDA(ctx, cudaMemcpyAsync(dp->h_npairs.data(), npairs, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
        CTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->d2h += sizeof(int32_t) * n;
        np_host = dp->h_npairs.data();
    }
    std::vector<size_t> pair_off;
    size_t total_pairs = 0;
    if (np_host) {
        pair_off.resize(n + 1);
        for (int i = 0; i < n; ++i) {
            pair_off[i] = total_pairs;
            if (np_host[i] < 0 || (long long)np_host[i] > (long long)dp->outs[i].nx * dp->outs[i].ny * dp->outs[i].nz)
                return WC_ERR_CORRUPT;
            total_pairs += (size_t)np_host[i];
        }
        pair_off[n] = total_pairs;
        if (total_pairs > 0 && !pairs) return WC_ERR_INVALID_ARG;
    }
    const wc_pair*  d_pairs  = pairs;
    const int32_t*  d_npairs = npairs;
    // (the slab-count runs of the big-box class are not in unit order: no chunking by unit range when there are any)
    const bool pipelined = in_space == WC_HOST && !dp->has_generic && dp->fl[FL_BIG].empty();
    if (in_space == WC_HOST) {
        CTX_CUDA(ctx, dp->d_pairs.reserve(sizeof(wc_pair) * std::max<size_t>(total_pairs, 1)));
        d_pairs  = dp->d_pairs.as<wc_pair>();
        d_npairs = dp->d_npairs.as<int32_t>();
        CTX_CUDA(ctx, cudaMemcpyAsync(dp->d_npairs.p, npairs, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
        ctx->h2d += sizeof(int32_t) * n;
        if (!pipelined && total_pairs) {
            CTX_CUDA(ctx, cudaMemcpyAsync(dp->d_pairs.p, pairs, sizeof(wc_pair) * total_pairs, cudaMemcpyHostToDevice, ctx->stream));
            ctx->h2d += sizeof(wc_pair) * total_pairs;
        }
    }
    if (dp->has_generic) {
        int rc = dplan_generic_fallback(dp, d_pairs, np_host);
        if (rc != WC_OK) return rc;
    } else {
        if (dp->tab_floats) CTX_CUDA(ctx, cudaMemsetAsync(dp->d_tab.p, 0, sizeof(float) * dp->tab_floats, ctx->stream));
        CTX_CUDA(ctx, launch_dec_prepare(dp->d_dec.as<DecUnitDev>(), n, d_pairs, d_npairs, dp->d_chain.as<unsigned long long>(),
                                         dp->d_err.as<int>(), ctx->stream, &ctx->ls));
    }
    size_t fi[FL_N] = {};
    if (!pipelined) {
        if (!dp->has_generic) {
            int rc = dplan_launch_range(dp, 0, n, fi);
            if (rc != WC_OK) return rc;
        }
        if (dp->out_space == WC_HOST) {
            CopyList cl;
            for (int i = 0; i < n; ++i)
                cl.add(dp->outs[i].data, dp->d_stage_out.as<char>() + dp->stage_off[i],
                       (size_t)dp->outs[i].nx * dp->outs[i].ny * dp->outs[i].nz * dtype_size(dp->outs[i].dtype));
            for (const CopyRange& r : cl.r) {
                CTX_CUDA(ctx, cudaMemcpyAsync(r.dst, r.src, r.bytes, cudaMemcpyDeviceToHost, ctx->stream));
                ctx->d2h += r.bytes;
            }
        }
    } else {
        // host stream in (and usually host boxes out): chunk c's pairs go H2D on one stream while chunk c-1 is
        // decoded on the ctx stream and chunk c-2's boxes go D2H on a third (PCIe is full duplex)
        std::vector<int> cut(NCH + 1, n);
        cut[0] = 0;
        {
            // chunk boundaries by output bytes (the larger direction)
            size_t acc = 0, per = dp->stage_bytes / NCH + 1;
            int c = 1;
            for (int i = 0; i < n && c < NCH; ++i) {
                acc += align_up((size_t)dp->outs[i].nx * dp->outs[i].ny * dp->outs[i].nz * dtype_size(dp->outs[i].dtype), 256);
                if (acc >= per * c) cut[c++] = i + 1;
            }
        }
        CTX_CUDA(ctx, cudaEventRecord(dp->ev[2 * NCH], ctx->stream));          // prepare done (needs only the counts)
        CTX_CUDA(ctx, cudaStreamWaitEvent(dp->s_h2d, dp->ev[2 * NCH], 0));
        for (int c = 0; c < NCH; ++c) {
            const int u0 = cut[c], u1 = cut[c + 1];
            const size_t p0 = pair_off[u0], p1 = pair_off[u1];
            if (p1 > p0) {
                CTX_CUDA(ctx, cudaMemcpyAsync(dp->d_pairs.as<wc_pair>() + p0, pairs + p0, sizeof(wc_pair) * (p1 - p0),
                                              cudaMemcpyHostToDevice, dp->s_h2d));
                ctx->h2d += sizeof(wc_pair) * (p1 - p0);
            }
            CTX_CUDA(ctx, cudaEventRecord(dp->ev[c], dp->s_h2d));
            CTX_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, dp->ev[c], 0));
            int rc = dplan_launch_range(dp, u0, u1, fi);
            if (rc != WC_OK) return rc;
            CTX_CUDA(ctx, cudaEventRecord(dp->ev[NCH + c], ctx->stream));
            if (dp->out_space == WC_HOST && u1 > u0) {
                CTX_CUDA(ctx, cudaStreamWaitEvent(dp->s_d2h, dp->ev[NCH + c], 0));
                CopyList cl;
                for (int i = u0; i < u1; ++i)
                    cl.add(dp->outs[i].data, dp->d_stage_out.as<char>() + dp->stage_off[i],
                           (size_t)dp->outs[i].nx * dp->outs[i].ny * dp->outs[i].nz * dtype_size(dp->outs[i].dtype));
                for (const CopyRange& r : cl.r) {
                    CTX_CUDA(ctx, cudaMemcpyAsync(r.dst, r.src, r.bytes, cudaMemcpyDeviceToHost, dp->s_d2h));
                    ctx->d2h += r.bytes;
                }
            }
        }
        if (dp->out_space == WC_HOST) {
            CTX_CUDA(ctx, cudaEventRecord(dp->ev[2 * NCH + 1], dp->s_d2h));
            CTX_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, dp->ev[2 * NCH + 1], 0));   // ctx stream = everything done
        }
    }
    CTX_CUDA(ctx, cudaMemcpyAsync(dp->h_err.p, dp->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    dp->decoded = true;
    return WC_OK;
}

int wc_dplan_finish(wc_dplan* dp) {
    if (!dp) return WC_ERR_INVALID_ARG;
    if (!dp->decoded) return WC_ERR_STATE;
    wc_ctx* ctx = dp->ctx;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    CTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (dp->n == 0) return WC_OK;
    return *dp->h_err.as<int>() ? WC_ERR_CORRUPT : WC_OK;
}

// per-unit results of the last compress that the reference derives on the host: min / max of the narrowed
// values (src/preprocess.cpp:82-88; needs WC_OPT_INGEST_STATS = 1 before the compress) and need32
// (src/compressor.cpp:224-229).  Any of the three arrays may be null.
int wc_plan_unit_stats(wc_plan* p, float* mins, float* maxs, int32_t* need32) {
    if (!p) return WC_ERR_INVALID_ARG;
    if (!p->compressed) return WC_ERR_STATE;
    wc_ctx* ctx = p->ctx;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((mins || maxs) && !p->has_stats) return WC_ERR_STATE;
    int rc = plan_read_states(p);
    if (rc != WC_OK) return rc;
    const UnitState* hs = p->h_states.as<UnitState>();
    auto decode = [](uint32_t code) -> float {        // inverse of float_order_code
        const uint32_t b = (code & 0x80000000u) ? (code & 0x7fffffffu) : ~code;
        float f;
        std::memcpy(&f, &b, 4);
        return f;
    };
    const float inf = __builtin_inff();
    for (int i = 0; i < p->n_units; ++i) {
        uint32_t cmin, cmax;
        std::memcpy(&cmin, &hs[i].vmin, 4);
        std::memcpy(&cmax, &hs[i].vmax, 4);
        if (mins) mins[i] = cmin ? decode(~cmin) : inf;
        if (maxs) maxs[i] = cmax ? decode(cmax) : -inf;
        if (need32) need32[i] = (hs[i].flags & UNIT_FLAG_NEED32) ? 1 : 0;
    }
    if ((mins || maxs) && !p->generic.empty()) {
        // units on the generic kernels: their own min / max pass (shapes no fused class takes)
        std::vector<wc_box_desc> b(p->generic.size());
        std::vector<float> lo(b.size()), hi(b.size());
        for (size_t j = 0; j < b.size(); ++j) {
            const UnitDev& u = p->h_units[p->generic[j]];
            b[j] = { u.in, u.dtype, u.nx, u.ny, u.nz };
        }
        rc = wc_minmax_batch(ctx, b.data(), (int)b.size(), WC_DEVICE, lo.data(), hi.data());
        if (rc != WC_OK) return rc;
        for (size_t j = 0; j < b.size(); ++j) {
            if (mins) mins[p->generic[j]] = lo[j];
            if (maxs) maxs[p->generic[j]] = hi[j];
        }
    }
    return WC_OK;
}

// ---------------------------------------------------------------------------------------------
// RMSE
// ---------------------------------------------------------------------------------------------
struct RmseJob {
    const void* a;
    const void* b;
    int32_t a_dtype, b_dtype;
    int32_t n;
};

// `cached_tiles`: in/out, the number of tiles already resident in d_tiles from an earlier call with the same
// unit sizes (a plan), or nullptr / -1 to build and upload the (unit, tile) table now.
static int run_rmse(wc_ctx* ctx, const std::vector<RmseJob>& jobs, DevBuf& d_units, DevBuf& d_tiles,
                    DevBuf& d_sum, DevBuf& d_rmse, double* rmse_host, long long* cached_tiles = nullptr) {
    int n = (int)jobs.size();
    if (n == 0) return WC_OK;
    const bool have_tiles = cached_tiles && *cached_tiles >= 0;
    std::vector<RmseUnitDev> ru(n);
    std::vector<int2> tiles;
    size_t n_tiles = 0;
    for (int i = 0; i < n; ++i) {
        ru[i].a = jobs[i].a; ru[i].b = jobs[i].b;
        ru[i].a_dtype = jobs[i].a_dtype; ru[i].b_dtype = jobs[i].b_dtype;
        ru[i].n = jobs[i].n;
        ru[i].ctile0  = (int32_t)n_tiles;
        ru[i].nctiles = ctile_count(jobs[i].n);
        if (!have_tiles)
            for (int t = 0; t < ru[i].nctiles; ++t) tiles.push_back(make_int2(i, t));
        n_tiles += (size_t)ru[i].nctiles;
    }
    CTX_CUDA(ctx, d_units.reserve(sizeof(RmseUnitDev) * n));
    CTX_CUDA(ctx, d_tiles.reserve(sizeof(int2) * std::max<size_t>(n_tiles, 1)));
    CTX_CUDA(ctx, d_sum.reserve(sizeof(double) * std::max<size_t>(n_tiles, 1)));
    CTX_CUDA(ctx, d_rmse.reserve(sizeof(double) * n));
    CTX_CUDA(ctx, cudaMemcpyAsync(d_units.p, ru.data(), sizeof(RmseUnitDev) * n,
                                  cudaMemcpyHostToDevice, ctx->stream));
    if (!have_tiles && n_tiles)
        CTX_CUDA(ctx, cudaMemcpyAsync(d_tiles.p, tiles.data(), sizeof(int2) * n_tiles,
                                      cudaMemcpyHostToDevice, ctx->stream));
    if (cached_tiles) *cached_tiles = (long long)n_tiles;
    CTX_CUDA(ctx, launch_rmse_generic(d_units.as<RmseUnitDev>(), n, d_tiles.as<int2>(),
                                      (int)n_tiles, d_sum.as<double>(), d_rmse.as<double>(),
                                      ctx->stream, &ctx->ls));
    CTX_CUDA(ctx, cudaMemcpyAsync(rmse_host, d_rmse.p, sizeof(double) * n, cudaMemcpyDeviceToHost,
                                  ctx->stream));
    ctx->d2h += sizeof(double) * n;
    CTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return WC_OK;
}

int wc_plan_rmse(wc_plan* p, const wc_box_desc* recon, double* rmse) {
    if (!p || (p->n_units > 0 && (!recon || !rmse))) return WC_ERR_INVALID_ARG;
    wc_ctx* ctx = p->ctx;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<RmseJob> jobs(p->n_units);
    for (int i = 0; i < p->n_units; ++i) {
        const UnitDev& u = p->h_units[i];
        if (recon[i].nx != u.nx || recon[i].ny != u.ny || recon[i].nz != u.nz) return WC_ERR_INVALID_ARG;
        if (recon[i].dtype != WC_F32 && recon[i].dtype != WC_F64) return WC_ERR_INVALID_ARG;
        jobs[i] = { u.in, recon[i].data, u.dtype, recon[i].dtype, u.n };
    }
    // the (unit, tile) table only depends on the unit sizes: uploaded once per plan
    return run_rmse(ctx, jobs, p->d_rmse_units, p->d_rmse_tiles, p->d_rmse_sum, p->d_rmse, rmse, &p->rmse_tiles);
}

int wc_rmse_batch(wc_ctx* ctx, const wc_box_desc* actual, const wc_box_desc* pred, int n_units,
                  int space, double* rmse) {
    if (!ctx || n_units < 0 || (n_units > 0 && (!actual || !pred || !rmse)) ||
        (space != WC_HOST && space != WC_DEVICE))
        return WC_ERR_INVALID_ARG;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<RmseJob> jobs(n_units);
    std::vector<size_t> off_a(n_units), off_b(n_units);
    size_t ta = 0, tb = 0;
    for (int i = 0; i < n_units; ++i) {
        int rc = check_dims(actual[i].nx, actual[i].ny, actual[i].nz);
        if (rc != WC_OK) return rc;
        if (pred[i].nx != actual[i].nx || pred[i].ny != actual[i].ny || pred[i].nz != actual[i].nz)
            return WC_ERR_INVALID_ARG;
        if ((actual[i].dtype != WC_F32 && actual[i].dtype != WC_F64) ||
            (pred[i].dtype != WC_F32 && pred[i].dtype != WC_F64))
            return WC_ERR_INVALID_ARG;
        size_t n = (size_t)actual[i].nx * actual[i].ny * actual[i].nz;
        if (n > 0 && (!actual[i].data || !pred[i].data)) return WC_ERR_INVALID_ARG;
        off_a[i] = ta; ta += align_up(n * dtype_size(actual[i].dtype), 256);
        off_b[i] = tb; tb += align_up(n * dtype_size(pred[i].dtype), 256);
    }
    if (space == WC_HOST) {
        CTX_CUDA(ctx, ctx->ws_a.reserve(std::max<size_t>(ta, 256)));
        CTX_CUDA(ctx, ctx->ws_b.reserve(std::max<size_t>(tb, 256)));
        CopyList ca, cb;
        for (int i = 0; i < n_units; ++i) {
            size_t n = (size_t)actual[i].nx * actual[i].ny * actual[i].nz;
            ca.add(ctx->ws_a.as<char>() + off_a[i], actual[i].data, n * dtype_size(actual[i].dtype));
            cb.add(ctx->ws_b.as<char>() + off_b[i], pred[i].data, n * dtype_size(pred[i].dtype));
        }
        for (const CopyList* cl : { &ca, &cb })
            for (const CopyRange& r : cl->r) {
                CTX_CUDA(ctx, cudaMemcpyAsync(r.dst, r.src, r.bytes, cudaMemcpyHostToDevice, ctx->stream));
                ctx->h2d += r.bytes;
            }
    }
    for (int i = 0; i < n_units; ++i) {
        size_t n = (size_t)actual[i].nx * actual[i].ny * actual[i].nz;
        jobs[i].a = space == WC_HOST ? (const void*)(ctx->ws_a.as<char>() + off_a[i]) : actual[i].data;
        jobs[i].b = space == WC_HOST ? (const void*)(ctx->ws_b.as<char>() + off_b[i]) : pred[i].data;
        jobs[i].a_dtype = actual[i].dtype;
        jobs[i].b_dtype = pred[i].dtype;
        jobs[i].n = (int32_t)n;
    }
    return run_rmse(ctx, jobs, ctx->ws_tbl0, ctx->ws_tiles0, ctx->ws_sum, ctx->ws_misc, rmse);
}


// ---------------------------------------------------------------------------------------------
// ingest statistics: per-unit min / max of the narrowed values
// ---------------------------------------------------------------------------------------------
int wc_minmax_batch(wc_ctx* ctx, const wc_box_desc* boxes, int n_units, int space, float* mins, float* maxs) {
    if (!ctx || n_units < 0 || (n_units > 0 && (!boxes || !mins || !maxs)) ||
        (space != WC_HOST && space != WC_DEVICE))
        return WC_ERR_INVALID_ARG;
    if (n_units == 0) return WC_OK;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<RmseUnitDev> ru(n_units);
    std::vector<int2> tiles;
    std::vector<size_t> off(n_units);
    size_t total = 0;
    for (int i = 0; i < n_units; ++i) {
        int rc = check_dims(boxes[i].nx, boxes[i].ny, boxes[i].nz);
        if (rc != WC_OK) return rc;
        if (boxes[i].dtype != WC_F32 && boxes[i].dtype != WC_F64) return WC_ERR_INVALID_ARG;
        size_t n = (size_t)boxes[i].nx * boxes[i].ny * boxes[i].nz;
        if (n > 0 && !boxes[i].data) return WC_ERR_INVALID_ARG;
        off[i] = total;
        total += align_up(n * dtype_size(boxes[i].dtype), 256);
    }
    if (space == WC_HOST) {
        CTX_CUDA(ctx, ctx->ws_a.reserve(std::max<size_t>(total, 256)));
        CopyList cl;
        for (int i = 0; i < n_units; ++i) {
            size_t n = (size_t)boxes[i].nx * boxes[i].ny * boxes[i].nz;
            cl.add(ctx->ws_a.as<char>() + off[i], boxes[i].data, n * dtype_size(boxes[i].dtype));
        }
        for (const CopyRange& r : cl.r) {
            CTX_CUDA(ctx, cudaMemcpyAsync(r.dst, r.src, r.bytes, cudaMemcpyHostToDevice, ctx->stream));
            ctx->h2d += r.bytes;
        }
    }
    for (int i = 0; i < n_units; ++i) {
        size_t n = (size_t)boxes[i].nx * boxes[i].ny * boxes[i].nz;
        ru[i].a = space == WC_HOST ? (const void*)(ctx->ws_a.as<char>() + off[i]) : boxes[i].data;
        ru[i].b = nullptr;
        ru[i].a_dtype = boxes[i].dtype; ru[i].b_dtype = WC_F32;
        ru[i].n = (int32_t)n;
        ru[i].ctile0  = (int32_t)tiles.size();
        ru[i].nctiles = ctile_count((long long)n);
        for (int t = 0; t < ru[i].nctiles; ++t) tiles.push_back(make_int2(i, t));
    }
    CTX_CUDA(ctx, ctx->ws_tbl0.reserve(sizeof(RmseUnitDev) * n_units));
    CTX_CUDA(ctx, ctx->ws_tiles0.reserve(sizeof(int2) * std::max<size_t>(tiles.size(), 1)));
    CTX_CUDA(ctx, ctx->ws_sum.reserve(sizeof(float2) * std::max<size_t>(tiles.size(), 1)));
    CTX_CUDA(ctx, ctx->ws_misc.reserve(sizeof(float2) * n_units));
    CTX_CUDA(ctx, cudaMemcpyAsync(ctx->ws_tbl0.p, ru.data(), sizeof(RmseUnitDev) * n_units, cudaMemcpyHostToDevice, ctx->stream));
    if (!tiles.empty())
        CTX_CUDA(ctx, cudaMemcpyAsync(ctx->ws_tiles0.p, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice, ctx->stream));
    CTX_CUDA(ctx, launch_minmax_generic(ctx->ws_tbl0.as<RmseUnitDev>(), n_units, ctx->ws_tiles0.as<int2>(), (int)tiles.size(),
                                        ctx->ws_sum.as<float2>(), ctx->ws_misc.as<float2>(), ctx->stream, &ctx->ls));
    std::vector<float2> mm(n_units);
    CTX_CUDA(ctx, cudaMemcpyAsync(mm.data(), ctx->ws_misc.p, sizeof(float2) * n_units, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->d2h += sizeof(float2) * n_units;
    CTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n_units; ++i) { mins[i] = mm[i].x; maxs[i] = mm[i].y; }
    return WC_OK;
}

// ---------------------------------------------------------------------------------------------
// blocking compress
// ---------------------------------------------------------------------------------------------
int wc_compress_batch(wc_ctx* ctx, const wc_box_desc* in, int n_units, int in_space, double keep,
                      int thresh_mode, wc_packed* out, int out_space) {
    if (!ctx || n_units < 0 || (n_units > 0 && (!in || !out))) return WC_ERR_INVALID_ARG;
    if (ctx->batch_plan) {
        wc_plan_destroy(ctx->batch_plan);
        ctx->batch_plan = nullptr;
    }
    wc_plan* p = nullptr;
    int rc = wc_plan_create(ctx, in, n_units, in_space, &p);
    if (rc != WC_OK) return rc;
    ctx->batch_plan = p;
    if (in_space == WC_HOST && out_space == WC_HOST && thresh_mode == WC_THRESH_PER_UNIT)
        return wc_plan_compress_to_host(p, keep, out);
    rc = wc_plan_compress(p, keep, thresh_mode);
    if (rc != WC_OK) return rc;
    return wc_plan_fetch(p, out, out_space);
}

// ---------------------------------------------------------------------------------------------
// un-fused primitives (always the generic kernels)
// ---------------------------------------------------------------------------------------------
static int stage_in(wc_ctx* ctx, DevBuf& buf, const void* src, size_t bytes, int space, const void** dev) {
    // always staged into an aligned internal buffer, so caller alignment never matters
    CTX_CUDA(ctx, buf.reserve(std::max<size_t>(bytes, 16)));
    if (bytes) {
        CTX_CUDA(ctx, cudaMemcpyAsync(buf.p, src, bytes,
                                      space == WC_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                                      ctx->stream));
        if (space == WC_HOST) ctx->h2d += bytes;
    }
    *dev = buf.p;
    return WC_OK;
}

static int stage_out(wc_ctx* ctx, void* dst, const void* dev, size_t bytes, int space) {
    if (bytes) {
        CTX_CUDA(ctx, cudaMemcpyAsync(dst, dev, bytes,
                                      space == WC_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice,
                                      ctx->stream));
        if (space == WC_HOST) ctx->d2h += bytes;
    }
    CTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return WC_OK;
}

int wc_haar_forward(wc_ctx* ctx, const wc_box_desc* in, int space, float* coef_out) {
    if (!ctx || !in || (space != WC_HOST && space != WC_DEVICE)) return WC_ERR_INVALID_ARG;
    int rc = check_dims(in->nx, in->ny, in->nz);
    if (rc != WC_OK) return rc;
    if (in->dtype != WC_F32 && in->dtype != WC_F64) return WC_ERR_INVALID_ARG;
    size_t n = (size_t)in->nx * in->ny * in->nz;
    if (n == 0) return WC_OK;
    if (!in->data || !coef_out) return WC_ERR_INVALID_ARG;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    const void* d_in;
    rc = stage_in(ctx, ctx->ws_a, in->data, n * dtype_size(in->dtype), space, &d_in);
    if (rc != WC_OK) return rc;
    CTX_CUDA(ctx, ctx->ws_coef.reserve(sizeof(float) * n));
    CTX_CUDA(ctx, ctx->ws_tbl0.reserve(sizeof(UnitDev)));
    CTX_CUDA(ctx, ctx->ws_tbl1.reserve(sizeof(UnitState)));
    UnitDev u;
    std::memset(&u, 0, sizeof(u));
    u.in = d_in; u.coef = ctx->ws_coef.as<float>(); u.nx = in->nx; u.ny = in->ny; u.nz = in->nz;
    u.n = (int32_t)n; u.dtype = in->dtype;
    int nt = xtile_count(in->nx, in->ny, in->nz);
    std::vector<int2> tiles(nt);
    for (int t = 0; t < nt; ++t) tiles[t] = make_int2(0, t);
    CTX_CUDA(ctx, ctx->ws_tiles0.reserve(sizeof(int2) * nt));
    CTX_CUDA(ctx, cudaMemcpyAsync(ctx->ws_tbl0.p, &u, sizeof(u), cudaMemcpyHostToDevice, ctx->stream));
    CTX_CUDA(ctx, cudaMemcpyAsync(ctx->ws_tiles0.p, tiles.data(), sizeof(int2) * nt, cudaMemcpyHostToDevice, ctx->stream));
    CTX_CUDA(ctx, cudaMemsetAsync(ctx->ws_tbl1.p, 0, sizeof(UnitState), ctx->stream));
    CTX_CUDA(ctx, launch_forward_generic(ctx->ws_tbl0.as<UnitDev>(), ctx->ws_tbl1.as<UnitState>(),
                                         ctx->ws_tiles0.as<int2>(), nt, ctx->stream, &ctx->ls));
    return stage_out(ctx, coef_out, ctx->ws_coef.p, sizeof(float) * n, space);
}

int wc_haar_inverse(wc_ctx* ctx, const float* coef, int nx, int ny, int nz, int space,
                    float* box_out) {
    if (!ctx || (space != WC_HOST && space != WC_DEVICE)) return WC_ERR_INVALID_ARG;
    int rc = check_dims(nx, ny, nz);
    if (rc != WC_OK) return rc;
    size_t n = (size_t)nx * ny * nz;
    if (n == 0) return WC_OK;
    if (!coef || !box_out) return WC_ERR_INVALID_ARG;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    const void* d_coef;
    rc = stage_in(ctx, ctx->ws_coef, coef, sizeof(float) * n, space, &d_coef);
    if (rc != WC_OK) return rc;
    CTX_CUDA(ctx, ctx->ws_boxes.reserve(sizeof(float) * n));
    CTX_CUDA(ctx, ctx->ws_tbl0.reserve(sizeof(InvUnitDev)));
    InvUnitDev iu;
    iu.coef = static_cast<const float*>(d_coef); iu.out = ctx->ws_boxes.p;
    iu.nx = nx; iu.ny = ny; iu.nz = nz; iu.dtype = WC_F32;
    int nt = xtile_count(nx, ny, nz);
    std::vector<int2> tiles(nt);
    for (int t = 0; t < nt; ++t) tiles[t] = make_int2(0, t);
    CTX_CUDA(ctx, ctx->ws_tiles0.reserve(sizeof(int2) * nt));
    CTX_CUDA(ctx, cudaMemcpyAsync(ctx->ws_tbl0.p, &iu, sizeof(iu), cudaMemcpyHostToDevice, ctx->stream));
    CTX_CUDA(ctx, cudaMemcpyAsync(ctx->ws_tiles0.p, tiles.data(), sizeof(int2) * nt, cudaMemcpyHostToDevice, ctx->stream));
    CTX_CUDA(ctx, launch_inverse_generic(ctx->ws_tbl0.as<InvUnitDev>(), ctx->ws_tiles0.as<int2>(), nt,
                                         ctx->stream, &ctx->ls));
    return stage_out(ctx, box_out, ctx->ws_boxes.p, sizeof(float) * n, space);
}

int wc_threshold_pack(wc_ctx* ctx, const float* coef, int n, double keep, int space,
                      wc_pair* pairs_out, int32_t* npairs_out) {
    if (!ctx || n < 0 || !npairs_out || (space != WC_HOST && space != WC_DEVICE)) return WC_ERR_INVALID_ARG;
    *npairs_out = 0;
    if (n == 0) return WC_OK;
    if (!coef || !pairs_out) return WC_ERR_INVALID_ARG;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    const void* d_coef;
    int rc = stage_in(ctx, ctx->ws_coef, coef, sizeof(float) * (size_t)n, space, &d_coef);
    if (rc != WC_OK) return rc;
    int nct = ctile_count(n);
    CTX_CUDA(ctx, ctx->ws_pairs.reserve(sizeof(wc_pair) * (size_t)n));
    CTX_CUDA(ctx, ctx->ws_tbl0.reserve(sizeof(UnitDev)));
    CTX_CUDA(ctx, ctx->ws_tbl1.reserve(sizeof(UnitState)));
    CTX_CUDA(ctx, ctx->ws_tiles0.reserve(sizeof(int2) * nct));
    CTX_CUDA(ctx, ctx->ws_sum.reserve(sizeof(int) * 4 * (size_t)nct));
    UnitDev u;
    std::memset(&u, 0, sizeof(u));
    u.coef = const_cast<float*>(static_cast<const float*>(d_coef));
    u.out  = ctx->ws_pairs.as<wc_pair>();
    u.nx = n; u.ny = 1; u.nz = 1; u.n = n; u.ctile0 = 0; u.nctiles = nct;
    std::vector<int2> tiles(nct);
    for (int t = 0; t < nct; ++t) tiles[t] = make_int2(0, t);
    CTX_CUDA(ctx, cudaMemcpyAsync(ctx->ws_tbl0.p, &u, sizeof(u), cudaMemcpyHostToDevice, ctx->stream));
    CTX_CUDA(ctx, cudaMemcpyAsync(ctx->ws_tiles0.p, tiles.data(), sizeof(int2) * nct, cudaMemcpyHostToDevice, ctx->stream));
    CTX_CUDA(ctx, cudaMemsetAsync(ctx->ws_tbl1.p, 0, sizeof(UnitState), ctx->stream));
    CTX_CUDA(ctx, launch_argmax_flat(ctx->ws_tbl0.as<UnitDev>(), ctx->ws_tbl1.as<UnitState>(),
                                     ctx->ws_tiles0.as<int2>(), nct, ctx->stream, &ctx->ls));
    volatile double one = 1.0;
    CTX_CUDA(ctx, launch_finalize_thresh(ctx->ws_tbl1.as<UnitState>(), 1, one - keep, nullptr,
                                         ctx->stream, &ctx->ls));
    int* ti = ctx->ws_sum.as<int>();
    CTX_CUDA(ctx, launch_pack_generic(ctx->ws_tbl0.as<UnitDev>(), ctx->ws_tbl1.as<UnitState>(), 1,
                                      ctx->ws_tiles0.as<int2>(), nct, ti, ti + nct, ti + 2 * nct,
                                      ti + 3 * nct, ctx->stream, &ctx->ls));
    UnitState hs;
    CTX_CUDA(ctx, cudaMemcpyAsync(&hs, ctx->ws_tbl1.p, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
    CTX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *npairs_out = hs.npairs;
    return stage_out(ctx, pairs_out, ctx->ws_pairs.p, sizeof(wc_pair) * (size_t)hs.npairs, space);
}

int wc_rle_decode(wc_ctx* ctx, const wc_pair* pairs, int npairs, int total, int space,
                  float* coef_out) {
    if (!ctx || npairs < 0 || total < 0 || (space != WC_HOST && space != WC_DEVICE)) return WC_ERR_INVALID_ARG;
    if (total == 0) return WC_OK;
    if (!coef_out || (npairs > 0 && !pairs)) return WC_ERR_INVALID_ARG;
    CTX_CUDA(ctx, cudaSetDevice(ctx->device));
    const void* d_pairs = nullptr;
    int rc = stage_in(ctx, ctx->ws_pairs, pairs, sizeof(wc_pair) * (size_t)npairs, space, &d_pairs);
    if (rc != WC_OK) return rc;
    int npt = ptile_count(npairs);
    CTX_CUDA(ctx, ctx->ws_coef.reserve(sizeof(float) * (size_t)total));
    CTX_CUDA(ctx, ctx->ws_tbl0.reserve(sizeof(DecUnitDev)));
    CTX_CUDA(ctx, ctx->ws_tiles0.reserve(sizeof(int2) * std::max(npt, 1)));
    CTX_CUDA(ctx, ctx->ws_sum.reserve(sizeof(long long) * std::max(npt, 1)));
    CTX_CUDA(ctx, ctx->ws_tbl2.reserve(64));
    DecUnitDev du;
    du.pairs = static_cast<const wc_pair*>(d_pairs); du.npairs_dev = nullptr;
    du.coef = ctx->ws_coef.as<float>(); du.npairs = npairs; du.total = total; du.ptile0 = 0; du.nptiles = npt;
    std::vector<int2> tiles(std::max(npt, 1));
    for (int t = 0; t < npt; ++t) tiles[t] = make_int2(0, t);
    CTX_CUDA(ctx, cudaMemcpyAsync(ctx->ws_tbl0.p, &du, sizeof(du), cudaMemcpyHostToDevice, ctx->stream));
    if (npt) CTX_CUDA(ctx, cudaMemcpyAsync(ctx->ws_tiles0.p, tiles.data(), sizeof(int2) * npt, cudaMemcpyHostToDevice, ctx->stream));
    CTX_CUDA(ctx, cudaMemsetAsync(ctx->ws_tbl2.p, 0, 64, ctx->stream));
    CTX_CUDA(ctx, cudaMemsetAsync(ctx->ws_coef.p, 0, sizeof(float) * (size_t)total, ctx->stream));
    CTX_CUDA(ctx, launch_rle_decode_generic(ctx->ws_tbl0.as<DecUnitDev>(), 1, ctx->ws_tiles0.as<int2>(), npt,
                                            ctx->ws_sum.as<long long>(), ctx->ws_tbl2.as<int>(),
                                            ctx->stream, &ctx->ls));
    int h_err = 0;
    CTX_CUDA(ctx, cudaMemcpyAsync(&h_err, ctx->ws_tbl2.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    rc = stage_out(ctx, coef_out, ctx->ws_coef.p, sizeof(float) * (size_t)total, space);
    if (rc != WC_OK) return rc;
    return h_err ? WC_ERR_CORRUPT : WC_OK;
}

} // extern "C"
