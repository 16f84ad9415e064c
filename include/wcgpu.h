/* wcgpu.h — C ABI of libwcgpu.so: the B200 (sm_100a) implementation of the numeric core of
 * carsonmw3/wavelet-compression.
 *
 * The reference has no plugin / FFI layer; its seam for this path is the C++ free-function
 * boundary that src/modes.cpp calls:
 *
 *     compress(multiBox3D&, std::vector<int>, double keep, int,int,int, std::string)
 *                                   src/compressor.h:9-15   (called src/modes.cpp:100-103,236-239)
 *     decompress(std::string, int,int,int,int) -> Box3D
 *                                   src/decompressor.h:6-10 (called src/modes.cpp:151-166,250-265)
 *     calc_rmse_per_box(const multiBox3D&, const multiBox3D&, int) -> std::vector<double>
 *                                   src/calc-loss.h:6-8     (called src/modes.cpp:271-280)
 *
 * Everything below replaces the NUMERIC part of those three functions; the LZMA container, the
 * file names and the side files stay with the host (wavelet-compression_b200/host/wc_dropin.hpp
 * re-creates the three signatures on top of this ABI; INTEGRATION.md shows the re-link).
 *
 * Conventions
 *   - A *unit* is one (timestep, level, box, component): one x-fastest slab of nx*ny*nz values,
 *     memory index m = i + nx*(j + ny*k)                                    (src/grid.h:18).
 *   - Coefficients are ordered z-fastest, f = (i*ny + j)*nz + k           (src/compressor.cpp:178-181).
 *   - A packed unit is K (int32 zero-run, float32 value) pairs in f order = bytes [20, 20+8K) of
 *     the reference's serialized buffer                                     (src/compressor.cpp:55-80);
 *     wc_serialize_header() produces bytes [0,20).
 *   - Plain C types only; no exceptions, no exit(): every call returns a wc_status.  There is no
 *     CPU fallback: without a usable CUDA device the calls fail with WC_ERR_NO_DEVICE.
 *   - One wc_ctx per GPU.  A ctx (and its plans) is not thread-safe; distinct ctxs are independent.
 */
#ifndef WCGPU_H
#define WCGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WCGPU_VERSION 200 /* 0.2.0 */

#if defined(__GNUC__)
#define WC_API __attribute__((visibility("default")))
#else
#define WC_API
#endif

typedef enum {
    WC_OK              = 0,
    WC_ERR_INVALID_ARG = 1, /* null pointer, negative count, unknown enum value             */
    WC_ERR_BAD_DIMS    = 2, /* nx/ny/nz < 0 or nx*ny*nz >= 2^31 (the reference's int limit) */
    WC_ERR_NO_DEVICE   = 3, /* no CUDA device / not an sm_100 part / driver missing         */
    WC_ERR_CUDA        = 4, /* a CUDA runtime call failed; see wc_last_error()              */
    WC_ERR_OOM         = 5, /* device or pinned-host allocation failed                      */
    WC_ERR_CAPACITY    = 6, /* a caller-provided output buffer is too small                 */
    WC_ERR_CORRUPT     = 7, /* packed stream is inconsistent (negative run, K > ncoef, ...) */
    WC_ERR_STATE       = 8  /* call order violated (e.g. fetch before compress)             */
} wc_status;

typedef enum { WC_F32 = 0, WC_F64 = 1 } wc_dtype;
typedef enum { WC_HOST = 0, WC_DEVICE = 1 } wc_space;

/* How the threshold of src/compressor.cpp:212-216 is scoped. */
typedef enum {
    WC_THRESH_PER_UNIT = 0, /* the reference: one threshold per (box, component)                    */
    WC_THRESH_GLOBAL   = 1, /* EXTENSION: one threshold for the whole batch, same max-rule applied to
                               the concatenation of the units in batch order (BASELINE config 5)   */
    WC_THRESH_QUANTILE = 2, /* EXTENSION (no counterpart in the reference, whose threshold is max * (1 - keep)): keep the
                               Kt = n - floor(keep * n) coefficients of largest magnitude of every unit — threshold = the
                               magnitude of rank Kt (descending, NaNs last), mask |c| > threshold, ties at the threshold
                               dropped.  Radix select over the unit's coefficients; needs a plan created under
                               WC_OPT_PATH = 1 (WC_ERR_STATE otherwise).                                              */
    WC_THRESH_QUANTILE_GLOBAL = 3 /* the same over the concatenation of all units of the batch */
} wc_thresh_mode;

typedef struct wc_ctx  wc_ctx;
typedef struct wc_plan wc_plan;

/* Input box of one unit.  dtype WC_F64 = raw AMReX FAB payload; it is narrowed to float32 with
 * round-to-nearest-even on the device exactly as src/preprocess.cpp:78 does on the host. */
typedef struct {
    const void* data;
    int32_t     dtype; /* wc_dtype */
    int32_t     nx, ny, nz;
} wc_box_desc;

/* Output box of one unit (decompression).  dtype WC_F64 widens the float32 result, as
 * src/writeplotfile.cpp:103 does before handing it to AMReX. */
typedef struct {
    void*   data;
    int32_t dtype; /* wc_dtype */
    int32_t nx, ny, nz;
} wc_box_out;

/* One (zero-run, value) pair: exactly 8 bytes of the reference's serialized stream. */
typedef struct {
    int32_t run;
    float   val;
} wc_pair;

/* One packed unit = the reference's CompressedWavelet (src/box-structs.h:65-70) without need32
 * (never serialized, src/compressor.cpp:55-80). */
typedef struct {
    int32_t  shape[3]; /* nx, ny, nz            -> serialized bytes [0,12)  */
    int32_t  ncoef;    /* nx*ny*nz              -> bytes [12,16)            */
    int32_t  npairs;   /* K                     -> bytes [16,20)            */
    int32_t  flags;    /* WC_PACKED_* bits, set by the compress calls; not serialized */
    wc_pair* pairs;    /* K pairs               -> bytes [20,20+8K)         */
} wc_packed;
/* CompressedWavelet::need32 (src/compressor.cpp:224-229, src/box-structs.h:69): some kept value has
 * |v| > INT16_MAX, i.e. the TODO'd 16-bit value stream (TODO.txt:1-2) could not hold this unit.  Computed on
 * the device, never part of the byte stream (the reference does not serialize it either). */
#define WC_PACKED_NEED32 1

/* ---- library / context ---------------------------------------------------------------------- */
WC_API int         wc_version(void);
WC_API const char* wc_strerror(int status);
WC_API int         wc_device_count(int* count);
/* Informational (no GPU needed, no reference counterpart): the name of the kernel class a box of this shape gets when its
 * pointers are 16-byte aligned — `decompress` = 0: the compress side, 1: the decompress side.  "cube8" / "cube16" /
 * "cube32" / "cube64" and "r1s" / "r1" / "r2" / "r4" / "r8" are the fused y-slab kernels (one CTA, or a cluster of 2 / 4 / 8),
 * "yslab" the two-pass y-slab kernels of boxes no cluster holds (128^3 ...), "xs1s" / "xs1" / "xs2" / "xs4" / "xs8" the x-slab
 * kernels that take any shape (odd dimensions, nz % 4 != 0 ...), "generic" the multi-kernel path, "empty" a box without
 * cells, "invalid" a negative dimension or an unknown dtype.  Results never depend on the class. */
WC_API const char* wc_box_kernel_class(int nx, int ny, int nz, int dtype, int decompress);

/* Creates a context on CUDA device `device_id` with its own non-blocking stream. */
WC_API int wc_create(wc_ctx** ctx, int device_id);
/* Same, but all work is issued on the caller's stream (a cudaStream_t, e.g. torch's current
 * stream), so the caller can bracket it with its own CUDA events. */
WC_API int wc_create_on_stream(wc_ctx** ctx, int device_id, void* cuda_stream);
WC_API int wc_destroy(wc_ctx* ctx);
WC_API int wc_sync(wc_ctx* ctx);
/* Text of the last failing CUDA call on this ctx ("" if none). */
WC_API const char* wc_last_error(const wc_ctx* ctx);

typedef enum {
    WC_OPT_PATH = 0,   /* 0 = auto (fused on-chip kernels when a unit fits, generic otherwise),
                          1 = force the generic multi-kernel path, 2 = force fused (error if a unit
                          does not fit) */
    WC_OPT_PROFILE = 1, /* 1 = bracket every kernel launch with CUDA events on the ctx stream and
                           accumulate per-kernel device time (read with wc_kernel_stats) */
    WC_OPT_OVERLAP = 2, /* 1 (default) = run the single-CTA and the cluster compress kernels of a step
                           concurrently (second stream, dynamic unit hand-out); 0 = back to back */
    WC_OPT_SEG_INDEX = 3, /* segment index of table-less packed streams: 0 (default) = chunk-parallel
                             single-pass scan (k_seg_index2), 1 = one CTA per unit (k_seg_index) */
    WC_OPT_COPY_ONLY = 4, /* measurement probe: 1 = wc_plan_compress_to_host issues its H2D / D2H copies with
                             the same chunking but launches no compress kernel (the pair counts of the previous
                             real call size the D2H) -> the host-link ceiling of that call */
    WC_OPT_INGEST_STATS = 5, /* 1 = the compress kernels also record each unit's min / max of the narrowed
                                input values (src/preprocess.cpp:82-88), read with wc_plan_unit_stats */
    WC_OPT_DECODE_PIPE = 6   /* decompress kernel of the 32^3 cubes: 0 = pairs are read straight from global memory (by
                                segment table when one came with the unit, else with a block-wide scan); 1 = units
                                without a table decode from a shared-memory staging area that TMA bulk copies fill one
                                unit ahead; 2 (default) = every unit does (the staged decode measured faster than the
                                table-driven one: 0.57 vs 0.64 ms for the 12288 units of the bench) */
} wc_option;
WC_API int wc_set_option(wc_ctx* ctx, int option, int64_t value);

typedef enum {
    WC_CTR_KERNEL_LAUNCHES = 0, /* kernels this ctx launched since creation / last reset */
    WC_CTR_H2D_BYTES       = 1,
    WC_CTR_D2H_BYTES       = 2
} wc_counter;
WC_API int wc_get_counter(const wc_ctx* ctx, int counter, uint64_t* value);
WC_API int wc_reset_counters(wc_ctx* ctx);
/* Per-kernel launch count and (with WC_OPT_PROFILE) accumulated device time since the last reset.
 * index runs from 0 until the call returns WC_ERR_INVALID_ARG.  Synchronizes the ctx stream. */
WC_API int wc_kernel_stats(wc_ctx* ctx, int index, const char** name, double* total_ms,
                           uint64_t* launches);

/* Pinned host memory (so host-resident boxes / packed streams move at PCIe speed). */
WC_API int wc_host_alloc(void** ptr, size_t bytes);
WC_API int wc_host_free(void* ptr);
/* Device memory + copies for hosts that do not otherwise link CUDA. kind: 0 H2D, 1 D2H, 2 D2D. */
WC_API int wc_device_alloc(wc_ctx* ctx, void** ptr, size_t bytes);
WC_API int wc_device_free(wc_ctx* ctx, void* ptr);
WC_API int wc_memcpy(wc_ctx* ctx, void* dst, const void* src, size_t bytes, int kind);

/* ---- blocking batch API: the numeric part of the three reference functions ------------------- */

/* compress(): F -> T -> M -> P of src/compressor.cpp:203-247 for n_units units.
 *   in / in_space : the boxes, all in host memory or all in device memory.
 *   keep          : as the reference's `double keep` (the CLI widens a float, src/argparse.h:13).
 *   out[u]        : shape/ncoef/npairs are filled in; out[u].pairs is set to ctx-owned memory in
 *                   out_space, valid until the next call on this ctx.  WC_HOST: pinned memory, the
 *                   units' pairs back to back in unit order.  WC_DEVICE: each unit's pairs start at
 *                   its own slot (capacity ncoef pairs). */
WC_API int wc_compress_batch(wc_ctx* ctx, const wc_box_desc* in, int n_units, int in_space, double keep,
                      int thresh_mode, wc_packed* out, int out_space);

/* decompress(): U -> I of src/decompressor.cpp:245-254 for n_units packed units.  out[u].data is
 * caller-owned memory in out_space with room for shape[0]*shape[1]*shape[2] values of out[u].dtype;
 * out[u].nx/ny/nz must equal in[u].shape. */
WC_API int wc_decompress_batch(wc_ctx* ctx, const wc_packed* in, int n_units, int in_space,
                        const wc_box_out* out, int out_space);

/* calc_rmse_per_box(): src/calc-loss.cpp:12-43 for n_units (actual, pred) pairs of float32 boxes
 * with equal dims.  rmse is a HOST array of n_units doubles. */
WC_API int wc_rmse_batch(wc_ctx* ctx, const wc_box_desc* actual, const wc_box_desc* pred, int n_units,
                  int space, double* rmse);

/* Ingest statistics: per-unit minimum and maximum of the (narrowed) float32 values, the quantities
 * src/preprocess.cpp:82-88 accumulates per component for the adjusted loss (src/modes.cpp:289).  A unit
 * without any comparable value (empty, all NaN) reports +inf / -inf.  mins / maxs are HOST arrays. */
WC_API int wc_minmax_batch(wc_ctx* ctx, const wc_box_desc* boxes, int n_units, int space, float* mins,
                           float* maxs);

/* ---- un-fused primitives (one unit, blocking) for parity tests -------------------------------- */
/* wavelet_decompose, src/compressor.cpp:85-185: coef_out gets nx*ny*nz float32 in f order. */
WC_API int wc_haar_forward(wc_ctx* ctx, const wc_box_desc* in, int space, float* coef_out);
/* inverse_wavelet_decompose, src/decompressor.cpp:79-159. */
WC_API int wc_haar_inverse(wc_ctx* ctx, const float* coef, int nx, int ny, int nz, int space,
                    float* box_out);
/* threshold + mask + rle_encode, src/compressor.cpp:212-237 + :24-42.  pairs_out has room for n
 * pairs; *npairs_out (host) receives K. */
WC_API int wc_threshold_pack(wc_ctx* ctx, const float* coef, int n, double keep, int space,
                      wc_pair* pairs_out, int32_t* npairs_out);
/* rle_decode, src/decompressor.cpp:14-30: coef_out gets `total` float32. */
WC_API int wc_rle_decode(wc_ctx* ctx, const wc_pair* pairs, int npairs, int total, int space,
                  float* coef_out);
/* bytes [0,20) of serialize_compressed_wavelet, src/compressor.cpp:59-71 (host-side helper). */
WC_API int wc_serialize_header(const wc_packed* unit, uint8_t header_out[20]);

/* ---- plan API: device-resident, asynchronous, allocation-free after creation ------------------ *
 * A plan fixes the unit list (dims, dtype, input addresses) of a batch — e.g. all boxes and
 * components of one timestep — builds the device descriptor tables once, and owns every scratch
 * and output buffer, so that running it is a fixed sequence of kernel launches on the ctx stream. */
WC_API int wc_plan_create(wc_ctx* ctx, const wc_box_desc* units, int n_units, int in_space,
                   wc_plan** plan);
WC_API int wc_plan_destroy(wc_plan* plan);
/* New input addresses for the same dims/dtypes (e.g. the next timestep on the same grids).  Device inputs:
 * stream-ordered, no host synchronisation (pinned pointer table + a patch kernel), so a timestep series
 * alternates wc_plan_set_inputs / wc_plan_compress without draining the GPU. */
WC_API int wc_plan_set_inputs(wc_plan* plan, const wc_box_desc* units);
/* Enqueue compression of every unit.  Returns after enqueueing when the inputs are on the device;
 * host inputs are first staged with (asynchronous, if pinned) H2D copies on the same stream. */
WC_API int wc_plan_compress(wc_plan* plan, double keep, int thresh_mode);
/* Wait for the last compress and describe its result.  WC_DEVICE: out[u].pairs = the unit's slot in
 * plan-owned device memory.  WC_HOST: the kept pairs of all units are gathered densely on the device,
 * copied D2H once into plan-owned pinned memory, and out[u].pairs points into it. */
WC_API int wc_plan_fetch(wc_plan* plan, wc_packed* out, int out_space);
/* compress + fetch(WC_HOST) for host-resident inputs in ONE pipelined call: the unit list is cut into
 * chunks and the H2D copy of chunk c overlaps the kernels of chunk c-1 and the gather + D2H of chunk
 * c-2 (three streams; PCIe is full duplex).  Per-unit thresholds only.  Falls back to
 * wc_plan_compress + wc_plan_fetch when the plan's inputs are device-resident. */
WC_API int wc_plan_compress_to_host(wc_plan* plan, double keep, wc_packed* out);
/* The same, handing finished chunks to the host while later chunks are still on the GPU, so that the host's
 * LZMA stage (src/compressor.cpp:256-291) overlaps the GPU work: on_chunk(user, first_unit, n_units, units) is
 * called on the calling thread, in unit order, as soon as the pairs of units [first_unit, first_unit + n_units)
 * have landed in pinned host memory; `units` = &out[first_unit], filled in.  The pointers stay valid until the
 * next compress call on this plan.  on_chunk may be NULL (= wc_plan_compress_to_host). */
typedef void (*wc_chunk_fn)(void* user, int first_unit, int n_units, const wc_packed* units);
WC_API int wc_plan_compress_to_host_chunked(wc_plan* plan, double keep, wc_packed* out, wc_chunk_fn on_chunk,
                                            void* user);
/* Per-unit by-products of the last compress (HOST arrays of n_units entries, any of them may be NULL):
 * mins / maxs = min / max of the unit's narrowed float32 input values, NaNs skipped, +inf / -inf when no value
 * is comparable (what src/preprocess.cpp:82-88 folds into the per-component range of the adjusted loss;
 * needs WC_OPT_INGEST_STATS = 1 at compress time, else WC_ERR_STATE); need32 = CompressedWavelet::need32. */
WC_API int wc_plan_unit_stats(wc_plan* plan, float* mins, float* maxs, int32_t* need32);
/* Total kept pairs of the last compress (waits for it). */
WC_API int wc_plan_total_pairs(wc_plan* plan, int64_t* total);
/* Enqueue decompression of the plan's current packed result into out[u] (device memory, or host
 * memory: D2H copies are enqueued after the kernels) — the on-device round trip that estimate mode
 * needs (src/modes.cpp:236-265). */
WC_API int wc_plan_decompress(wc_plan* plan, const wc_box_out* out, int out_space);
/* Enqueue + wait: RMSE of recon[u] (float32, device) against the plan's own input boxes. */
WC_API int wc_plan_rmse(wc_plan* plan, const wc_box_desc* recon, double* rmse);
/* EXTENSION for multi-GPU WC_THRESH_GLOBAL: split compress at the threshold so the caller can
 * all-reduce the arg-max key between the two halves (NCCL ncclMax on one uint64 per plan).
 *   wc_plan_transform       : forward transform + local arg-max key -> *key_dev (device uint64*)
 *   wc_plan_pack_with_key   : threshold from *key_dev (device) + mask + pack                     */
WC_API int wc_plan_transform(wc_plan* plan, uint64_t** key_dev);
WC_API int wc_plan_pack_with_key(wc_plan* plan, double keep, const uint64_t* key_dev);
/* EXTENSION for multi-GPU WC_THRESH_QUANTILE_GLOBAL: the radix select split at its histograms, so that the caller can
 * all-reduce them (NCCL ncclSum over 2048 uint64) and every rank picks the same bucket — the kept set is then the same
 * as on one GPU.
 *   wc_plan_quantile_begin : forward transform into the scratch; n_total = coefficients of ALL ranks (0: this plan's)
 *   per pass 0, 1, 2       : wc_plan_quantile_hist (-> *hist_dev, device uint64[2048], add the other ranks' in place)
 *                            then wc_plan_quantile_pick
 *   wc_plan_quantile_pack  : thresholds -> mask + ordered packing
 * (global = 0 runs the per-unit mode through the same calls; its histograms are one row per unit.) */
WC_API int wc_plan_quantile_begin(wc_plan* plan, double keep, int global, uint64_t n_total);
WC_API int wc_plan_quantile_hist(wc_plan* plan, int pass, uint64_t** hist_dev);
WC_API int wc_plan_quantile_pick(wc_plan* plan, int pass);
WC_API int wc_plan_quantile_pack(wc_plan* plan);

/* ---- decode plans: decompress() (src/decompressor.cpp:238-255) for a batch, stream-in ------------------ *
 * A decode plan fixes the output boxes (dims, dtype, addresses) of a batch and owns the device tables; each
 * wc_dplan_decode then takes the batch's packed units as ONE dense pair stream — the units' pairs back to back
 * in unit order, i.e. the concatenation of bytes [20, 20+8K) of the reference's files, which is also what
 * wc_plan_fetch(WC_HOST) / wc_plan_compress_to_host return — plus the K of every unit.  Nothing but the two
 * arrays is needed from the host: per-unit offsets, the segment index of the large units and the inverse
 * transform all run on the device.  Host streams are pipelined (H2D of chunk c | kernels of c-1 | D2H of the
 * boxes of c-2). */
typedef struct wc_dplan wc_dplan;
WC_API int wc_dplan_create(wc_ctx* ctx, const wc_box_out* outs, int n_units, int out_space, wc_dplan** dplan);
WC_API int wc_dplan_destroy(wc_dplan* dplan);
/* Enqueue: pairs (sum of npairs entries) and npairs (n_units entries) both in `in_space`.  Asynchronous for
 * device streams; results are complete after wc_dplan_finish. */
WC_API int wc_dplan_decode(wc_dplan* dplan, const wc_pair* pairs, const int32_t* npairs, int in_space);
/* Wait for the last decode; WC_ERR_CORRUPT if a unit's stream was inconsistent (negative run, K > ncoef). */
WC_API int wc_dplan_finish(wc_dplan* dplan);

#ifdef __cplusplus
}
#endif

#endif /* WCGPU_H */
