// Host-visible declarations of the kernel launchers (wc_generic.cu, wc_fused.cu) used by wc_api.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/wcgpu.h"

namespace wc {

typedef unsigned long long u64;
struct UnitDev;
struct UnitState;

// ---- launch accounting / optional per-kernel CUDA-event timing (WC_OPT_PROFILE) -----------------
enum KernelId {
    KID_FORWARD_GENERIC, KID_ARGMAX_FLAT, KID_FINALIZE, KID_GLOBAL_KEY, KID_COUNT, KID_SCAN, KID_EMIT,
    KID_RLE_SUMS, KID_RLE_SCAN, KID_RLE_SCATTER, KID_INVERSE_GENERIC, KID_RMSE_TILES, KID_RMSE_FINAL,
    KID_OFFSETS, KID_GATHER, KID_FUSED_C1, KID_FUSED_C8, KID_FUSED_C1S, KID_FUSED_C8S, KID_FUSED_D1, KID_FUSED_D8, KID_FUSED_D1S, KID_FUSED_D8S, KID_SEG_INDEX, KID_FUSED_C1T, KID_FUSED_C16, KID_FUSED_D1T, KID_FUSED_D16, KID_FUSED_C8C, KID_FUSED_D8C, KID_FUSED_C4, KID_FUSED_C2, KID_FUSED_D4, KID_FUSED_D2, KID_MINMAX_TILES, KID_MINMAX_FINAL, KID_SEG_INDEX2, KID_DEC_PREPARE, KID_PATCH_INPUTS, KID_STAGED_D1S, KID_FUSED_DBIG, KID_BIG_FORWARD, KID_BIG_PACK, KID_Q_HIST, KID_Q_PICK, KID_FUSED_D128, KID_XS_C1, KID_XS_C2, KID_XS_C4, KID_XS_C8, KID_XS_D, KID_XS_C1S, KID_XS_DS, KID_N
};
inline const char* kernel_name(int id) {
    static const char* n[KID_N] = {
        "k_forward_generic", "k_argmax_flat", "k_finalize_thresh", "k_global_key", "k_count_tiles",
        "k_scan_tiles", "k_emit_tiles", "k_rle_tile_sums", "k_rle_scan", "k_rle_scatter",
        "k_inverse_generic", "k_rmse_tiles", "k_rmse_final", "k_unit_offsets", "k_gather_dense",
        "k_fused_compress<1>", "k_fused_compress<8>", "k_fused_compress<1,cube32>", "k_fused_compress<8,cube64>",
        "k_fused_decompress<1>",
        "k_fused_decompress<8>", "k_fused_decompress<1,cube32>", "k_fused_decompress<8,cube64>", "k_seg_index", "k_fused_compress<1,small>", "k_fused_compress<1,cube16>",
        "k_fused_decompress<1,small>", "k_fused_decompress<1,cube16>", "k_fused_compress<1,cube8>",
        "k_fused_decompress<1,cube8>", "k_fused_compress<4>", "k_fused_compress<2>", "k_fused_decompress<4>",
        "k_fused_decompress<2>", "k_minmax_tiles", "k_minmax_final", "k_seg_index3", "k_dec_prepare",
        "k_patch_inputs", "k_staged_decompress<1,cube32>", "k_fused_decompress<slabs>", "k_big_forward", "k_big_pack", "k_q_hist", "k_q_pick", "k_fused_decompress<64,cube128>",
        "k_xs_compress<1>", "k_xs_compress<2>", "k_xs_compress<4>", "k_xs_compress<8>", "k_xs_decompress",
        "k_xs_compress<1,small>", "k_xs_decompress<small>" };
    return (id >= 0 && id < KID_N) ? n[id] : "?";
}
struct LaunchStats {
    uint64_t launches = 0;
    bool     profile  = false;
    uint64_t count[KID_N] = {};
    double   ms[KID_N]    = {};
    int      occ[KID_N]   = {};   // occupancy of each kernel on this ctx's device (0 = not queried yet)
    int      occ_mm[KID_N] = {};  // same for the ingest-statistics instantiations of the compress kernels
    struct Pending { int id; cudaEvent_t a, b; };
    std::vector<Pending>     pending;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get_event() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    void begin(int id, cudaStream_t st) {
        ++launches;
        ++count[id];
        if (!profile) return;
        Pending p { id, get_event(), get_event() };
        cudaEventRecord(p.a, st);
        pending.push_back(p);
    }
    void end(cudaStream_t st) {
        if (!profile || pending.empty()) return;
        cudaEventRecord(pending.back().b, st);
    }
    // call after the stream has been synchronised
    void collect() {
        for (Pending& p : pending) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) ms[p.id] += t;
            pool.push_back(p.a);
            pool.push_back(p.b);
        }
        pending.clear();
        cudaGetLastError();
    }
    void reset() {
        launches = 0;
        for (int i = 0; i < KID_N; ++i) { count[i] = 0; ms[i] = 0.0; }
    }
    void destroy() {
        collect();
        for (cudaEvent_t e : pool) cudaEventDestroy(e);
        pool.clear();
    }
};

// ---- generic path geometry ---------------------------------------------------------------------
constexpr int XT_THREADS    = 256;
constexpr int XT_BLOCKS     = 1024;             // 2x2x2 blocks per transform tile
constexpr int XT_OCT_FLOATS = 1536;             // max rows*stride per sub-band (TA=2: 512 rows * 3)
constexpr int XT_SMEM_BYTES = 8 * XT_OCT_FLOATS * 4;
constexpr int CT_THREADS    = 256;
constexpr int CT_ELEMS      = 2048;             // coefficients per flat tile
constexpr int PT_THREADS    = 256;
constexpr int PT_PAIRS      = 2048;             // pairs per rle-decode tile

int xtile_count(int nx, int ny, int nz);
inline int ctile_count(long long n) { return (int)((n + CT_ELEMS - 1) / CT_ELEMS); }
inline int ptile_count(long long k) { return (int)((k + PT_PAIRS - 1) / PT_PAIRS); }

struct DecUnitDev {
    const wc_pair* pairs;
    const int32_t* npairs_dev; // when set, K is read from the device (plan round trip) and
                               // `npairs` is only the bound the tile table was sized for
    float*         coef;       // zero-filled scratch, `total` floats (generic path); for the slab-decoded
                               // fused classes (R8 / CUBE64): the unit's segment table, int2[2*nx*8 + 1]
    int32_t        npairs;
    int32_t        total;      // ncoef
    int32_t        ptile0;
    int32_t        nptiles;
};

struct InvUnitDev {
    const float* coef;
    void*        out;
    int32_t      nx, ny, nz;
    int32_t      dtype;
};

struct RmseUnitDev {
    const void* a; // actual:  float32, or the raw float64 FAB slab (narrowed on load, A1)
    const void* b; // pred:    float32, or float64 holding widened float32 values
    int32_t     a_dtype, b_dtype;
    int32_t     n;
    int32_t     ctile0;
    int32_t     nctiles;
    int32_t     reserved;
};

// ---- generic launchers (all asynchronous on `st`; every kernel is accounted in *ls) -------------
cudaError_t launch_forward_generic(const UnitDev* units, UnitState* states, const int2* tiles,
                                   int n_tiles, cudaStream_t st, LaunchStats* ls);
cudaError_t launch_argmax_flat(const UnitDev* units, UnitState* states, const int2* ctiles,
                               int n_ctiles, cudaStream_t st, LaunchStats* ls);
cudaError_t launch_finalize_thresh(UnitState* states, int n_units, double one_minus_keep,
                                   const u64* global_key, cudaStream_t st, LaunchStats* ls);
// quantile thresholds (extension): radix select over the coefficient scratch
constexpr int Q_BINS = 2048;
struct QState { unsigned long long rank, nvalid; uint32_t prefix; int32_t done; float thresh; int32_t pad; };
cudaError_t launch_q_init(QState* q, const unsigned long long* rank, int nq, cudaStream_t st);
cudaError_t launch_q_hist(const UnitDev* units, const int2* ctiles, int n_ctiles, const QState* q, unsigned long long* hist,
                          int pass, bool global, cudaStream_t st, LaunchStats* ls);
cudaError_t launch_q_pick(QState* q, unsigned long long* hist, int nq, int pass, cudaStream_t st, LaunchStats* ls);
cudaError_t launch_q_apply(UnitState* states, const QState* q, int n_units, bool global, cudaStream_t st);
cudaError_t launch_global_key(const UnitDev* units, const UnitState* states, int n_units, u64* out,
                              cudaStream_t st, LaunchStats* ls);
cudaError_t launch_patch_inputs(UnitDev* units, const void* const* ptrs, int n_units, cudaStream_t st,
                                LaunchStats* ls);
cudaError_t launch_pack_generic(const UnitDev* units, UnitState* states, int n_units,
                                const int2* ctiles, int n_ctiles, int* tile_cnt, int* tile_last,
                                int* tile_base, int* tile_prev, cudaStream_t st,
                                LaunchStats* ls);
cudaError_t launch_rle_decode_generic(const DecUnitDev* units, int n_units, const int2* ptiles,
                                      int n_ptiles, long long* tile_sum, int* err, cudaStream_t st,
                                      LaunchStats* ls);
cudaError_t launch_inverse_generic(const InvUnitDev* units, const int2* tiles, int n_tiles,
                                   cudaStream_t st, LaunchStats* ls);
cudaError_t launch_rmse_generic(const RmseUnitDev* units, int n_units, const int2* ctiles,
                                int n_ctiles, double* tile_sum, double* rmse, cudaStream_t st,
                                LaunchStats* ls);
cudaError_t launch_minmax_generic(const RmseUnitDev* units, int n_units, const int2* ctiles, int n_ctiles,
                                  float2* tile_mm, float2* out, cudaStream_t st, LaunchStats* ls);
cudaError_t launch_gather_dense(const UnitDev* units, const UnitState* states, int n_units,
                                long long* offsets, wc_pair* dense, bool offsets_only,
                                cudaStream_t st, LaunchStats* ls, long long* running = nullptr);

} // namespace wc
