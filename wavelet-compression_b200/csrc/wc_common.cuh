// Shared device-side definitions for libwcgpu (sm_100a).
//
// Arithmetic contract (SURVEY.md §0 "fp32 equivalence"): the reference's forward pass
// `(a + b) / 2.0` (float add, double divide, float store; src/compressor.cpp:108-110) and inverse
// pass `avg + diff` in double stored to float (src/decompressor.cpp:99-108) are bit-identical to
// single fp32 operations add/sub then multiply-by-0.5 (forward) and add/sub (inverse), provided
// nothing is contracted into an FMA, nothing is re-associated and denormals are kept.  Hence the
// explicit __f*_rn intrinsics below and no --use_fast_math / -ftz in the build.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wcgpu.h"

namespace wc {

typedef unsigned long long u64;

// ---- per-unit device records -------------------------------------------------------------------
struct UnitDev {
    const void* in;    // box, x-fastest, dtype below (device)
    wc_pair*    out;   // pair slot, capacity >= n pairs (device)
    float*      coef;  // coefficient scratch in f order (generic path; may be null for fused units)
    int32_t     nx, ny, nz;
    int32_t     n;     // nx*ny*nz
    int32_t     dtype; // wc_dtype of `in`
    int32_t     ctile0; // first flat (compaction / rmse) tile of this unit
    int32_t     nctiles;
    int32_t     reserved;
};

struct UnitState {
    u64     key;      // arg-max key, see make_key
    float   thresh_f; // largest float <= thresh (see threshold_float)
    int32_t npairs;   // K
    int32_t flags;    // bit0: coefficient f=0 is NaN; bit1: need32 (a kept |value| > INT16_MAX,
                      // src/compressor.cpp:224-229)
    float   vmin;     // with WC_OPT_INGEST_STATS: min / max of the narrowed input values, NaNs skipped
    float   vmax;     //   (src/preprocess.cpp:82-88); +inf / -inf when nothing is comparable
    int32_t reserved;
};
static_assert(sizeof(UnitState) == 32, "UnitState layout");
constexpr int UNIT_FLAG_NAN0 = 1, UNIT_FLAG_NEED32 = 2;

// need32 of src/compressor.cpp:224-229: some KEPT coefficient has |value| > INT16_MAX.  The coefficient of
// largest magnitude M is kept whenever anything above INT16_MAX is (|c| > tf is monotone in |c|), so the
// flag is "M > 32767 and M passes the mask".
__device__ __forceinline__ bool unit_need32(float M, float tf) { return M > 32767.0f && M > tf; }

// ---- arg-max key -----------------------------------------------------------------------------------
// std::max_element with comp(a,b) = |a| < |b| (src/compressor.cpp:212-215) returns the FIRST element
// of maximal magnitude and skips NaN candidates.  Encoded as an unsigned 64-bit maximum:
//   [63:32] bits of |c|   [31:1] 0x7fffffff - f (smaller f wins ties)   [0] sign of c
// NaN candidates map to 0 (= "no candidate"); every real candidate is >= 2 because f <= 2^31-2.
// A NaN at f = 0 (never replaced in the sequential rule) is tracked separately (UnitState.flags).
__device__ __forceinline__ u64 make_key(float c, uint32_t f) {
    uint32_t b = __float_as_uint(c);
    uint32_t a = b & 0x7fffffffu;
    if (a > 0x7f800000u) return 0ull;
    return ((u64)a << 32) | ((u64)(0x7fffffffu - f) << 1) | (u64)(b >> 31);
}
__device__ __forceinline__ float key_value(u64 key) {
    uint32_t a = (uint32_t)(key >> 32);
    uint32_t s = (uint32_t)(key & 1ull);
    return __uint_as_float(a | (s << 31));
}
__device__ __forceinline__ u64 max_u64(u64 a, u64 b) { return a > b ? a : b; }

__device__ __forceinline__ u64 warp_max_u64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max_u64(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// thresh = max_val * (1 - keep) in double (src/compressor.cpp:216); the mask test
// |(double)c| > thresh (:226) is equivalent, for every float c, to |c| > tf with tf the largest
// float <= thresh (round toward -inf): thresh NaN -> tf NaN (nothing kept), thresh < 0 -> tf < 0
// (every non-NaN kept), thresh = +inf -> tf = +inf (nothing kept).
__device__ __forceinline__ float threshold_float(u64 key, bool first_is_nan, double one_minus_keep) {
    if (first_is_nan) return __uint_as_float(0x7fc00000u);
    double max_val = (double)key_value(key);
    double thresh  = __dmul_rn(max_val, one_minus_keep);
    return __double2float_rd(thresh);
}

__device__ __forceinline__ bool keep_coef(float c, float tf) { return fabsf(c) > tf; }

// ---- the 2x2x2 Haar block ---------------------------------------------------------------------------
// v[zi*4 + yi*2 + xi] holds the (narrowed) inputs of one block; wx/wy/wz say whether the block is a
// full pair along that axis (false = the trailing singleton of an odd dimension, which passes through).
// On return v[sz*4 + sy*2 + sx] is the coefficient of sub-band (sx,sy,sz).  Pass order Z, Y, X as the
// reference (src/compressor.cpp:98-175).
__device__ __forceinline__ void haar_pair(float& lo, float& hi) {
    float s = __fadd_rn(lo, hi);
    float d = __fsub_rn(lo, hi);
    lo      = __fmul_rn(s, 0.5f);
    hi      = __fmul_rn(d, 0.5f);
}
__device__ __forceinline__ void haar_block_forward(float v[8], bool wx, bool wy, bool wz) {
    if (wz) {
#pragma unroll
        for (int q = 0; q < 4; ++q) haar_pair(v[q], v[4 + q]);
    }
    if (wy) {
#pragma unroll
        for (int zi = 0; zi < 2; ++zi)
#pragma unroll
            for (int xi = 0; xi < 2; ++xi) haar_pair(v[zi * 4 + xi], v[zi * 4 + 2 + xi]);
    }
    if (wx) {
#pragma unroll
        for (int q = 0; q < 4; ++q) haar_pair(v[2 * q], v[2 * q + 1]);
    }
}
// fast version for full blocks
__device__ __forceinline__ void haar_block_forward_full(float v[8]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) haar_pair(v[q], v[4 + q]);
#pragma unroll
    for (int zi = 0; zi < 2; ++zi)
#pragma unroll
        for (int xi = 0; xi < 2; ++xi) haar_pair(v[zi * 4 + xi], v[zi * 4 + 2 + xi]);
#pragma unroll
    for (int q = 0; q < 4; ++q) haar_pair(v[2 * q], v[2 * q + 1]);
}

// Inverse of a full block, pass order X, Y, Z (src/decompressor.cpp:90-156): v[sz*4+sy*2+sx] in,
// v[zi*4+yi*2+xi] out.
__device__ __forceinline__ void ihaar_pair(float& avg, float& diff) {
    float p = __fadd_rn(avg, diff);
    float m = __fsub_rn(avg, diff);
    avg     = p;
    diff    = m;
}
__device__ __forceinline__ void haar_block_inverse_full(float v[8]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) ihaar_pair(v[2 * q], v[2 * q + 1]);
#pragma unroll
    for (int zi = 0; zi < 2; ++zi)
#pragma unroll
        for (int xi = 0; xi < 2; ++xi) ihaar_pair(v[zi * 4 + xi], v[zi * 4 + 2 + xi]);
#pragma unroll
    for (int q = 0; q < 4; ++q) ihaar_pair(v[q], v[4 + q]);
}

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

} // namespace wc
