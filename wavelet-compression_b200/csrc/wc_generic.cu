// Generic (any box size, any dims incl. odd) multi-kernel path of libwcgpu.
//
// Coefficients live in an HBM/L2 scratch array in f order between the transform and the packing
// kernels.  This is the fallback for units that do not fit the fused on-chip kernels
// (wc_fused.cu) and the implementation behind the un-fused parity primitives of the C ABI.
//
//   forward : k_forward_generic  (F + arg-max key of T)     src/compressor.cpp:85-185, :212-215
//             k_finalize_thresh  (T)                        src/compressor.cpp:216
//             k_count_tiles / k_scan_tiles / k_emit_tiles   src/compressor.cpp:222-237, :24-42 (M, P)
//   inverse : k_rle_tile_sums / k_rle_scan / k_rle_scatter  src/decompressor.cpp:14-30 (U)
//             k_inverse_generic                              src/decompressor.cpp:79-159 (I)
//   loss    : k_rmse_tiles / k_rmse_final                    src/calc-loss.cpp:12-43 (R)
#include "wc_common.cuh"
#include "wc_kernels.h"

namespace wc {

// ============================================================================================
// transform tiles
// ============================================================================================
// A tile is a TA x TB x TC brick of 2x2x2 blocks (<= XT_BLOCKS blocks) of one unit; blocks are
// enumerated with a fastest, so global loads run along x, and the 8 sub-band values of every block
// are staged in shared memory so the scratch writes run along k' (the f order's fastest axis).

__host__ __device__ inline void xtile_shape(int nbx, int nby, int nbz, int& TA, int& TB, int& TC) {
    TA = nbx < XT_BLOCKS ? nbx : XT_BLOCKS;
    if (TA < 1) TA = 1;
    int r = XT_BLOCKS / TA;
    if (r < 1) r = 1;
    TC = nbz < r ? nbz : r;
    if (TC < 1) TC = 1;
    r = XT_BLOCKS / (TA * TC);
    if (r < 1) r = 1;
    TB = nby < r ? nby : r;
    if (TB < 1) TB = 1;
}

int xtile_count(int nx, int ny, int nz) {
    if (nx <= 0 || ny <= 0 || nz <= 0) return 0;
    int nbx = (nx + 1) / 2, nby = (ny + 1) / 2, nbz = (nz + 1) / 2;
    int TA, TB, TC;
    xtile_shape(nbx, nby, nbz, TA, TB, TC);
    return ((nbx + TA - 1) / TA) * ((nby + TB - 1) / TB) * ((nbz + TC - 1) / TC);
}

struct XTileGeom {
    int nbx, nby, nbz, hx, hy, hz;
    int TA, TB, TC, S, rows;
    int a0, b0, c0;
};

__device__ __forceinline__ XTileGeom xtile_geom(const UnitDev& u, int local) {
    XTileGeom g;
    g.hx  = u.nx / 2;
    g.hy  = u.ny / 2;
    g.hz  = u.nz / 2;
    g.nbx = (u.nx + 1) / 2;
    g.nby = (u.ny + 1) / 2;
    g.nbz = (u.nz + 1) / 2;
    xtile_shape(g.nbx, g.nby, g.nbz, g.TA, g.TB, g.TC);
    int ta = (g.nbx + g.TA - 1) / g.TA;
    int tc = (g.nbz + g.TC - 1) / g.TC;
    int ia = local % ta;
    int r  = local / ta;
    int ic = r % tc;
    int ib = r / tc;
    g.a0   = ia * g.TA;
    g.b0   = ib * g.TB;
    g.c0   = ic * g.TC;
    g.S    = g.TA | 1; // odd row stride: conflict-free both along a and along c
    g.rows = g.TB * g.TC;
    return g;
}

template <typename T>
__device__ __forceinline__ float load_narrow(const T* p);
template <>
__device__ __forceinline__ float load_narrow<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_narrow<double>(const double* p) {
    return __double2float_rn(__ldg(p)); // src/preprocess.cpp:78
}

// loads the 2x2x2 (or clipped) block at (x0,y0,z0) into v[zi*4+yi*2+xi]
template <typename T>
__device__ __forceinline__ void load_block(const T* __restrict__ box, int X, int Y, int x0, int y0,
                                           int z0, bool wx, bool wy, bool wz, bool vec_ok,
                                           float v[8]) {
#pragma unroll
    for (int zi = 0; zi < 2; ++zi) {
#pragma unroll
        for (int yi = 0; yi < 2; ++yi) {
            bool row_ok = (zi == 0 || wz) && (yi == 0 || wy);
            float e0 = 0.f, e1 = 0.f;
            if (row_ok) {
                const T* p = box + (size_t)x0 + (size_t)X * ((size_t)(y0 + yi) + (size_t)Y * (size_t)(z0 + zi));
                if (wx && vec_ok) {
                    if constexpr (sizeof(T) == 8) {
                        double2 d = __ldg(reinterpret_cast<const double2*>(p));
                        e0 = __double2float_rn(d.x);
                        e1 = __double2float_rn(d.y);
                    } else {
                        float2 d = __ldg(reinterpret_cast<const float2*>(p));
                        e0 = d.x;
                        e1 = d.y;
                    }
                } else {
                    e0 = load_narrow<T>(p);
                    if (wx) e1 = load_narrow<T>(p + 1);
                }
            }
            v[zi * 4 + yi * 2]     = e0;
            v[zi * 4 + yi * 2 + 1] = e1;
        }
    }
}

__device__ __forceinline__ u64 block_max_u64(u64 v, u64* s_red) {
    v = warp_max_u64(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) s_red[w] = v;
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        u64 x  = l < nw ? s_red[l] : 0ull;
        x      = warp_max_u64(x);
        if (l == 0) s_red[0] = x;
    }
    __syncthreads();
    return s_red[0];
}

__global__ void __launch_bounds__(XT_THREADS)
k_forward_generic(const UnitDev* __restrict__ units, UnitState* __restrict__ states,
                  const int2* __restrict__ tiles) {
    extern __shared__ float s_oct[]; // 8 * XT_OCT_FLOATS
    __shared__ u64 s_red[XT_THREADS / 32];

    const int2    tl = tiles[blockIdx.x];
    const UnitDev u  = units[tl.x];
    const XTileGeom g = xtile_geom(u, tl.y);
    const int X = u.nx, Y = u.ny, Z = u.nz;
    const int NB = g.TA * g.TB * g.TC;

    const bool vec_ok = (X % 2 == 0) &&
                        ((reinterpret_cast<uintptr_t>(u.in) & (u.dtype == WC_F64 ? 15u : 7u)) == 0);

    // phase 1: load + transform, a fastest
    for (int q = threadIdx.x; q < NB; q += XT_THREADS) {
        int a_l = q % g.TA;
        int r   = q / g.TA;
        int c_l = r % g.TC;
        int b_l = r / g.TC;
        int a = g.a0 + a_l, b = g.b0 + b_l, c = g.c0 + c_l;
        if (a >= g.nbx || b >= g.nby || c >= g.nbz) continue;
        bool wx = a < g.hx, wy = b < g.hy, wz = c < g.hz;
        float v[8];
        if (u.dtype == WC_F64)
            load_block<double>(static_cast<const double*>(u.in), X, Y, 2 * a, 2 * b, 2 * c, wx, wy, wz, vec_ok, v);
        else
            load_block<float>(static_cast<const float*>(u.in), X, Y, 2 * a, 2 * b, 2 * c, wx, wy, wz, vec_ok, v);
        haar_block_forward(v, wx, wy, wz);
        int si = (b_l * g.TC + c_l) * g.S + a_l;
#pragma unroll
        for (int o = 0; o < 8; ++o) s_oct[o * XT_OCT_FLOATS + si] = v[o];
    }
    __syncthreads();

    // phase 2: write in f order, c fastest; fold the arg-max key
    u64 key = 0ull;
    for (int o = 0; o < 8; ++o) {
        const int sx = o & 1, sy = (o >> 1) & 1, sz = o >> 2;
        for (int e = threadIdx.x; e < NB; e += XT_THREADS) {
            int c_l = e % g.TC;
            int r   = e / g.TC;
            int a_l = r % g.TA;
            int b_l = r / g.TA;
            int a = g.a0 + a_l, b = g.b0 + b_l, c = g.c0 + c_l;
            if (a >= g.nbx || b >= g.nby || c >= g.nbz) continue;
            bool wx = a < g.hx, wy = b < g.hy, wz = c < g.hz;
            if ((sx && !wx) || (sy && !wy) || (sz && !wz)) continue;
            // a singleton block's only value sits at the last index of an odd axis (= a + hx)
            int ip = wx ? a + sx * g.hx : X - 1;
            int jp = wy ? b + sy * g.hy : Y - 1;
            int kp = wz ? c + sz * g.hz : Z - 1;
            uint32_t f = (uint32_t)((ip * Y + jp) * Z + kp);
            float val  = s_oct[o * XT_OCT_FLOATS + (b_l * g.TC + c_l) * g.S + a_l];
            u.coef[f]  = val;
            key        = max_u64(key, make_key(val, f));
            if (f == 0 && isnan(val)) atomicOr(&states[tl.x].flags, 1);
        }
    }
    key = block_max_u64(key, s_red);
    if (threadIdx.x == 0 && key != 0ull) atomicMax(&states[tl.x].key, key);
}

// arg-max key of an already-flat coefficient array (un-fused primitive wc_threshold_pack)
__global__ void __launch_bounds__(CT_THREADS)
k_argmax_flat(const UnitDev* __restrict__ units, UnitState* __restrict__ states,
              const int2* __restrict__ tiles) {
    __shared__ u64 s_red[CT_THREADS / 32];
    const int2    tl = tiles[blockIdx.x];
    const UnitDev u  = units[tl.x];
    u64 key = 0ull;
    for (int j = threadIdx.x; j < CT_ELEMS; j += CT_THREADS) {
        int f = tl.y * CT_ELEMS + j;
        if (f < u.n) {
            float c = u.coef[f];
            key     = max_u64(key, make_key(c, (uint32_t)f));
            if (f == 0 && isnan(c)) atomicOr(&states[tl.x].flags, 1);
        }
    }
    key = block_max_u64(key, s_red);
    if (threadIdx.x == 0 && key != 0ull) atomicMax(&states[tl.x].key, key);
}

__global__ void k_finalize_thresh(UnitState* __restrict__ states, int n_units,
                                  double one_minus_keep, const u64* __restrict__ global_key) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_units) return;
    UnitState s = states[i];
    if (global_key) {
        // EXTENSION (WC_THRESH_GLOBAL): bit 63 of the global key word flags "the very first
        // coefficient of the batch is NaN"
        u64 gk = *global_key;
        states[i].thresh_f = threshold_float(gk & ~(1ull << 63), (gk >> 63) != 0, one_minus_keep);
    } else {
        states[i].thresh_f = threshold_float(s.key, (s.flags & 1) != 0, one_minus_keep);
    }
}

// EXTENSION: batch-wide key = the reference's sequential max rule over the concatenation of the
// units in batch order.  Per-unit keys order by (|c|, lowest f); across units the lowest unit index
// wins ties, so reduce (|c| bits, ~unit) and keep the sign of that unit's winner.
__global__ void k_global_key(const UnitDev* __restrict__ units, const UnitState* __restrict__ states, int n_units,
                             u64* out) {
    __shared__ u64 s_red[32];
    __shared__ int s_first;
    if (threadIdx.x == 0) s_first = n_units;
    __syncthreads();
    u64 best = 0ull;
    for (int i = threadIdx.x; i < n_units; i += blockDim.x) {
        if (units[i].n > 0) atomicMin(&s_first, i);     // the first coefficient of the concatenation lives here
        u64 k = states[i].key;
        if (k == 0ull) continue;
        u64 g = (k & 0xffffffff00000000ull) | ((u64)(0x7fffffffu - (uint32_t)i) << 1) | (k & 1ull);
        best  = max_u64(best, g);
    }
    best = block_max_u64(best, s_red);
    if (threadIdx.x == 0) {
        // empty units contribute no coefficient: the "first coefficient is NaN" flag is the one of the first
        // NON-EMPTY unit (block_max_u64 has synchronised the CTA, s_first is final)
        bool first_nan = s_first < n_units && (states[s_first].flags & 1);
        *out           = best | (first_nan ? (1ull << 63) : 0ull);
        // a second word for multi-rank reductions: does this batch own a non-empty unit at all
        out[1] = s_first < n_units ? 1ull : 0ull;
    }
}

// ============================================================================================
// EXTENSION: quantile thresholds by radix select (WC_THRESH_QUANTILE / WC_THRESH_QUANTILE_GLOBAL)
// ============================================================================================
// The reference's threshold is max * (1 - keep) (src/compressor.cpp:212-216), not a quantile; this mode exists because
// the north star names it: keep the Kt = n - floor(keep * n) coefficients of largest magnitude, i.e. thresh = the
// magnitude of rank Kt (0-based, descending; NaNs rank last and are never kept), mask |c| > thresh as everywhere else —
// ties at the threshold are all dropped, so the kept set does not depend on any order.  Three passes over the
// coefficient scratch, each a histogram of 11 / 11 / 9 key bits (key = the float's bit pattern without the sign: for
// non-NaN magnitudes integer order = float order) among the coefficients that match the prefix chosen so far; a pick
// kernel walks the histogram from the top.  One histogram row per unit, or ONE row for the whole batch (global mode:
// the row is what ranks all-reduce with NCCL between hist and pick, so every rank picks the same bucket).
__global__ void k_q_init(QState* __restrict__ q, const unsigned long long* __restrict__ rank, int nq) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    QState s;
    s.rank = rank[i]; s.nvalid = 0; s.prefix = 0; s.done = 0; s.thresh = __int_as_float(0x7f800000); s.pad = 0;
    q[i] = s;
}

__global__ void __launch_bounds__(CT_THREADS)
k_q_hist(const UnitDev* __restrict__ units, const int2* __restrict__ tiles, const QState* __restrict__ q,
         unsigned long long* __restrict__ hist, int pass, int global) {
    __shared__ uint32_t s_h[Q_BINS];
    const int2    tl = tiles[blockIdx.x];
    const UnitDev u  = units[tl.x];
    const int row = global ? 0 : tl.x;
    const QState st = q[row];
    if (st.done) return;
    for (int i = threadIdx.x; i < Q_BINS; i += CT_THREADS) s_h[i] = 0;
    __syncthreads();
    const int shift = pass == 0 ? 20 : pass == 1 ? 9 : 0;
    const uint32_t mask = pass == 2 ? 511u : 2047u;
    const int pshift = pass == 1 ? 20 : 9;                 // bits below the prefix of the previous passes
    for (int j = threadIdx.x; j < CT_ELEMS; j += CT_THREADS) {
        const int f = tl.y * CT_ELEMS + j;
        if (f < u.n) {
            const uint32_t key = __float_as_uint(u.coef[f]) & 0x7fffffffu;
            if (key <= 0x7f800000u && (pass == 0 || (key >> pshift) == st.prefix))
                atomicAdd(&s_h[(key >> shift) & mask], 1u);
        }
    }
    __syncthreads();
    unsigned long long* h = hist + (size_t)row * Q_BINS;
    for (int i = threadIdx.x; i < Q_BINS; i += CT_THREADS)
        if (s_h[i]) atomicAdd(&h[i], (unsigned long long)s_h[i]);
}

// one warp per row: the bucket that holds rank `rank` counted from the top, then the row is cleared for the next pass
__global__ void k_q_pick(QState* __restrict__ q, unsigned long long* __restrict__ hist, int nq, int pass) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= nq) return;
    QState st = q[row];
    unsigned long long* h = hist + (size_t)row * Q_BINS;
    constexpr int PER = Q_BINS / 32;
    if (!st.done) {
        unsigned long long mine = 0;
        for (int i = 0; i < PER; ++i) mine += h[lane * PER + i];
        // inclusive suffix sums over the lanes (lane 31 holds the top buckets)
        unsigned long long suf = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long x = __shfl_down_sync(0xffffffffu, suf, o);
            if (lane + o < 32) suf += x;
        }
        const unsigned long long total = __shfl_sync(0xffffffffu, suf, 0);
        if (pass == 0) st.nvalid = total;
        if (st.rank >= total) {
            // fewer (matching) coefficients than the rank asks for: in pass 0 this means "keep every non-NaN
            // coefficient"; in later passes it cannot happen (the bucket of the previous pass held the rank)
            st.done = 1;
            st.thresh = -1.0f;
        } else {
            const unsigned long long above = suf - mine;                       // coefficients in higher lanes' buckets
            const bool here = above <= st.rank && st.rank < suf;               // exactly one lane
            const int src = __ffs(__ballot_sync(0xffffffffu, here)) - 1;
            unsigned long long r = st.rank - above;
            int b = 0;
            if (here) {
                unsigned long long c = 0;
                for (b = lane * PER + PER - 1;; --b) {                         // from this lane's top bucket down
                    c += h[b];
                    if (c > r) { r -= c - h[b]; break; }
                }
            }
            b = __shfl_sync(0xffffffffu, b, src);
            r = __shfl_sync(0xffffffffu, r, src);
            const int bits = pass == 2 ? 9 : 11;
            st.prefix = (st.prefix << bits) | (uint32_t)b;
            st.rank = r;
            if (pass == 2) { st.done = 1; st.thresh = __uint_as_float(st.prefix); }
        }
        if (lane == 0) q[row] = st;
    }
    __syncwarp();
    for (int i = 0; i < PER; ++i) h[lane * PER + i] = 0ull;
}

__global__ void k_q_apply(UnitState* __restrict__ states, const QState* __restrict__ q, int n_units, int global) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_units) states[i].thresh_f = q[global ? 0 : i].thresh;
}

cudaError_t launch_q_init(QState* q, const unsigned long long* rank, int nq, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    k_q_init<<<(nq + 255) / 256, 256, 0, st>>>(q, rank, nq);
    return cudaGetLastError();
}
cudaError_t launch_q_hist(const UnitDev* units, const int2* ctiles, int n_ctiles, const QState* q, unsigned long long* hist,
                          int pass, bool global, cudaStream_t st, LaunchStats* ls) {
    if (n_ctiles <= 0) return cudaSuccess;
    ls->begin(KID_Q_HIST, st);
    k_q_hist<<<n_ctiles, CT_THREADS, 0, st>>>(units, ctiles, q, hist, pass, global ? 1 : 0);
    ls->end(st);
    return cudaGetLastError();
}
cudaError_t launch_q_pick(QState* q, unsigned long long* hist, int nq, int pass, cudaStream_t st, LaunchStats* ls) {
    if (nq <= 0) return cudaSuccess;
    ls->begin(KID_Q_PICK, st);
    k_q_pick<<<(nq + 7) / 8, 256, 0, st>>>(q, hist, nq, pass);
    ls->end(st);
    return cudaGetLastError();
}
cudaError_t launch_q_apply(UnitState* states, const QState* q, int n_units, bool global, cudaStream_t st) {
    if (n_units <= 0) return cudaSuccess;
    k_q_apply<<<(n_units + 255) / 256, 256, 0, st>>>(states, q, n_units, global ? 1 : 0);
    return cudaGetLastError();
}

// ============================================================================================
// flat tiles: mask + ordered (run, value) packing
// ============================================================================================
// A flat tile is CT_ELEMS consecutive coefficients of one unit.  Warp w of the CTA owns elements
// [w*256, w*256+256) of the tile, in two rounds of 128 (one float4 per lane per round), so the
// order inside a tile is (warp, round, lane, j).

__device__ __forceinline__ void load4(const float* __restrict__ coef, int f, int n, float c[4]) {
    if (f + 3 < n) {
        float4 v = *reinterpret_cast<const float4*>(coef + f);
        c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) c[j] = (f + j < n) ? coef[f + j] : 0.f;
    }
}

__global__ void __launch_bounds__(CT_THREADS)
k_count_tiles(const UnitDev* __restrict__ units, const UnitState* __restrict__ states,
              const int2* __restrict__ tiles, int* __restrict__ tile_cnt,
              int* __restrict__ tile_last) {
    __shared__ int s_cnt[CT_THREADS / 32], s_last[CT_THREADS / 32];
    const int2    tl = tiles[blockIdx.x];
    const UnitDev u  = units[tl.x];
    const float   tf = states[tl.x].thresh_f;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int base = tl.y * CT_ELEMS + w * 256;
    int cnt = 0, last = -1;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        int f = base + r * 128 + l * 4;
        if (f < u.n) {
            float c[4];
            load4(u.coef, f, u.n, c);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (f + j < u.n && keep_coef(c[j], tf)) { ++cnt; last = f + j; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    }
    if (l == 0) { s_cnt[w] = cnt; s_last[w] = last; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int c = 0, m = -1;
#pragma unroll
        for (int i = 0; i < CT_THREADS / 32; ++i) { c += s_cnt[i]; m = max(m, s_last[i]); }
        tile_cnt[u.ctile0 + tl.y]  = c;
        tile_last[u.ctile0 + tl.y] = m;
    }
}

// one warp per unit: exclusive sum of counts, exclusive running max of last-kept
__global__ void k_scan_tiles(const UnitDev* __restrict__ units, UnitState* __restrict__ states,
                             int n_units, const int* __restrict__ tile_cnt,
                             const int* __restrict__ tile_last, int* __restrict__ tile_base,
                             int* __restrict__ tile_prev) {
    int unit = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int l    = threadIdx.x & 31;
    if (unit >= n_units) return;
    const UnitDev u = units[unit];
    int carry_cnt = 0, carry_last = -1;
    for (int t0 = 0; t0 < u.nctiles; t0 += 32) {
        int t   = t0 + l;
        int c   = t < u.nctiles ? tile_cnt[u.ctile0 + t] : 0;
        int m   = t < u.nctiles ? tile_last[u.ctile0 + t] : -1;
        int ic = c, im = m;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int pc = __shfl_up_sync(0xffffffffu, ic, o);
            int pm = __shfl_up_sync(0xffffffffu, im, o);
            if (l >= o) { ic += pc; im = max(im, pm); }
        }
        int ec = __shfl_up_sync(0xffffffffu, ic, 1);
        int em = __shfl_up_sync(0xffffffffu, im, 1);
        if (l == 0) { ec = 0; em = -1; }
        if (t < u.nctiles) {
            tile_base[u.ctile0 + t] = carry_cnt + ec;
            tile_prev[u.ctile0 + t] = max(carry_last, em);
        }
        carry_cnt += __shfl_sync(0xffffffffu, ic, 31);
        carry_last = max(carry_last, __shfl_sync(0xffffffffu, im, 31));
    }
    if (l == 0) {
        states[unit].npairs = carry_cnt;
        const UnitState st = states[unit];
        if (carry_cnt > 0 && unit_need32(fabsf(key_value(st.key)), st.thresh_f))
            states[unit].flags = st.flags | UNIT_FLAG_NEED32;
    }
}

__global__ void __launch_bounds__(CT_THREADS)
k_emit_tiles(const UnitDev* __restrict__ units, const UnitState* __restrict__ states,
             const int2* __restrict__ tiles, const int* __restrict__ tile_base,
             const int* __restrict__ tile_prev) {
    constexpr int NE = (CT_THREADS / 32) * 2; // (warp, round) entries
    __shared__ int s_cnt[NE], s_last[NE];
    const int2    tl = tiles[blockIdx.x];
    const UnitDev u  = units[tl.x];
    const float   tf = states[tl.x].thresh_f;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int base = tl.y * CT_ELEMS + w * 256;
    const uint32_t lt = lanemask_lt();

    float c[2][4];
    int   flags[2], lpre[2], llast[2], lsrc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        int f = base + r * 128 + l * 4;
        int fl = 0, last = -1;
        if (f < u.n) {
            load4(u.coef, f, u.n, c[r]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (f + j < u.n && keep_coef(c[r][j], tf)) { fl |= 1 << j; last = f + j; }
        }
        int cnt = __popc(fl);
        uint32_t b0 = __ballot_sync(0xffffffffu, cnt & 1);
        uint32_t b1 = __ballot_sync(0xffffffffu, cnt & 2);
        uint32_t b2 = __ballot_sync(0xffffffffu, cnt & 4);
        uint32_t any = b0 | b1 | b2;
        flags[r] = fl;
        lpre[r]  = __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
        llast[r] = last;
        uint32_t lower = any & lt;
        lsrc[r]  = lower ? 31 - __clz(lower) : -1; // nearest lower lane that kept something
        if (l == 31) {
            s_cnt[w * 2 + r] = lpre[r] + cnt;
        }
        // last kept of the whole (warp, round): highest lane with any
        int hi = any ? 31 - __clz(any) : 0;
        int wl = __shfl_sync(0xffffffffu, last, hi);
        if (l == 0) s_last[w * 2 + r] = any ? wl : -1;
    }
    __syncthreads();
    const int tbase = tile_base[u.ctile0 + tl.y];
    const int tprev = tile_prev[u.ctile0 + tl.y];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        int e = w * 2 + r;
        int pos = tbase, prev = tprev;
        for (int i = 0; i < e; ++i) { pos += s_cnt[i]; prev = max(prev, s_last[i]); }
        pos += lpre[r];
        int nb = __shfl_sync(0xffffffffu, llast[r], lsrc[r] < 0 ? 0 : lsrc[r]);
        if (lsrc[r] >= 0) prev = nb;
        int f = base + r * 128 + l * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (flags[r] & (1 << j)) {
                wc_pair p;
                p.run = f + j - prev - 1;
                p.val = c[r][j];
                u.out[pos++] = p;
                prev = f + j;
            }
        }
    }
}

// ============================================================================================
// rle_decode: idx_i = sum_{j<=i} run_j + i; coef[idx_i] = val_i if idx_i < total
// ============================================================================================
__device__ __forceinline__ long long block_excl_scan_ll(long long v, long long* s_w, long long& total) {
    // inclusive warp scan
    int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long p = __shfl_up_sync(0xffffffffu, inc, o);
        if (l >= o) inc += p;
    }
    if (l == 31) s_w[w] = inc;
    __syncthreads();
    long long wpre = 0, tot = 0;
    int nw = blockDim.x >> 5;
    for (int i = 0; i < nw; ++i) {
        long long x = s_w[i];
        if (i < w) wpre += x;
        tot += x;
    }
    total = tot;
    __syncthreads();
    return wpre + inc - v;
}

__global__ void __launch_bounds__(PT_THREADS)
k_rle_tile_sums(const DecUnitDev* __restrict__ units, const int2* __restrict__ tiles,
                long long* __restrict__ tile_sum, int* __restrict__ err) {
    __shared__ long long s_w[PT_THREADS / 32];
    const int2       tl = tiles[blockIdx.x];
    const DecUnitDev u  = units[tl.x];
    const int K = u.npairs_dev ? *u.npairs_dev : u.npairs;
    long long s = 0;
    int p0 = tl.y * PT_PAIRS + threadIdx.x * (PT_PAIRS / PT_THREADS);
#pragma unroll
    for (int j = 0; j < PT_PAIRS / PT_THREADS; ++j) {
        int p = p0 + j;
        if (p < K) {
            int run = u.pairs[p].run;
            if (run < 0) atomicOr(err, 1);
            s += (long long)run + 1;
        }
    }
    long long tot;
    block_excl_scan_ll(s, s_w, tot);
    if (threadIdx.x == 0) tile_sum[u.ptile0 + tl.y] = tot;
}

__global__ void k_rle_scan(const DecUnitDev* __restrict__ units, int n_units,
                           long long* __restrict__ tile_sum) {
    // in place: tile_sum -> exclusive prefix (one warp per unit)
    int unit = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int l    = threadIdx.x & 31;
    if (unit >= n_units) return;
    const DecUnitDev u = units[unit];
    long long carry = 0;
    for (int t0 = 0; t0 < u.nptiles; t0 += 32) {
        int t = t0 + l;
        long long v = t < u.nptiles ? tile_sum[u.ptile0 + t] : 0;
        long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long p = __shfl_up_sync(0xffffffffu, inc, o);
            if (l >= o) inc += p;
        }
        if (t < u.nptiles) tile_sum[u.ptile0 + t] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

__global__ void __launch_bounds__(PT_THREADS)
k_rle_scatter(const DecUnitDev* __restrict__ units, const int2* __restrict__ tiles,
              const long long* __restrict__ tile_pre) {
    __shared__ long long s_w[PT_THREADS / 32];
    const int2       tl = tiles[blockIdx.x];
    const DecUnitDev u  = units[tl.x];
    constexpr int PER = PT_PAIRS / PT_THREADS;
    const int K = u.npairs_dev ? *u.npairs_dev : u.npairs;
    wc_pair pr[PER];
    long long s = 0;
    int p0 = tl.y * PT_PAIRS + threadIdx.x * PER;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        int p = p0 + j;
        if (p < K) {
            pr[j] = u.pairs[p];
            s += (long long)pr[j].run + 1;
        } else {
            pr[j].run = 0;
            pr[j].val = 0.f;
        }
    }
    long long tot;
    long long pre = tile_pre[u.ptile0 + tl.y] + block_excl_scan_ll(s, s_w, tot);
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        int p = p0 + j;
        if (p < K) {
            long long idx = pre + pr[j].run; // = sum_{q<p}(run_q+1) + run_p
            if (idx >= 0 && idx < (long long)u.total) u.coef[idx] = pr[j].val;
            pre += (long long)pr[j].run + 1;
        }
    }
}

// ============================================================================================
// inverse transform
// ============================================================================================
template <typename T>
__device__ __forceinline__ void store_out(T* p, float v);
template <>
__device__ __forceinline__ void store_out<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_out<double>(double* p, float v) { *p = (double)v; }

template <typename T>
__device__ __forceinline__ void store_block(T* __restrict__ box, int X, int Y, int x0, int y0,
                                            int z0, bool wx, bool wy, bool wz, bool vec_ok,
                                            const float v[8]) {
#pragma unroll
    for (int zi = 0; zi < 2; ++zi) {
#pragma unroll
        for (int yi = 0; yi < 2; ++yi) {
            bool row_ok = (zi == 0 || wz) && (yi == 0 || wy);
            if (!row_ok) continue;
            T* p = box + (size_t)x0 + (size_t)X * ((size_t)(y0 + yi) + (size_t)Y * (size_t)(z0 + zi));
            float e0 = v[zi * 4 + yi * 2], e1 = v[zi * 4 + yi * 2 + 1];
            if (wx && vec_ok) {
                if constexpr (sizeof(T) == 8) {
                    *reinterpret_cast<double2*>(p) = make_double2((double)e0, (double)e1);
                } else {
                    *reinterpret_cast<float2*>(p) = make_float2(e0, e1);
                }
            } else {
                store_out<T>(p, e0);
                if (wx) store_out<T>(p + 1, e1);
            }
        }
    }
}

__global__ void __launch_bounds__(XT_THREADS)
k_inverse_generic(const InvUnitDev* __restrict__ units, const int2* __restrict__ tiles) {
    extern __shared__ float s_oct[];
    const int2       tl = tiles[blockIdx.x];
    const InvUnitDev iu = units[tl.x];
    UnitDev u;
    u.nx = iu.nx; u.ny = iu.ny; u.nz = iu.nz;
    const XTileGeom g = xtile_geom(u, tl.y);
    const int X = u.nx, Y = u.ny, Z = u.nz;
    const int NB = g.TA * g.TB * g.TC;

    // phase 1: gather the 8 sub-band values of every block, c fastest (coalesced along k')
    for (int o = 0; o < 8; ++o) {
        const int sx = o & 1, sy = (o >> 1) & 1, sz = o >> 2;
        for (int e = threadIdx.x; e < NB; e += XT_THREADS) {
            int c_l = e % g.TC;
            int r   = e / g.TC;
            int a_l = r % g.TA;
            int b_l = r / g.TA;
            int a = g.a0 + a_l, b = g.b0 + b_l, c = g.c0 + c_l;
            if (a >= g.hx || b >= g.hy || c >= g.hz) continue; // only full blocks contribute
            int ip = a + sx * g.hx, jp = b + sy * g.hy, kp = c + sz * g.hz;
            s_oct[o * XT_OCT_FLOATS + (b_l * g.TC + c_l) * g.S + a_l] =
                iu.coef[(size_t)(ip * Y + jp) * Z + kp];
        }
    }
    __syncthreads();

    const bool vec_ok = (X % 2 == 0) &&
                        ((reinterpret_cast<uintptr_t>(iu.out) & (iu.dtype == WC_F64 ? 15u : 7u)) == 0);
    // phase 2: inverse + store, a fastest
    for (int q = threadIdx.x; q < NB; q += XT_THREADS) {
        int a_l = q % g.TA;
        int r   = q / g.TA;
        int c_l = r % g.TC;
        int b_l = r / g.TC;
        int a = g.a0 + a_l, b = g.b0 + b_l, c = g.c0 + c_l;
        if (a >= g.nbx || b >= g.nby || c >= g.nbz) continue;
        bool wx = a < g.hx, wy = b < g.hy, wz = c < g.hz;
        float v[8];
        if (wx && wy && wz) {
            int si = (b_l * g.TC + c_l) * g.S + a_l;
#pragma unroll
            for (int o = 0; o < 8; ++o) v[o] = s_oct[o * XT_OCT_FLOATS + si];
            haar_block_inverse_full(v);
        } else {
            // a block that touches the trailing plane of an odd axis: the reference's inverse
            // zero-fills that plane (src/decompressor.cpp:98, :123, :144 — `restored` is
            // value-initialised and only 2*(n/2) entries are written) and every later pass only
            // combines it with itself, so the whole clipped block is +0.
#pragma unroll
            for (int o = 0; o < 8; ++o) v[o] = 0.f;
        }
        if (iu.dtype == WC_F64)
            store_block<double>(static_cast<double*>(iu.out), X, Y, 2 * a, 2 * b, 2 * c, wx, wy, wz, vec_ok, v);
        else
            store_block<float>(static_cast<float*>(iu.out), X, Y, 2 * a, 2 * b, 2 * c, wx, wy, wz, vec_ok, v);
    }
}

// ============================================================================================
// RMSE
// ============================================================================================
// Fixed reduction tree (deterministic): thread-sequential over 8 consecutive elements, warp xor
// tree, 8 warp partials summed in order, then per unit lane-strided sequential + warp tree.
__global__ void __launch_bounds__(CT_THREADS)
k_rmse_tiles(const RmseUnitDev* __restrict__ units, const int2* __restrict__ tiles,
             double* __restrict__ tile_sum) {
    __shared__ double s_w[CT_THREADS / 32];
    const int2        tl = tiles[blockIdx.x];
    const RmseUnitDev u  = units[tl.x];
    // thread t takes the element pairs (2t, 2t+1) + 2 * CT_THREADS * j: a warp reads whole 512-byte (f64) /
    // 256-byte (f32) runs with 16- / 8-byte vector loads; the order of the additions is fixed
    const int tile0 = tl.y * CT_ELEMS;
    double s = 0.0;
    auto term = [&](float av, float bv) {
        float  df = __fsub_rn(av, bv); // float subtraction, src/calc-loss.cpp:33
        double d  = (double)df;
        s = __dadd_rn(s, __dmul_rn(d, d));
    };
    const bool a64 = u.a_dtype == WC_F64, b64 = u.b_dtype == WC_F64;
    const bool vec = tile0 + CT_ELEMS <= u.n &&
                     (reinterpret_cast<uintptr_t>(u.a) & (a64 ? 15u : 7u)) == 0 &&
                     (reinterpret_cast<uintptr_t>(u.b) & (b64 ? 15u : 7u)) == 0;
    if (vec) {
        float2 av[CT_ELEMS / CT_THREADS / 2], bv[CT_ELEMS / CT_THREADS / 2];
#pragma unroll
        for (int j = 0; j < CT_ELEMS / CT_THREADS / 2; ++j) {
            const int f = tile0 + 2 * (threadIdx.x + CT_THREADS * j);
            if (a64) {
                const double2 v = __ldcs(reinterpret_cast<const double2*>(static_cast<const double*>(u.a) + f));
                av[j] = make_float2(__double2float_rn(v.x), __double2float_rn(v.y));   // src/preprocess.cpp:78
            } else {
                av[j] = __ldcs(reinterpret_cast<const float2*>(static_cast<const float*>(u.a) + f));
            }
            if (b64) {
                const double2 v = __ldcs(reinterpret_cast<const double2*>(static_cast<const double*>(u.b) + f));
                bv[j] = make_float2(__double2float_rn(v.x), __double2float_rn(v.y));
            } else {
                bv[j] = __ldcs(reinterpret_cast<const float2*>(static_cast<const float*>(u.b) + f));
            }
        }
#pragma unroll
        for (int j = 0; j < CT_ELEMS / CT_THREADS / 2; ++j) { term(av[j].x, bv[j].x); term(av[j].y, bv[j].y); }
    } else {
        // ragged last tile / unaligned boxes: same element-to-thread map, scalar loads
#pragma unroll
        for (int j = 0; j < CT_ELEMS / CT_THREADS / 2; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int f = tile0 + 2 * (threadIdx.x + CT_THREADS * j) + e;
                if (f < u.n) {
                    float av = a64 ? __double2float_rn(static_cast<const double*>(u.a)[f]) : static_cast<const float*>(u.a)[f];
                    float bv = b64 ? __double2float_rn(static_cast<const double*>(u.b)[f]) : static_cast<const float*>(u.b)[f];
                    term(av, bv);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) s_w[w] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < CT_THREADS / 32; ++i) t = __dadd_rn(t, s_w[i]);
        tile_sum[u.ctile0 + tl.y] = t;
    }
}

__global__ void k_rmse_final(const RmseUnitDev* __restrict__ units, int n_units,
                             const double* __restrict__ tile_sum, double* __restrict__ rmse) {
    int unit = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int l    = threadIdx.x & 31;
    if (unit >= n_units) return;
    const RmseUnitDev u = units[unit];
    double s = 0.0;
    for (int t = l; t < u.nctiles; t += 32) s = __dadd_rn(s, tile_sum[u.ctile0 + t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
    // sqrt(sum / (xdim*ydim*zdim)), src/calc-loss.cpp:39
    if (l == 0) rmse[unit] = __dsqrt_rn(__ddiv_rn(s, (double)u.n));
}

// ============================================================================================
// per-unit min / max of the narrowed values (ingest statistics, src/preprocess.cpp:82-88)
// ============================================================================================
// `if (value < min) min = value; if (value > max) max = value;` — NaN never updates either.
__global__ void __launch_bounds__(CT_THREADS)
k_minmax_tiles(const RmseUnitDev* __restrict__ units, const int2* __restrict__ tiles,
               float2* __restrict__ tile_mm) {
    __shared__ float s_lo[CT_THREADS / 32], s_hi[CT_THREADS / 32];
    const int2        tl = tiles[blockIdx.x];
    const RmseUnitDev u  = units[tl.x];
    float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);   // +inf / -inf = "no value yet"
    for (int j = threadIdx.x; j < CT_ELEMS; j += CT_THREADS) {
        int f = tl.y * CT_ELEMS + j;
        if (f < u.n) {
            float v = u.a_dtype == WC_F64 ? __double2float_rn(static_cast<const double*>(u.a)[f])
                                          : static_cast<const float*>(u.a)[f];
            if (v < lo) lo = v;
            if (v > hi) hi = v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s_lo[w] = lo; s_hi[w] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < CT_THREADS / 32; ++i) { lo = fminf(lo, s_lo[i]); hi = fmaxf(hi, s_hi[i]); }
        tile_mm[u.ctile0 + tl.y] = make_float2(lo, hi);
    }
}

__global__ void k_minmax_final(const RmseUnitDev* __restrict__ units, int n_units,
                               const float2* __restrict__ tile_mm, float2* __restrict__ out) {
    int unit = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int l    = threadIdx.x & 31;
    if (unit >= n_units) return;
    const RmseUnitDev u = units[unit];
    float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
    for (int t = l; t < u.nctiles; t += 32) {
        float2 m = tile_mm[u.ctile0 + t];
        lo = fminf(lo, m.x);
        hi = fmaxf(hi, m.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (l == 0) out[unit] = make_float2(lo, hi);
}

// ============================================================================================
// dense gather of the packed slots (for the single D2H of wc_plan_fetch(WC_HOST))
// ============================================================================================
__global__ void k_unit_offsets(const UnitState* __restrict__ states, int n_units,
                               long long* __restrict__ offsets /* n_units+1 */,
                               long long* __restrict__ running /* optional: in = base, out = base + total */) {
    // single CTA exclusive scan over units
    __shared__ long long s_w[32];
    __shared__ long long s_carry;
    if (threadIdx.x == 0) s_carry = running ? *running : 0;
    __syncthreads();
    for (int i0 = 0; i0 < n_units; i0 += blockDim.x) {
        int i = i0 + threadIdx.x;
        long long v = i < n_units ? states[i].npairs : 0;
        long long tot;
        long long ex = block_excl_scan_ll(v, s_w, tot);
        if (i < n_units) offsets[i] = s_carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        offsets[n_units] = s_carry;
        if (running) *running = s_carry;
    }
}

__global__ void k_gather_dense(const UnitDev* __restrict__ units, const UnitState* __restrict__ states,
                               const long long* __restrict__ offsets, wc_pair* __restrict__ dense,
                               int n_units, int ctas_per_unit) {
    int unit = blockIdx.x / ctas_per_unit;
    int part = blockIdx.x % ctas_per_unit;
    if (unit >= n_units) return;
    const wc_pair* src = units[unit].out;
    wc_pair*       dst = dense + offsets[unit];
    int k = states[unit].npairs;
    for (int i = part * blockDim.x + threadIdx.x; i < k; i += ctas_per_unit * blockDim.x) dst[i] = src[i];
}

// ============================================================================================
// launchers
// ============================================================================================
#define WC_LAUNCH_CHECK()                        \
    do {                                         \
        cudaError_t e__ = cudaGetLastError();    \
        if (e__ != cudaSuccess) return e__;      \
    } while (0)

// Function attributes are per device: set on every launch (a cheap driver call) rather than cached in a
// process-wide static, so that ctxs on different GPUs — and different host threads — stay independent.
static cudaError_t ensure_xt_smem() {
    cudaError_t e;
    e = cudaFuncSetAttribute(k_forward_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, XT_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_inverse_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, XT_SMEM_BYTES);
}

cudaError_t launch_forward_generic(const UnitDev* units, UnitState* states, const int2* tiles,
                                   int n_tiles, cudaStream_t st, LaunchStats* ls) {
    if (n_tiles <= 0) return cudaSuccess;
    cudaError_t e = ensure_xt_smem();
    if (e != cudaSuccess) return e;
    ls->begin(KID_FORWARD_GENERIC, st);
    k_forward_generic<<<n_tiles, XT_THREADS, XT_SMEM_BYTES, st>>>(units, states, tiles);
    ls->end(st);
    WC_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_argmax_flat(const UnitDev* units, UnitState* states, const int2* ctiles,
                               int n_ctiles, cudaStream_t st, LaunchStats* ls) {
    if (n_ctiles <= 0) return cudaSuccess;
    ls->begin(KID_ARGMAX_FLAT, st);
    k_argmax_flat<<<n_ctiles, CT_THREADS, 0, st>>>(units, states, ctiles);
    ls->end(st);
    WC_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_finalize_thresh(UnitState* states, int n_units, double one_minus_keep,
                                   const u64* global_key, cudaStream_t st, LaunchStats* ls) {
    if (n_units <= 0) return cudaSuccess;
    ls->begin(KID_FINALIZE, st);
    k_finalize_thresh<<<(n_units + 255) / 256, 256, 0, st>>>(states, n_units, one_minus_keep, global_key);
    ls->end(st);
    WC_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_global_key(const UnitDev* units, const UnitState* states, int n_units, u64* out,
                              cudaStream_t st, LaunchStats* ls) {
    ls->begin(KID_GLOBAL_KEY, st);
    k_global_key<<<1, 1024, 0, st>>>(units, states, n_units, out);
    ls->end(st);
    WC_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_pack_generic(const UnitDev* units, UnitState* states, int n_units,
                                const int2* ctiles, int n_ctiles, int* tile_cnt, int* tile_last,
                                int* tile_base, int* tile_prev, cudaStream_t st,
                                LaunchStats* ls) {
    if (n_units <= 0) return cudaSuccess;
    if (n_ctiles > 0) {
        ls->begin(KID_COUNT, st);
        k_count_tiles<<<n_ctiles, CT_THREADS, 0, st>>>(units, states, ctiles, tile_cnt, tile_last);
        ls->end(st);
        WC_LAUNCH_CHECK();
    }
    ls->begin(KID_SCAN, st);
    k_scan_tiles<<<(n_units + 7) / 8, 256, 0, st>>>(units, states, n_units, tile_cnt, tile_last,
                                                    tile_base, tile_prev);
    ls->end(st);
    WC_LAUNCH_CHECK();
    if (n_ctiles > 0) {
        ls->begin(KID_EMIT, st);
        k_emit_tiles<<<n_ctiles, CT_THREADS, 0, st>>>(units, states, ctiles, tile_base, tile_prev);
        ls->end(st);
        WC_LAUNCH_CHECK();
    }
    return cudaSuccess;
}

cudaError_t launch_rle_decode_generic(const DecUnitDev* units, int n_units, const int2* ptiles,
                                      int n_ptiles, long long* tile_sum, int* err, cudaStream_t st,
                                      LaunchStats* ls) {
    if (n_units <= 0 || n_ptiles <= 0) return cudaSuccess;
    ls->begin(KID_RLE_SUMS, st);
    k_rle_tile_sums<<<n_ptiles, PT_THREADS, 0, st>>>(units, ptiles, tile_sum, err);
    ls->end(st);
    WC_LAUNCH_CHECK();
    ls->begin(KID_RLE_SCAN, st);
    k_rle_scan<<<(n_units + 7) / 8, 256, 0, st>>>(units, n_units, tile_sum);
    ls->end(st);
    WC_LAUNCH_CHECK();
    ls->begin(KID_RLE_SCATTER, st);
    k_rle_scatter<<<n_ptiles, PT_THREADS, 0, st>>>(units, ptiles, tile_sum);
    ls->end(st);
    WC_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_inverse_generic(const InvUnitDev* units, const int2* tiles, int n_tiles,
                                   cudaStream_t st, LaunchStats* ls) {
    if (n_tiles <= 0) return cudaSuccess;
    cudaError_t e = ensure_xt_smem();
    if (e != cudaSuccess) return e;
    ls->begin(KID_INVERSE_GENERIC, st);
    k_inverse_generic<<<n_tiles, XT_THREADS, XT_SMEM_BYTES, st>>>(units, tiles);
    ls->end(st);
    WC_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_rmse_generic(const RmseUnitDev* units, int n_units, const int2* ctiles,
                                int n_ctiles, double* tile_sum, double* rmse, cudaStream_t st,
                                LaunchStats* ls) {
    if (n_units <= 0) return cudaSuccess;
    if (n_ctiles > 0) {
        ls->begin(KID_RMSE_TILES, st);
        k_rmse_tiles<<<n_ctiles, CT_THREADS, 0, st>>>(units, ctiles, tile_sum);
        ls->end(st);
        WC_LAUNCH_CHECK();
    }
    ls->begin(KID_RMSE_FINAL, st);
    k_rmse_final<<<(n_units + 7) / 8, 256, 0, st>>>(units, n_units, tile_sum, rmse);
    ls->end(st);
    WC_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_minmax_generic(const RmseUnitDev* units, int n_units, const int2* ctiles, int n_ctiles,
                                  float2* tile_mm, float2* out, cudaStream_t st, LaunchStats* ls) {
    if (n_units <= 0) return cudaSuccess;
    if (n_ctiles > 0) {
        ls->begin(KID_MINMAX_TILES, st);
        k_minmax_tiles<<<n_ctiles, CT_THREADS, 0, st>>>(units, ctiles, tile_mm);
        ls->end(st);
        WC_LAUNCH_CHECK();
    }
    ls->begin(KID_MINMAX_FINAL, st);
    k_minmax_final<<<(n_units + 7) / 8, 256, 0, st>>>(units, n_units, tile_mm, out);
    ls->end(st);
    WC_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_gather_dense(const UnitDev* units, const UnitState* states, int n_units,
                                long long* offsets, wc_pair* dense, bool offsets_only,
                                cudaStream_t st, LaunchStats* ls, long long* running) {
    if (n_units <= 0) return cudaSuccess;
    if (offsets_only) {
        ls->begin(KID_OFFSETS, st);
        k_unit_offsets<<<1, 1024, 0, st>>>(states, n_units, offsets, running);
        ls->end(st);
        WC_LAUNCH_CHECK();
        return cudaSuccess;
    }
    const int cpu = 4;
    ls->begin(KID_GATHER, st);
    k_gather_dense<<<n_units * cpu, 256, 0, st>>>(units, states, offsets, dense, n_units, cpu);
    ls->end(st);
    WC_LAUNCH_CHECK();
    return cudaSuccess;
}

// wc_plan_set_inputs: new input addresses for the same unit table, stream-ordered (no host synchronisation)
__global__ void k_patch_inputs(UnitDev* __restrict__ units, const void* const* __restrict__ ptrs, int n_units) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_units) units[i].in = ptrs[i];
}
cudaError_t launch_patch_inputs(UnitDev* units, const void* const* ptrs, int n_units, cudaStream_t st,
                                LaunchStats* ls) {
    if (n_units <= 0) return cudaSuccess;
    ls->begin(KID_PATCH_INPUTS, st);
    k_patch_inputs<<<(n_units + 255) / 256, 256, 0, st>>>(units, ptrs, n_units);
    ls->end(st);
    WC_LAUNCH_CHECK();
    return cudaSuccess;
}

} // namespace wc
