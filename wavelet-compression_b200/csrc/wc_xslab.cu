// Fused on-chip kernels for boxes of ANY shape (odd dimensions, nz % 4 != 0, rows that are not 16-byte multiples,
// unaligned pointers): the x-slab classes FUSED_CLS_XS1 / XS2 / XS4 / XS8.
//
// The y-slab kernels of wc_fused.cu owe their speed to 16-byte vector accesses, and with them to even dimensions and
// nz % 4 == 0.  Everything else used to take the generic multi-kernel path (coefficient scratch in HBM: 20N + 8K bytes of
// traffic).  These kernels keep such a unit on chip as well: HBM sees the input once and the pairs once.
//
// Decomposition by X-SLABS.  CTA r of S takes the block columns a in [r*na, (r+1)*na) (na = ceil(hx / S)); the trailing
// plane x = X-1 of an odd X (which passes through the x stage, src/compressor.cpp:153-175) goes to the last CTA.  The
// coefficients of a slab are whole PLANES i' of the flat order f = (i'*Y + j')*Z + k' (src/compressor.cpp:178-181): the
// low planes [a0, a1) and the high planes [hx+a0, hx+a1) (+ the plane X-1) — two contiguous flat ranges.  So
//   compress:   a cluster of S CTAs exchanges, through distributed shared memory, one arg-max key and four words (kept
//               count and last kept flat index of either range) per CTA — nothing else is needed to place a CTA's pairs
//               in the unit's ordered pair list and to know the zero run in front of its first pair;
//   decompress: the S slab items of a unit are independent once the position of every plane in the pair list is known:
//               tab[i'] = (first pair at/after flat index i'*Y*Z, flat index of the pair before it).  The compress kernel
//               writes that table for free; for streams the index kernels of wc_fused.cu build it (seglen = Y*Z).
//               S = 1 needs no table: the whole list is walked once.
// The trailing element of an odd axis passes through the forward stage of that axis and is ZEROED by the inverse
// (src/decompressor.cpp:90-156: `restored` starts zero-filled and only 2*(n/2) entries are written), so the decoder never
// reads the singleton planes / rows / columns of the coefficient array: it writes +0 into the trailing cells.
//
// C, the coefficient array in shared memory, holds the CTA's planes at a stride of PS = (Y*Z | 31) + 2 words (PS % 32 == 1):
// the lanes of a warp work on consecutive block columns a, i.e. on consecutive planes, and hit 32 different banks.
// All global accesses are element-wise (4 or 8 bytes per lane, consecutive lanes on consecutive block columns): coalesced
// without any alignment requirement.  Simplicity over the last 20 %: these shapes are rare in AMR plotfiles (boxes are
// multiples of the blocking factor); the point is that they no longer fall off the fused path.
#include <cooperative_groups.h>

#include "wc_common.cuh"
#include "wc_fused.h"

namespace cg = cooperative_groups;

namespace wc {

constexpr int XS_NT     = 512;
constexpr int XS_NW     = XS_NT / 32;
constexpr int XS_CWORDS = 56320;               // words of C per CTA (220 KB: one CTA per SM either way)
constexpr int XS_MAXPL  = 272;                 // local planes per CTA: at most 2 * 128 + 1 (nx <= 256)
constexpr int XS_PPT    = 8;                   // pairs per thread per tile of the decoder's block scan
// shared memory layout (bytes)
constexpr int XS_OFF_CNT  = XS_CWORDS * 4;                 // int[XS_MAXPL]   kept coefficients per local plane
constexpr int XS_OFF_LAST = XS_OFF_CNT + XS_MAXPL * 4;     // int[XS_MAXPL]   flat index of the last kept one, or -1
constexpr int XS_OFF_BASE = XS_OFF_LAST + XS_MAXPL * 4;    // int[XS_MAXPL]   pairs of the same range in front of the plane
constexpr int XS_OFF_PREV = XS_OFF_BASE + XS_MAXPL * 4;    // int[XS_MAXPL]   last kept flat index of the range before it
constexpr int XS_OFF_RED  = XS_OFF_PREV + XS_MAXPL * 4;    // u64[64]         block reductions
constexpr int XS_OFF_X1   = XS_OFF_RED + 64 * 8;           // u64[2][2]       exchange 1: arg-max key, NaN-at-f=0
constexpr int XS_OFF_X2   = XS_OFF_X1 + 32;                // int[2][4]       exchange 2: count / last of either range
constexpr int XS_OFF_MISC = XS_OFF_X2 + 32;                // int[8]
constexpr int XS_SMEM     = XS_OFF_MISC + 32;
static_assert(XS_SMEM <= 232448, "exceeds the 227 KB per-CTA shared memory of sm_100");

__host__ __device__ inline int xs_plane_stride(int yz) { return (yz | 31) + 2; }

// Cluster size / slab count of a box on the x-slab classes: the smallest S in {1, 2, 4, 8} whose slab fits C; 0 = none.
int xs_slabs(int nx, int ny, int nz) {
    if (nx < 1 || ny < 1 || nz < 1 || nx > 256) return 0;
    const long long yz = (long long)ny * nz;
    if (yz > XS_CWORDS) return 0;
    const int ps = xs_plane_stride((int)yz), hx = nx / 2;
    for (int S = 1; S <= 8; S *= 2) {
        const int na = (hx + S - 1) / S;
        if ((long long)(2 * na + (nx & 1)) * ps <= XS_CWORDS) return S;
    }
    return 0;
}

struct XGeom {
    int X, Y, Z, hx, hy, hz, ox, oy, oz;
    int YZ, PS;
    int a0, nl, own1;      // this slab: first block column, block columns, owns the trailing plane x = X-1
    int npl;               // local planes: 2 * nl + own1
    __device__ __forceinline__ void init(int nx, int ny, int nz, int S, int rank) {
        X = nx; Y = ny; Z = nz;
        hx = nx >> 1; hy = ny >> 1; hz = nz >> 1;
        ox = nx & 1; oy = ny & 1; oz = nz & 1;
        YZ = ny * nz;
        PS = xs_plane_stride(YZ);
        const int na = (hx + S - 1) / S;
        a0 = min(hx, rank * na);
        nl = min(hx, a0 + na) - a0;
        own1 = (ox && rank == S - 1) ? 1 : 0;
        npl = 2 * nl + own1;
    }
    // global plane i' of local plane p (low planes, high planes, then the trailing plane)
    __device__ __forceinline__ int gplane(int p) const { return p < nl ? a0 + p : (p < 2 * nl ? hx + a0 + (p - nl) : X - 1); }
};

template <class T> __device__ __forceinline__ float xs_load(const void* base, size_t idx);
template <> __device__ __forceinline__ float xs_load<double>(const void* base, size_t idx) {
    return __double2float_rn(__ldg(static_cast<const double*>(base) + idx));      // src/preprocess.cpp:78
}
template <> __device__ __forceinline__ float xs_load<float>(const void* base, size_t idx) {
    return __ldg(static_cast<const float*>(base) + idx);
}

// Phase A of one slab: every thread takes generalized blocks (al fastest, then b, then c): up to 2 x 2 x 2 cells, a single
// cell wide along an axis whose trailing element it holds.  Coefficients go to C, the running arg-max key (make_key:
// largest |c|, lowest flat index, NaNs skipped — std::max_element of src/compressor.cpp:212-215) stays in a register.
template <class T, bool MM>
__device__ __forceinline__ void xs_phase_a(const XGeom& g, const void* in, float* C, u64& key, float& vmn, float& vmx) {
    const int nbx = g.nl + g.own1, nby = g.hy + g.oy, nbz = g.hz + g.oz;
    const int nblk = nbx * nby * nbz;
#pragma unroll 1
    for (int q = threadIdx.x; q < nblk; q += XS_NT) {
        const int al = q % nbx, t = q / nbx, b = t % nby, c = t / nby;
        const bool wx = al < g.nl, wy = b < g.hy, wz = c < g.hz;
        const int x0 = wx ? 2 * (g.a0 + al) : g.X - 1, y0 = wy ? 2 * b : g.Y - 1, z0 = wz ? 2 * c : g.Z - 1;
        float v[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const int xi = o & 1, yi = (o >> 1) & 1, zi = o >> 2;
            const bool valid = (xi == 0 || wx) && (yi == 0 || wy) && (zi == 0 || wz);
            v[o] = 0.f;
            if (valid) {
                v[o] = xs_load<T>(in, ((size_t)(z0 + zi) * g.Y + (size_t)(y0 + yi)) * g.X + (size_t)(x0 + xi));
                if (MM) { vmn = fminf(vmn, v[o]); vmx = fmaxf(vmx, v[o]); }
            }
        }
        haar_block_forward(v, wx, wy, wz);
        const int pl0 = wx ? al : 2 * g.nl, pl1 = g.nl + al;           // local planes of the low / high x band
        const int gi0 = wx ? g.a0 + al : g.X - 1, gi1 = g.hx + g.a0 + al;
        const int j0 = wy ? b : g.Y - 1, j1 = g.hy + b, k0 = wz ? c : g.Z - 1, k1 = g.hz + c;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const int sx = o & 1, sy = (o >> 1) & 1, sz = o >> 2;
            const bool valid = (sx == 0 || wx) && (sy == 0 || wy) && (sz == 0 || wz);
            if (valid) {
                const int j = sy ? j1 : j0, k = sz ? k1 : k0;
                C[(sx ? pl1 : pl0) * g.PS + j * g.Z + k] = v[o];
                const uint32_t f = (uint32_t)(((sx ? gi1 : gi0) * g.Y + j) * g.Z + k);
                key = max_u64(key, make_key(v[o], f));
            }
        }
    }
}

__device__ __forceinline__ uint32_t xs_order_code(float f) {      // monotone float -> uint (as float_order_code, wc_fused.cu)
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// ---- compress ------------------------------------------------------------------------------------------------------
// One cluster of S CTAs per unit, units taken round-robin by the clusters.  mode: FUSED_FULL / FUSED_KEYS_ONLY /
// FUSED_GIVEN_THRESH, FUSED_MINMAX or-ed in (wc_fused.h), with the meaning they have for k_fused_compress.
__global__ void __launch_bounds__(XS_NT, 1)
k_xs_compress(const UnitDev* __restrict__ units, UnitState* __restrict__ states, const int* __restrict__ unit_list,
              int n_list, double one_minus_keep, const u64* __restrict__ global_key, int mode_flags) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* const C      = reinterpret_cast<float*>(smem);
    int* const   s_cnt  = reinterpret_cast<int*>(smem + XS_OFF_CNT);
    int* const   s_last = reinterpret_cast<int*>(smem + XS_OFF_LAST);
    int* const   s_base = reinterpret_cast<int*>(smem + XS_OFF_BASE);
    int* const   s_prev = reinterpret_cast<int*>(smem + XS_OFF_PREV);
    u64* const   s_red  = reinterpret_cast<u64*>(smem + XS_OFF_RED);
    u64* const   s_x1   = reinterpret_cast<u64*>(smem + XS_OFF_X1);
    int* const   s_x2   = reinterpret_cast<int*>(smem + XS_OFF_X2);

    cg::cluster_group cluster = cg::this_cluster();
    const int S    = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int cid  = (int)blockIdx.x / S, ncl = (int)gridDim.x / S;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t lt = lanemask_lt();
    const int  mode = mode_flags & 15;
    const bool mm   = (mode_flags & FUSED_MINMAX) != 0;

    int it = 0;
    for (int ui = cid; ui < n_list; ui += ncl, ++it) {
        const int     uid = unit_list[ui];
        const UnitDev u   = units[uid];
        const int     par = it & 1;
        XGeom g;
        g.init(u.nx, u.ny, u.nz, S, rank);

        // ---------------- phase A ----------------
        u64   key = 0ull;
        float vmn = __int_as_float(0x7f800000), vmx = __int_as_float(0xff800000);
        if (u.dtype == WC_F64) {
            if (mm) xs_phase_a<double, true>(g, u.in, C, key, vmn, vmx);
            else    xs_phase_a<double, false>(g, u.in, C, key, vmn, vmx);
        } else {
            if (mm) xs_phase_a<float, true>(g, u.in, C, key, vmn, vmx);
            else    xs_phase_a<float, false>(g, u.in, C, key, vmn, vmx);
        }
        key = warp_max_u64(key);
        if (mm) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                vmn = fminf(vmn, __shfl_xor_sync(0xffffffffu, vmn, o));
                vmx = fmaxf(vmx, __shfl_xor_sync(0xffffffffu, vmx, o));
            }
        }
        if (lane == 0) {
            s_red[warp]      = key;
            s_red[32 + warp] = ((u64)__float_as_uint(vmn) << 32) | (u64)__float_as_uint(vmx);
        }
        __syncthreads();                       // C complete, warp keys visible

        // ---------------- phase B: the unit's arg-max key over the cluster, the threshold ----------------
        if (warp == 0) {
            u64 k = lane < XS_NW ? s_red[lane] : 0ull;
            k = warp_max_u64(k);
            if (lane == 0) {
                s_x1[par * 2]     = k;
                // the coefficient at f = 0 sits in rank 0's first local plane (the trailing plane when X == 1)
                s_x1[par * 2 + 1] = (rank == 0 && isnan(C[0])) ? 1ull : 0ull;
            }
            __syncwarp();      // explicit reconvergence behind one-lane regions that precede warp collectives
            if (mm) {
                const u64 x = lane < XS_NW ? s_red[32 + lane] : (((u64)0x7f800000u << 32) | 0xff800000u);
                float a = __uint_as_float((uint32_t)(x >> 32)), b = __uint_as_float((uint32_t)x);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
                    b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
                }
                // same encoding as k_fused_compress: atomicMax over zeroed words, decoded by wc_plan_unit_stats
                if (lane == 0 && a <= b) {
                    atomicMax(reinterpret_cast<unsigned int*>(&states[uid].vmin), ~xs_order_code(a));
                    atomicMax(reinterpret_cast<unsigned int*>(&states[uid].vmax), xs_order_code(b));
                }
            }
        }
        cluster.sync();                        // exchange 1 (for S == 1 a CTA barrier)
        u64  ukey = 0ull;
        bool first_nan = false;
        for (int r = 0; r < S; ++r) {
            const u64* px = cluster.map_shared_rank(s_x1 + par * 2, r);
            ukey = max_u64(ukey, px[0]);
            first_nan = first_nan || px[1] != 0ull;
        }
        float tf;
        if (mode == FUSED_GIVEN_THRESH) {
            const u64 gk = *global_key;
            tf = threshold_float(gk & ~(1ull << 63), (gk >> 63) != 0, one_minus_keep);
        } else {
            tf = threshold_float(ukey, first_nan, one_minus_keep);
        }
        if (tid == 0 && rank == 0) {
            states[uid].key      = ukey;
            states[uid].flags    = first_nan ? UNIT_FLAG_NAN0 : 0;
            states[uid].thresh_f = tf;
        }
        if (mode == FUSED_KEYS_ONLY) {
            __syncthreads();                   // C and s_red are rewritten by the next unit
            continue;
        }

        // ---------------- phase C1: kept count and last kept coefficient of every local plane ----------------
#pragma unroll 1
        for (int p = warp; p < g.npl; p += XS_NW) {
            const float* cs = C + p * g.PS;
            int cnt = 0, last = -1;
            for (int w0 = 0; w0 < g.YZ; w0 += 32) {
                const int  w  = w0 + lane;
                const bool kf = w < g.YZ && keep_coef(cs[w], tf);
                const uint32_t bal = __ballot_sync(0xffffffffu, kf);
                if (bal) { cnt += __popc(bal); last = w0 + 31 - __clz(bal); }
            }
            if (lane == 0) {
                s_cnt[p]  = cnt;
                s_last[p] = last >= 0 ? g.gplane(p) * g.YZ + last : -1;
            }
            __syncwarp();
        }
        __syncthreads();
        // scan inside either range (low planes, high planes + trailing plane): pairs in front of every plane, last kept
        // flat index in front of it; the totals go to the cluster
        if (warp == 0) {
            int tot[2], lastf[2];
#pragma unroll
            for (int grp = 0; grp < 2; ++grp) {
                const int lo = grp ? g.nl : 0, hi = grp ? g.npl : g.nl;
                int carry = 0, cmax = -1;
                for (int base = lo; base < hi; base += 32) {
                    const int p = base + lane;
                    const int c = p < hi ? s_cnt[p] : 0, l = p < hi ? s_last[p] : -1;
                    int isum = c, imax = l;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int ps = __shfl_up_sync(0xffffffffu, isum, o), pm = __shfl_up_sync(0xffffffffu, imax, o);
                        if (lane >= o) { isum += ps; imax = max(imax, pm); }
                    }
                    int emax = __shfl_up_sync(0xffffffffu, imax, 1);
                    if (lane == 0) emax = -1;
                    if (p < hi) { s_base[p] = carry + isum - c; s_prev[p] = max(cmax, emax); }
                    __syncwarp();
                    carry += __shfl_sync(0xffffffffu, isum, 31);
                    cmax = max(cmax, __shfl_sync(0xffffffffu, imax, 31));
                }
                tot[grp] = carry; lastf[grp] = cmax;
            }
            if (lane == 0) {
                int* x2 = s_x2 + par * 4;
                x2[0] = tot[0]; x2[1] = lastf[0]; x2[2] = tot[1]; x2[3] = lastf[1];
            }
        }
        cluster.sync();                        // exchange 2
        // position of this CTA's two ranges in the unit's pair list: low ranges of ranks 0 .. S-1, then the high ranges
        int base_lo = 0, prev_lo = -1, base_hi = 0, prev_hi = -1, lo_all = 0, lastlo_all = -1, K = 0, last_all = -1;
        for (int r = 0; r < S; ++r) {
            const int* x = cluster.map_shared_rank(s_x2 + par * 4, r);
            const int c0 = x[0], l0 = x[1], c1 = x[2], l1 = x[3];
            if (r < rank) { base_lo += c0; prev_lo = max(prev_lo, l0); base_hi += c1; prev_hi = max(prev_hi, l1); }
            lo_all += c0; lastlo_all = max(lastlo_all, l0);
            K += c0 + c1; last_all = max(last_all, max(l0, l1));
        }
        base_hi += lo_all;
        prev_hi = max(prev_hi, lastlo_all);
        int2* const tab = reinterpret_cast<int2*>(u.coef);     // decode-side plane table (cluster classes), or null
        if (tid == 0 && rank == 0) {
            const float M = fabsf(key_value(ukey));
            states[uid].npairs = K;
            states[uid].flags  = (first_nan ? UNIT_FLAG_NAN0 : 0) | ((K > 0 && unit_need32(M, tf)) ? UNIT_FLAG_NEED32 : 0);
            if (tab) tab[g.X] = make_int2(K, last_all);
        }

        // ---------------- phase C2: emit (run, value) pairs, a warp per plane ----------------
        int2* const out = reinterpret_cast<int2*>(u.out);
#pragma unroll 1
        for (int p = warp; p < g.npl; p += XS_NW) {
            const bool hi_grp = p >= g.nl;
            int pos  = (hi_grp ? base_hi : base_lo) + s_base[p];
            int prev = max(hi_grp ? prev_hi : prev_lo, s_prev[p]);
            const int fstart = g.gplane(p) * g.YZ;
            if (tab && lane == 0) tab[g.gplane(p)] = make_int2(pos, prev);
            __syncwarp();
            if (s_cnt[p] == 0) continue;
            const float* cs = C + p * g.PS;
            for (int w0 = 0; w0 < g.YZ; w0 += 32) {
                const int   w  = w0 + lane;
                const float c  = w < g.YZ ? cs[w] : 0.f;
                const bool  kf = w < g.YZ && keep_coef(c, tf);
                const uint32_t bal = __ballot_sync(0xffffffffu, kf);
                if (bal) {
                    const uint32_t lower = bal & lt;
                    // flat index of the previous kept coefficient: a lower lane of this group, or the carried one
                    const int pf = lower ? fstart + w0 + 31 - __clz(lower) : prev;
                    if (kf) out[pos + __popc(lower)] = make_int2(fstart + w - pf - 1, __float_as_int(c));
                    pos += __popc(bal);
                    prev = fstart + w0 + 31 - __clz(bal);
                }
            }
        }
        __syncthreads();                       // C and the plane arrays are rewritten by the next unit
    }
    cluster.sync();                            // no CTA may exit while a peer can still read its shared memory
}

// ---- decompress ----------------------------------------------------------------------------------------------------
// Block-wide exclusive prefix of run + 1 over one tile of XS_NT * XS_PPT pairs (saturating at 2^30, so corrupt streams
// cannot wrap; negative runs are flagged, count as 0 and are skipped by the caller).  wt: 32 words per tile parity.
__device__ __forceinline__ uint32_t xs_sat_add(uint32_t a, uint32_t b) {
    const uint32_t s = a + b;
    return s > 0x40000000u ? 0x40000000u : s;
}
__device__ __forceinline__ uint32_t xs_tile_scan(const int2 (&pr)[XS_PPT], int nvalid, uint32_t* wt, bool& bad,
                                                 uint32_t& ttot) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < XS_PPT; ++j) {
        const bool in = j < nvalid;
        if (in && pr[j].x < 0) bad = true;
        s = xs_sat_add(s, (in && pr[j].x >= 0) ? (uint32_t)pr[j].x + 1u : 0u);
    }
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = xs_sat_add(inc, v);
    }
    if (lane == 31) wt[warp] = inc;
    __syncthreads();
    uint32_t winc = lane < XS_NW ? wt[lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc = xs_sat_add(winc, v);
    }
    ttot = __shfl_sync(0xffffffffu, winc, 31);
    uint32_t wpre = __shfl_sync(0xffffffffu, winc, (warp + 31) & 31);
    if (warp == 0) wpre = 0;
    uint32_t exl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exl = 0;
    return xs_sat_add(wpre, exl);
}

// rle_decode (src/decompressor.cpp:14-30) of the pairs [pb, pe) whose first run starts at flat index `cur`: every pair
// that lands in [f0, f1) goes to C (plane (f - f0) / YZ + pbase); pairs past f1 — and with them every later one, flat
// indices only grow — are dropped, which for f1 <= total is the reference's `if (idx < total)`.
__device__ __forceinline__ void xs_decode_range(const int2* __restrict__ pairs, int pb, int pe, uint32_t cur, uint32_t f0,
                                                uint32_t f1, int pbase, const XGeom& g, float* C, uint32_t* s_wt, int& tile,
                                                bool& bad) {
#pragma unroll 1
    for (int p0 = pb; p0 < pe; p0 += XS_NT * XS_PPT, ++tile) {
        const int p = p0 + (int)threadIdx.x * XS_PPT;
        int2 pr[XS_PPT];
#pragma unroll
        for (int j = 0; j < XS_PPT; ++j) pr[j] = (p + j < pe) ? __ldg(pairs + p + j) : make_int2(0, 0);
        uint32_t ttot;
        uint32_t rp = xs_sat_add(cur, xs_tile_scan(pr, pe - p, s_wt + (tile & 1) * 32, bad, ttot));
#pragma unroll
        for (int j = 0; j < XS_PPT; ++j) {
            if (p + j < pe && pr[j].x >= 0) {
                const uint32_t f = xs_sat_add(rp, (uint32_t)pr[j].x);
                if (f >= f0 && f < f1) {
                    const uint32_t d = f - f0, pl = d / (uint32_t)g.YZ;
                    C[(pbase + (int)pl) * g.PS + (int)(d - pl * (uint32_t)g.YZ)] = __int_as_float(pr[j].y);
                }
                rp = xs_sat_add(rp, (uint32_t)pr[j].x + 1u);
            }
        }
        cur = xs_sat_add(cur, ttot);
        __syncwarp();
    }
}

template <class T>
__device__ __forceinline__ void xs_inverse_store(const XGeom& g, const float* C, T* __restrict__ out) {
    const int tid = threadIdx.x;
    // full 2 x 2 x 2 blocks: X, then Y, then Z (src/decompressor.cpp:90-156)
    const int nblk = g.nl * g.hy * g.hz;
#pragma unroll 1
    for (int q = tid; q < nblk; q += XS_NT) {
        const int al = q % g.nl, t = q / g.nl, b = t % g.hy, c = t / g.hy;
        float v[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const int sx = o & 1, sy = (o >> 1) & 1, sz = o >> 2;
            v[o] = C[(sx ? g.nl + al : al) * g.PS + (sy ? g.hy + b : b) * g.Z + (sz ? g.hz + c : c)];
        }
        haar_block_inverse_full(v);
        const int x0 = 2 * (g.a0 + al);
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const int xi = o & 1, yi = (o >> 1) & 1, zi = o >> 2;
            out[((size_t)(2 * c + zi) * g.Y + (size_t)(2 * b + yi)) * g.X + (size_t)(x0 + xi)] = (T)v[o];
        }
    }
    // trailing cells of odd axes: the inverse leaves them at +0
    const int nxs = 2 * g.nl + g.own1;                    // x positions of this slab (incl. the trailing plane)
    auto xpos = [&](int i) { return i < 2 * g.nl ? 2 * g.a0 + i : g.X - 1; };
    if (g.oz) {
        for (int q = tid; q < nxs * g.Y; q += XS_NT) {
            const int i = q % nxs, y = q / nxs;
            out[((size_t)(g.Z - 1) * g.Y + (size_t)y) * g.X + (size_t)xpos(i)] = (T)0;
        }
    }
    if (g.oy) {
        const int nz2 = g.Z - g.oz;
        for (int q = tid; q < nxs * nz2; q += XS_NT) {
            const int i = q % nxs, z = q / nxs;
            out[((size_t)z * g.Y + (size_t)(g.Y - 1)) * g.X + (size_t)xpos(i)] = (T)0;
        }
    }
    if (g.own1) {
        const int ny2 = g.Y - g.oy, nz2 = g.Z - g.oz;
        for (int q = tid; q < ny2 * nz2; q += XS_NT) {
            const int y = q % ny2, z = q / ny2;
            out[((size_t)z * g.Y + (size_t)y) * g.X + (size_t)(g.X - 1)] = (T)0;
        }
    }
}

// Work item = (unit, x-slab r of S); one CTA per item, items handed out through a global counter.
__global__ void __launch_bounds__(XS_NT, 1)
k_xs_decompress(const DecUnitDev* __restrict__ dec, const InvUnitDev* __restrict__ inv, const int* __restrict__ unit_list,
                int n_list, int S, int* __restrict__ err, int* __restrict__ work_counter) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* const    C      = reinterpret_cast<float*>(smem);
    uint32_t* const s_wt   = reinterpret_cast<uint32_t*>(smem + XS_OFF_RED);      // [2][32]
    int* const      s_item = reinterpret_cast<int*>(smem + XS_OFF_MISC);
    const int tid = threadIdx.x;
    const int n_items = n_list * S;
    bool bad = false;
    for (int k = 0;; ++k) {
        if (tid == 0) *s_item = work_counter ? atomicAdd(work_counter, 1) : (int)blockIdx.x + k * (int)gridDim.x;
        __syncthreads();
        const int item = *s_item;
        if (item >= n_items) break;
        const int uid = unit_list[item / S], rank = item % S;
        const DecUnitDev du = dec[uid];
        const InvUnitDev iu = inv[uid];
        XGeom g;
        g.init(iu.nx, iu.ny, iu.nz, S, rank);
        if (g.npl > 0) {
            // the slab's low and high planes start out zero (src/decompressor.cpp:16)
            for (int i = tid; i < 2 * g.nl * g.PS; i += XS_NT) C[i] = 0.f;
            __syncthreads();
            const int2* pairs = reinterpret_cast<const int2*>(du.pairs);
            int K = du.npairs_dev ? *du.npairs_dev : du.npairs;
            K = max(0, min(K, du.total));
            int tile = 0;
            if (g.nl > 0) {
                if (S == 1) {
                    xs_decode_range(pairs, 0, K, 0u, 0u, (uint32_t)(2 * g.hx * g.YZ), 0, g, C, s_wt, tile, bad);
                } else {
                    const int2* tab = reinterpret_cast<const int2*>(du.coef);
                    const int2 l0 = tab[g.a0], l1 = tab[g.a0 + g.nl];
                    const int2 h0 = tab[g.hx + g.a0], h1 = tab[g.hx + g.a0 + g.nl];
                    xs_decode_range(pairs, max(0, l0.x), min(K, l1.x), (uint32_t)(l0.y + 1), (uint32_t)(g.a0 * g.YZ),
                                    (uint32_t)((g.a0 + g.nl) * g.YZ), 0, g, C, s_wt, tile, bad);
                    xs_decode_range(pairs, max(0, h0.x), min(K, h1.x), (uint32_t)(h0.y + 1), (uint32_t)((g.hx + g.a0) * g.YZ),
                                    (uint32_t)((g.hx + g.a0 + g.nl) * g.YZ), g.nl, g, C, s_wt, tile, bad);
                }
            }
            __syncthreads();
            if (iu.dtype == WC_F64) xs_inverse_store<double>(g, C, static_cast<double*>(iu.out));
            else                    xs_inverse_store<float>(g, C, static_cast<float*>(iu.out));
        }
        __syncthreads();                       // C and s_item are rewritten by the next item
    }
    if (bad) atomicOr(err, 1);
}

// ---- launchers -----------------------------------------------------------------------------------------------------
int xs_class_slabs(int fused_cls) {
    switch (fused_cls) {
    case FUSED_CLS_XS1: return 1;
    case FUSED_CLS_XS2: return 2;
    case FUSED_CLS_XS4: return 4;
    case FUSED_CLS_XS8: return 8;
    }
    return 0;
}
int xs_class_of(int nx, int ny, int nz) {
    switch (xs_slabs(nx, ny, nz)) {
    case 1: return FUSED_CLS_XS1;
    case 2: return FUSED_CLS_XS2;
    case 4: return FUSED_CLS_XS4;
    case 8: return FUSED_CLS_XS8;
    }
    return FUSED_CLS_NONE;
}

cudaError_t launch_xs_compress(int fused_cls, int mode, const UnitDev* units, UnitState* states, const int* unit_list,
                               int n_list, double one_minus_keep, const u64* global_key, int sm_count, cudaStream_t st,
                               LaunchStats* ls) {
    const int S = xs_class_slabs(fused_cls);
    if (S == 0) return cudaErrorInvalidValue;
    if (n_list <= 0) return cudaSuccess;
    const int kid = S == 1 ? KID_XS_C1 : S == 2 ? KID_XS_C2 : S == 4 ? KID_XS_C4 : KID_XS_C8;
    cudaError_t e = cudaFuncSetAttribute(k_xs_compress, cudaFuncAttributeMaxDynamicSharedMemorySize, XS_SMEM);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim         = dim3(XS_NT);
    cfg.dynamicSmemBytes = XS_SMEM;
    cfg.stream           = st;
    cudaLaunchAttribute attr[1];
    attr[0].id               = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs    = attr;
    cfg.numAttrs = 1;
    int& max_clusters = ls->occ[kid];          // resident clusters on this ctx's device
    if (max_clusters == 0) {
        cfg.gridDim = dim3(S * sm_count);
        int nc = 0;
        e = cudaOccupancyMaxActiveClusters(&nc, k_xs_compress, &cfg);
        if (e != cudaSuccess) return e;
        if (nc < 1) return cudaErrorLaunchOutOfResources;
        max_clusters = nc;
    }
    const int nc = max_clusters < n_list ? max_clusters : n_list;
    cfg.gridDim = dim3(nc * S);
    ls->begin(kid, st);
    e = cudaLaunchKernelEx(&cfg, k_xs_compress, units, states, unit_list, n_list, one_minus_keep, global_key, mode);
    ls->end(st);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

cudaError_t launch_xs_decompress(int fused_cls, const DecUnitDev* dec, const InvUnitDev* inv, const int* unit_list,
                                 int n_list, int* err, int sm_count, cudaStream_t st, LaunchStats* ls, int* work_counter) {
    const int S = xs_class_slabs(fused_cls);
    if (S == 0) return cudaErrorInvalidValue;
    if (n_list <= 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_xs_decompress, cudaFuncAttributeMaxDynamicSharedMemorySize, XS_SMEM);
    if (e != cudaSuccess) return e;
    const long long items = (long long)n_list * S;
    const int nc = (int)(items < sm_count ? items : sm_count);
    ls->begin(KID_XS_D, st);
    k_xs_decompress<<<nc, XS_NT, XS_SMEM, st>>>(dec, inv, unit_list, n_list, S, err, work_counter);
    ls->end(st);
    return cudaGetLastError();
}

} // namespace wc
