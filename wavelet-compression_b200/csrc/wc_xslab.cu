// Fused on-chip kernels for boxes of ANY shape (odd dimensions, nz % 4 != 0, rows that are not 16-byte multiples,
// unaligned pointers): the x-slab classes FUSED_CLS_XS1S / XS1 / XS2 / XS4 / XS8.
//
// The y-slab kernels of wc_fused.cu owe their speed to 16-byte vector accesses, and with them to even dimensions and
// nz % 4 == 0.  Everything else used to take the generic multi-kernel path (coefficient scratch in HBM: 20N + 8K bytes of
// traffic).  These kernels keep such a unit on chip as well: HBM sees the input once and the pairs once.
//
// Decomposition by X-SLABS.  CTA r of S takes the block columns a in [r*na, (r+1)*na) (na = ceil(hx / S)); the trailing
// plane x = X-1 of an odd X (which passes through the x stage, src/compressor.cpp:153-175) goes to the last CTA.  The
// coefficients of a slab are whole PLANES i' of the flat order f = (i'*Y + j')*Z + k' (src/compressor.cpp:178-181): the
// low planes [a0, a1) and the high planes [hx+a0, hx+a1) (+ the plane X-1) — two contiguous flat ranges.  So
//   compress:   a cluster of S CTAs exchanges, through distributed shared memory, the arg-max (max c, max -c, NaN at f = 0)
//               and four words (kept count and last kept flat index of either range) per CTA — nothing else is needed to
//               place a CTA's pairs in the unit's ordered pair list and to know the zero run in front of its first pair;
//               behind its own loads a CTA prefetches its share of the cluster's NEXT unit into L2;
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
// multiples of the blocking factor); the point is that they no longer fall off the fused path.  Measurements, the ncu
// tables and what was tried: profiles/r02_xslab.md; design: DESIGN.md §4.6.
#include <cooperative_groups.h>

#include "wc_common.cuh"
#include "wc_fused.h"

namespace cg = cooperative_groups;

namespace wc {

// Two configurations: 512 threads and a coefficient array of 220 KB (one CTA per SM), and — for the units that fit 33 KB
// (FUSED_CLS_XS1S: 16^3-sized boxes and smaller) — 128 threads with five CTAs per SM, which overlap each other's barriers
// and load latencies the way the small-unit variants of the y-slab kernels do.
constexpr int XS_CWORDS   = 56320;             // words of C per CTA, large configuration
constexpr int XS_CWORDS_S = 8448;              // ... small configuration
constexpr int XS_SEG      = 512;               // coefficients per segment: the unit of work of the packing phases
constexpr int XS_MAXSEG   = 384;               // local segments: npl * ceil(YZ / 512) <= CWORDS / 512 + 257 (nx <= 256)
constexpr int XS_PPT      = 8;                 // pairs per thread per tile of the decoder's block scan
// shared memory layout (bytes) behind the CW words of C
template <int CW>
struct XSmem {
    static constexpr int CNT   = CW * 4;                      // int[XS_MAXSEG]  kept coefficients per local segment
    static constexpr int LAST  = CNT + XS_MAXSEG * 4;         // int[XS_MAXSEG]  flat index of the last kept one, or -1
    static constexpr int BASE  = LAST + XS_MAXSEG * 4;        // int[XS_MAXSEG]  pairs of the same range in front of it
    static constexpr int PREV  = BASE + XS_MAXSEG * 4;        // int[XS_MAXSEG]  last kept flat index of the range before it
    static constexpr int RED   = PREV + XS_MAXSEG * 4;        // u64[64]         block reductions / scan scratch
    static constexpr int X1    = RED + 64 * 8;                // u64[2][2]       exchange 1: arg-max key, NaN-at-f=0
    static constexpr int X2    = X1 + 32;                     // int[2][4]       exchange 2: count / last of either range
    static constexpr int X3    = X2 + 32;                     // u64[2]          exchange 3 (+M / -M ties): lowest flat index
    static constexpr int MISC  = X3 + 16;                     // int[8]
    static constexpr int TOTAL = MISC + 32;
};
static_assert(XSmem<XS_CWORDS>::TOTAL <= 232448, "exceeds the 227 KB per-CTA shared memory of sm_100");

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
static bool xs_fits_small(int nx, int ny, int nz) {
    if (nx < 1 || ny < 1 || nz < 1 || nx > 256) return false;
    const long long yz = (long long)ny * nz;
    return yz <= XS_CWORDS_S && (long long)nx * xs_plane_stride((int)yz) <= XS_CWORDS_S;
}

struct XGeom {
    int X, Y, Z, hx, hy, hz, ox, oy, oz;
    int YZ, PS;
    int a0, nl, own1;      // this slab: first block column, block columns, owns the trailing plane x = X-1
    int npl;               // local planes: 2 * nl + own1
    int spp;               // segments per plane: ceil(YZ / XS_SEG)
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
        spp = (YZ + XS_SEG - 1) / XS_SEG;
    }
    // global plane i' of local plane p (low planes, high planes, then the trailing plane)
    __device__ __forceinline__ int gplane(int p) const { return p < nl ? a0 + p : (p < 2 * nl ? hx + a0 + (p - nl) : X - 1); }
};

// q / d by a magic multiply: exact while q * d < 2^32 (block and cell counts here stay below 2^17)
__device__ __forceinline__ uint32_t xs_magic(uint32_t d) { return d <= 1 ? 0u : (0xffffffffu / d) + 1u; }
__device__ __forceinline__ uint32_t xs_div(uint32_t q, uint32_t m) { return m ? __umulhi(q, m) : q; }

// Phase A of one slab: every thread takes generalized blocks (al fastest, then b, then c): up to 2 x 2 x 2 cells, a single
// cell wide along an axis whose trailing element it holds.  Coefficients go to C; the arg-max is tracked as two running float
// maxima (max c, max -c; fmaxf skips NaNs as std::max_element's comparison does, src/compressor.cpp:212-215) — only an exact
// +M == -M tie needs the flat index, and that is searched for in C afterwards (k_xs_compress, phase B).
// Full blocks (all but the trailing planes) take a path without predicates: one 64-bit address per block, 32-bit offsets.
template <class T> __device__ __forceinline__ float xs_ld(const T* p);
template <> __device__ __forceinline__ float xs_ld<double>(const double* p) { return __double2float_rn(__ldg(p)); }   // src/preprocess.cpp:78
template <> __device__ __forceinline__ float xs_ld<float>(const float* p) { return __ldg(p); }
// the (x, x+1) outputs of a full block: one vector store where the pair happens to be aligned (even X and an aligned box:
// every row; odd X: every other row), two element stores otherwise: decode +3..6 %.  (The same for the loads of phase A
// measured 1..7 % SLOWER — the alignment test sits in front of the loads it guards — and was dropped.)
__device__ __forceinline__ void xs_st2(double* p, float a, float b) {
    if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) *reinterpret_cast<double2*>(p) = make_double2((double)a, (double)b);
    else { p[0] = (double)a; p[1] = (double)b; }
}
__device__ __forceinline__ void xs_st2(float* p, float a, float b) {
    if ((reinterpret_cast<uintptr_t>(p) & 7u) == 0) *reinterpret_cast<float2*>(p) = make_float2(a, b);
    else { p[0] = a; p[1] = b; }
}

template <class T, bool MM, int NT>
__device__ __forceinline__ void xs_phase_a(const XGeom& g, const void* in, float* C, float& bp, float& bn, float& vmn,
                                           float& vmx) {
    const int nbx = g.nl + g.own1, nby = g.hy + g.oy, nbz = g.hz + g.oz;
    const int nblk = nbx * nby * nbz;
    const uint32_t m_bx = xs_magic(nbx), m_by = xs_magic(nby);
    const T* const src = static_cast<const T*>(in);
    const int sy = g.X, sz = g.X * g.Y;                                  // element strides of the box
    const int o_hx = g.nl * g.PS, o_hy = g.hy * g.Z, o_hz = g.hz;        // offsets of the high bands in C
#pragma unroll 1
    for (int q = threadIdx.x; q < nblk; q += NT) {
        const int t = (int)xs_div(q, m_bx), al = q - t * nbx, c = (int)xs_div(t, m_by), b = t - c * nby;
        const bool wx = al < g.nl, wy = b < g.hy, wz = c < g.hz;
        float v[8];
        if (wx && wy && wz) {
            const T* p = src + ((size_t)(2 * c) * sz + (size_t)(2 * b) * sy + (size_t)(2 * (g.a0 + al)));
#pragma unroll
            for (int o = 0; o < 8; ++o) v[o] = xs_ld<T>(p + ((o >> 2) * sz + ((o >> 1) & 1) * sy + (o & 1)));
            if (MM) {
#pragma unroll
                for (int o = 0; o < 8; ++o) { vmn = fminf(vmn, v[o]); vmx = fmaxf(vmx, v[o]); }
            }
            haar_block_forward_full(v);
            float* cd = C + al * g.PS + b * g.Z + c;
#pragma unroll
            for (int o = 0; o < 8; o += 2) {
                cd[((o >> 1) & 1) * o_hy + (o >> 2) * o_hz]        = v[o];
                cd[o_hx + ((o >> 1) & 1) * o_hy + (o >> 2) * o_hz] = v[o + 1];
                bp = fmaxf(fmaxf(bp, v[o]), v[o + 1]);
                bn = fmaxf(fmaxf(bn, -v[o]), -v[o + 1]);
            }
        } else {
            const int x0 = wx ? 2 * (g.a0 + al) : g.X - 1, y0 = wy ? 2 * b : g.Y - 1, z0 = wz ? 2 * c : g.Z - 1;
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const int xi = o & 1, yi = (o >> 1) & 1, zi = o >> 2;
                const bool valid = (xi == 0 || wx) && (yi == 0 || wy) && (zi == 0 || wz);
                v[o] = 0.f;
                if (valid) {
                    v[o] = xs_ld<T>(src + ((size_t)(z0 + zi) * sz + (size_t)(y0 + yi) * sy + (size_t)(x0 + xi)));
                    if (MM) { vmn = fminf(vmn, v[o]); vmx = fmaxf(vmx, v[o]); }
                }
            }
            haar_block_forward(v, wx, wy, wz);
            const int pl0 = wx ? al : 2 * g.nl, pl1 = g.nl + al;           // local planes of the low / high x band
            const int j0 = wy ? b : g.Y - 1, j1 = g.hy + b, k0 = wz ? c : g.Z - 1, k1 = g.hz + c;
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const int sx = o & 1, sy2 = (o >> 1) & 1, sz2 = o >> 2;
                const bool valid = (sx == 0 || wx) && (sy2 == 0 || wy) && (sz2 == 0 || wz);
                if (valid) {
                    C[(sx ? pl1 : pl0) * g.PS + (sy2 ? j1 : j0) * g.Z + (sz2 ? k1 : k0)] = v[o];
                    bp = fmaxf(bp, v[o]);
                    bn = fmaxf(bn, -v[o]);
                }
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
// The packing phases work on SEGMENTS of at most XS_SEG consecutive coefficients of one plane (a warp per segment at a
// time): with whole planes as the unit of work a 63^3 slab had 9 work items of 3969 coefficients for 16 warps.
template <int NT, int CW, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_xs_compress(const UnitDev* __restrict__ units, UnitState* __restrict__ states, const int* __restrict__ unit_list,
              int n_list, double one_minus_keep, const u64* __restrict__ global_key, int mode_flags) {
    typedef XSmem<CW> SM;
    constexpr int NW = NT / 32;
    extern __shared__ __align__(128) unsigned char smem[];
    float* const C      = reinterpret_cast<float*>(smem);
    int* const   s_cnt  = reinterpret_cast<int*>(smem + SM::CNT);
    int* const   s_last = reinterpret_cast<int*>(smem + SM::LAST);
    int* const   s_base = reinterpret_cast<int*>(smem + SM::BASE);
    int* const   s_prev = reinterpret_cast<int*>(smem + SM::PREV);
    u64* const   s_red  = reinterpret_cast<u64*>(smem + SM::RED);
    u64* const   s_x1   = reinterpret_cast<u64*>(smem + SM::X1);
    int* const   s_x2   = reinterpret_cast<int*>(smem + SM::X2);
    u64* const   s_x3   = reinterpret_cast<u64*>(smem + SM::X3);      // [2]  exchange 3 (ties only)
    int* const   s_seed = reinterpret_cast<int*>(smem + SM::MISC);    // [8]  reduced exchange values of this CTA

    cg::cluster_group cluster = cg::this_cluster();
    const int S    = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int cid  = (int)blockIdx.x / S, ncl = (int)gridDim.x / S;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t lt = lanemask_lt();
    const int  mode = mode_flags & 15;
    const bool mm   = (mode_flags & FUSED_MINMAX) != 0;
    // cluster-wide barrier with release / acquire of shared memory; a plain CTA barrier when the cluster is one CTA
    auto xsync = [&]() { if (S == 1) __syncthreads(); else cluster.sync(); };

    int it = 0;
    for (int ui = cid; ui < n_list; ui += ncl, ++it) {
        const int     uid = unit_list[ui];
        const UnitDev u   = units[uid];
        const int     par = it & 1;
        XGeom g;
        g.init(u.nx, u.ny, u.nz, S, rank);

        // ---------------- phase A ----------------
        float bp = 0.f, bn = 0.f;              // running max of +c and of -c
        float vmn = __int_as_float(0x7f800000), vmx = __int_as_float(0xff800000);
        if (u.dtype == WC_F64) {
            if (mm) xs_phase_a<double, true, NT>(g, u.in, C, bp, bn, vmn, vmx);
            else    xs_phase_a<double, false, NT>(g, u.in, C, bp, bn, vmn, vmx);
        } else {
            if (mm) xs_phase_a<float, true, NT>(g, u.in, C, bp, bn, vmn, vmx);
            else    xs_phase_a<float, false, NT>(g, u.in, C, bp, bn, vmn, vmx);
        }
        // L2 prefetch of this cluster's NEXT unit (every CTA takes 1/S of its 128-byte lines), issued behind this unit's own
        // loads: the packing phases below read nothing from HBM, so the next phase A finds its input in L2
        if (ui + ncl < n_list) {
            const UnitDev* un = units + unit_list[ui + ncl];
            const char*  nb  = static_cast<const char*>(un->in);
            const size_t nby = (size_t)un->n * (un->dtype == WC_F64 ? 8 : 4);
            const int lines = (int)((nby + 127) >> 7), per = (lines + S - 1) / S;
            const int l0 = rank * per, l1 = min(lines, l0 + per);
            for (int i = l0 + tid; i < l1; i += NT)
                asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(nb + (size_t)i * 128));
        }
        // both maxima are >= 0 (a -0 cannot win against the initial +0; masked all the same): the halves reduce separately
        u64 key = ((u64)(__float_as_uint(bp) & 0x7fffffffu) << 32) | (u64)(__float_as_uint(bn) & 0x7fffffffu);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const u64 x = __shfl_xor_sync(0xffffffffu, key, o);
            const uint32_t hi = max((uint32_t)(key >> 32), (uint32_t)(x >> 32)), lo = max((uint32_t)key, (uint32_t)x);
            key = ((u64)hi << 32) | lo;        // non-negative floats order like their bit patterns
        }
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
            u64 k = lane < NW ? s_red[lane] : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const u64 x = __shfl_xor_sync(0xffffffffu, k, o);
                const uint32_t hi = max((uint32_t)(k >> 32), (uint32_t)(x >> 32)), lo = max((uint32_t)k, (uint32_t)x);
                k = ((u64)hi << 32) | lo;
            }
            if (lane == 0) {
                s_x1[par * 2]     = k;
                // the coefficient at f = 0 sits in rank 0's first local plane (the trailing plane when X == 1)
                s_x1[par * 2 + 1] = (rank == 0 && isnan(C[0])) ? 1ull : 0ull;
            }
            __syncwarp();      // explicit reconvergence behind one-lane regions that precede warp collectives
            if (mm) {
                const u64 x = lane < NW ? s_red[32 + lane] : (((u64)0x7f800000u << 32) | 0xff800000u);
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
        xsync();                               // exchange 1
        u64  ukey = 0ull;
        bool first_nan = false;
        {
            // one warp reads the peers (lane r <- rank r) and leaves the reduced values in this CTA's shared memory: with
            // every thread reading every peer a 63^3 slab spent more time in remote loads than in its cluster barriers
            if (warp == 0) {
                u64 x = 0ull, f = 0ull;
                if (lane < S) {
                    const u64* px = cluster.map_shared_rank(s_x1 + par * 2, lane);
                    x = px[0]; f = px[1];
                }
                const uint32_t rp = __reduce_max_sync(0xffffffffu, (uint32_t)(x >> 32));
                const uint32_t rn = __reduce_max_sync(0xffffffffu, (uint32_t)x);
                const bool     fn = __any_sync(0xffffffffu, f != 0ull);
                if (lane == 0) { s_seed[0] = (int)rp; s_seed[1] = (int)rn; s_seed[2] = fn ? 1 : 0; }
            }
            __syncthreads();
            const uint32_t mp = (uint32_t)s_seed[0], mn = (uint32_t)s_seed[1];
            first_nan = s_seed[2] != 0;
            const uint32_t mb = max(mp, mn);
            uint32_t sign = mn > mp ? 1u : 0u;
            if (mp == mn && mb != 0u && !first_nan && mode != FUSED_GIVEN_THRESH) {
                // +M and -M tie: the FIRST one in f order decides (std::max_element) -> lowest flat index with |c| == M
                // over the cluster; every CTA takes this branch together (mp, mn are cluster-wide values)
                const float M = __uint_as_float(mb);
                u64 best = ~0ull;
                for (int p = warp; p < g.npl; p += NW) {
                    const float* cs = C + p * g.PS;
                    const uint32_t f0 = (uint32_t)(g.gplane(p) * g.YZ);
                    for (int w = lane; w < g.YZ; w += 32) {
                        const float c = cs[w];
                        if (fabsf(c) == M) {
                            const u64 cand = ((u64)(f0 + (uint32_t)w) << 1) | (u64)(__float_as_uint(c) >> 31);
                            best = cand < best ? cand : best;
                        }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const u64 x = __shfl_xor_sync(0xffffffffu, best, o);
                    best = x < best ? x : best;
                }
                __syncthreads();               // s_red reuse
                if (lane == 0) s_red[warp] = best;
                __syncthreads();
                if (warp == 0) {
                    best = lane < NW ? s_red[lane] : ~0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const u64 x = __shfl_xor_sync(0xffffffffu, best, o);
                        best = x < best ? x : best;
                    }
                    if (lane == 0) s_x3[par] = best;
                }
                xsync();
                best = ~0ull;
                for (int r = 0; r < S; ++r) {
                    const u64 x = *cluster.map_shared_rank(s_x3 + par, r);
                    best = x < best ? x : best;
                }
                sign = (uint32_t)(best & 1ull);
            }
            ukey = ((u64)mb << 32) | 2ull | (u64)sign;      // the canonical key of k_fused_compress
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

        // ---------------- phase C1: kept count and last kept coefficient of every local segment ----------------
        const int nsl = g.npl * g.spp;         // local segments, in flat order: e = p * spp + q
#pragma unroll 1
        for (int e = warp; e < nsl; e += NW) {
            const int p = e / g.spp, w_lo = (e - p * g.spp) * XS_SEG, w_hi = min(g.YZ, w_lo + XS_SEG);
            const float* cs = C + p * g.PS;
            int cnt = 0, lw = 0;
            uint32_t lb = 0u;
            // 128 coefficients per trip: four independent loads and ballots (reads past w_hi stay inside the CTA's
            // shared memory and are masked out)
#pragma unroll 1
            for (int w0 = w_lo; w0 < w_hi; w0 += 128) {
                uint32_t bal[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int w = w0 + 32 * j + lane;
                    bal[j] = __ballot_sync(0xffffffffu, w < w_hi && keep_coef(cs[w], tf));
                }
                cnt += __popc(bal[0]) + __popc(bal[1]) + __popc(bal[2]) + __popc(bal[3]);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (bal[j]) { lb = bal[j]; lw = w0 + 32 * j; }
            }
            if (lane == 0) {
                s_cnt[e]  = cnt;
                s_last[e] = lb ? g.gplane(p) * g.YZ + lw + 31 - __clz(lb) : -1;
            }
            __syncwarp();
        }
        __syncthreads();
        // scan inside either range (low planes; high planes + trailing plane): pairs in front of every segment, last
        // kept flat index in front of it; the totals go to the cluster
        if (warp == 0) {
            int tot[2], lastf[2];
#pragma unroll
            for (int grp = 0; grp < 2; ++grp) {
                const int lo = grp ? g.nl * g.spp : 0, hi = grp ? nsl : g.nl * g.spp;
                int carry = 0, cmax = -1;
                for (int base = lo; base < hi; base += 32) {
                    const int e = base + lane;
                    const int c = e < hi ? s_cnt[e] : 0, l = e < hi ? s_last[e] : -1;
                    int isum = c, imax = l;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int ps = __shfl_up_sync(0xffffffffu, isum, o), pm = __shfl_up_sync(0xffffffffu, imax, o);
                        if (lane >= o) { isum += ps; imax = max(imax, pm); }
                    }
                    int emax = __shfl_up_sync(0xffffffffu, imax, 1);
                    if (lane == 0) emax = -1;
                    if (e < hi) { s_base[e] = carry + isum - c; s_prev[e] = max(cmax, emax); }
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
        xsync();                               // exchange 2
        // position of this CTA's two ranges in the unit's pair list: low ranges of ranks 0 .. S-1, then the high ranges
        if (warp == 0) {
            int c0 = 0, l0 = -1, c1 = 0, l1 = -1;
            if (lane < S) {
                const int* x = cluster.map_shared_rank(s_x2 + par * 4, lane);
                c0 = x[0]; l0 = x[1]; c1 = x[2]; l1 = x[3];
            }
            const bool before = lane < rank;
            const int blo = __reduce_add_sync(0xffffffffu, before ? c0 : 0), plo = __reduce_max_sync(0xffffffffu, before ? l0 : -1);
            const int bhi = __reduce_add_sync(0xffffffffu, before ? c1 : 0), phi = __reduce_max_sync(0xffffffffu, before ? l1 : -1);
            const int loa = __reduce_add_sync(0xffffffffu, c0), lla = __reduce_max_sync(0xffffffffu, l0);
            const int kk  = __reduce_add_sync(0xffffffffu, c0 + c1), la = __reduce_max_sync(0xffffffffu, max(l0, l1));
            if (lane == 0) {
                s_seed[0] = blo; s_seed[1] = plo; s_seed[2] = bhi + loa; s_seed[3] = max(phi, lla);
                s_seed[4] = kk;  s_seed[5] = la;
            }
        }
        __syncthreads();
        const int base_lo = s_seed[0], prev_lo = s_seed[1], base_hi = s_seed[2], prev_hi = s_seed[3], K = s_seed[4],
                  last_all = s_seed[5];
        int2* const tab = reinterpret_cast<int2*>(u.coef);     // decode-side plane table (cluster classes), or null
        if (tid == 0 && rank == 0) {
            const float M = fabsf(key_value(ukey));
            states[uid].npairs = K;
            states[uid].flags  = (first_nan ? UNIT_FLAG_NAN0 : 0) | ((K > 0 && unit_need32(M, tf)) ? UNIT_FLAG_NEED32 : 0);
            if (tab) tab[g.X] = make_int2(K, last_all);
        }

        // ---------------- phase C2: emit (run, value) pairs, a warp per segment ----------------
        int2* const out = reinterpret_cast<int2*>(u.out);
#pragma unroll 1
        for (int e = warp; e < nsl; e += NW) {
            const int p = e / g.spp, q = e - p * g.spp, w_lo = q * XS_SEG, w_hi = min(g.YZ, w_lo + XS_SEG);
            const bool hi_grp = p >= g.nl;
            int pos  = (hi_grp ? base_hi : base_lo) + s_base[e];
            int prev = max(hi_grp ? prev_hi : prev_lo, s_prev[e]);
            const int fstart = g.gplane(p) * g.YZ;
            if (tab && q == 0 && lane == 0) tab[g.gplane(p)] = make_int2(pos, prev);
            __syncwarp();
            const int scnt = s_cnt[e];
            if (scnt == 0) continue;
            const float* cs = C + p * g.PS;
            if (scnt == w_hi - w_lo) {
                // every coefficient kept (e.g. a negative max, SURVEY.md D3'): runs are 0, ranks are w - w_lo
                for (int w = w_lo + lane; w < w_hi; w += 32)
                    out[pos + (w - w_lo)] = make_int2(w == w_lo ? fstart + w_lo - prev - 1 : 0, __float_as_int(cs[w]));
                continue;
            }
#pragma unroll 1
            for (int w0 = w_lo; w0 < w_hi; w0 += 128) {
                float    c[4];
                bool     kf[4];
                uint32_t bal[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int w = w0 + 32 * j + lane;
                    c[j]   = cs[w];
                    kf[j]  = w < w_hi && keep_coef(c[j], tf);
                    bal[j] = __ballot_sync(0xffffffffu, kf[j]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (bal[j]) {                                      // warp-uniform
                        const uint32_t lower = bal[j] & lt;
                        const int f0 = fstart + w0 + 32 * j;           // flat index of lane 0's coefficient
                        // previous kept coefficient: a lower lane of this group, or the carried one
                        const int pf = lower ? f0 + 31 - __clz(lower) : prev;
                        if (kf[j]) out[pos + __popc(lower)] = make_int2(f0 + lane - pf - 1, __float_as_int(c[j]));
                        pos += __popc(bal[j]);
                        prev = f0 + 31 - __clz(bal[j]);
                    }
                }
            }
        }
        __syncthreads();                       // C and the segment arrays are rewritten by the next unit
    }
    xsync();                                   // no CTA may exit while a peer can still read its shared memory
}

// ---- decompress ----------------------------------------------------------------------------------------------------
// Block-wide exclusive prefix of run + 1 over one tile of NT * XS_PPT pairs (saturating at 2^30, so corrupt streams
// cannot wrap; negative runs are flagged, count as 0 and are skipped by the caller).  wt: 32 words per tile parity.
__device__ __forceinline__ uint32_t xs_sat_add(uint32_t a, uint32_t b) {
    const uint32_t s = a + b;
    return s > 0x40000000u ? 0x40000000u : s;
}
template <int NT>
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
    uint32_t winc = lane < NT / 32 ? wt[lane] : 0u;
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
template <int NT>
__device__ __forceinline__ void xs_decode_range(const int2* __restrict__ pairs, int pb, int pe, uint32_t cur, uint32_t f0,
                                                uint32_t f1, int pbase, const XGeom& g, float* C, uint32_t* s_wt, int& tile,
                                                bool& bad) {
    const uint32_t m_yz = xs_magic((uint32_t)g.YZ);            // d / YZ: d * YZ < 2^32 (d < 2^17 per range, YZ < 2^16)
    int2 pr[XS_PPT], nxt[XS_PPT];
    auto load = [&](int p, int2 (&dst)[XS_PPT]) {
#pragma unroll
        for (int j = 0; j < XS_PPT; ++j) dst[j] = (p + j < pe) ? __ldg(pairs + p + j) : make_int2(0, 0);
    };
    if (pb < pe) load(pb + (int)threadIdx.x * XS_PPT, nxt);
#pragma unroll 1
    for (int p0 = pb; p0 < pe; p0 += NT * XS_PPT, ++tile) {
        const int p = p0 + (int)threadIdx.x * XS_PPT;
#pragma unroll
        for (int j = 0; j < XS_PPT; ++j) pr[j] = nxt[j];
        if (p0 + NT * XS_PPT < pe) load(p + NT * XS_PPT, nxt);     // the next tile is in flight during this one's scan
        uint32_t ttot;
        uint32_t rp = xs_sat_add(cur, xs_tile_scan<NT>(pr, pe - p, s_wt + (tile & 1) * 32, bad, ttot));
#pragma unroll
        for (int j = 0; j < XS_PPT; ++j) {
            if (p + j < pe && pr[j].x >= 0) {
                const uint32_t f = xs_sat_add(rp, (uint32_t)pr[j].x);
                if (f >= f0 && f < f1) {
                    const uint32_t d = f - f0, pl = xs_div(d, m_yz);
                    C[(pbase + (int)pl) * g.PS + (int)(d - pl * (uint32_t)g.YZ)] = __int_as_float(pr[j].y);
                }
                rp = xs_sat_add(rp, (uint32_t)pr[j].x + 1u);
            }
        }
        cur = xs_sat_add(cur, ttot);
        __syncwarp();
    }
}

template <class T, int NT>
__device__ __forceinline__ void xs_inverse_store(const XGeom& g, const float* C, T* __restrict__ out) {
    const int tid = threadIdx.x;
    // full 2 x 2 x 2 blocks: X, then Y, then Z (src/decompressor.cpp:90-156)
    const int nblk = g.nl * g.hy * g.hz;
    const uint32_t m_nl = xs_magic(g.nl), m_hy = xs_magic(g.hy);
    const int sy = g.X, sz = g.X * g.Y;                                  // element strides of the box
    const int o_hx = g.nl * g.PS, o_hy = g.hy * g.Z, o_hz = g.hz;        // offsets of the high bands in C
#pragma unroll 1
    for (int q = tid; q < nblk; q += NT) {
        const int t = (int)xs_div(q, m_nl), al = q - t * g.nl, c = (int)xs_div(t, m_hy), b = t - c * g.hy;
        float v[8];
        const float* cd = C + al * g.PS + b * g.Z + c;
#pragma unroll
        for (int o = 0; o < 8; ++o) v[o] = cd[(o & 1) * o_hx + ((o >> 1) & 1) * o_hy + (o >> 2) * o_hz];
        haar_block_inverse_full(v);
        T* po = out + ((size_t)(2 * c) * sz + (size_t)(2 * b) * sy + (size_t)(2 * (g.a0 + al)));
#pragma unroll
        for (int o = 0; o < 8; o += 2) xs_st2(po + ((o >> 2) * sz + ((o >> 1) & 1) * sy), v[o], v[o + 1]);
    }
    // trailing cells of odd axes: the inverse leaves them at +0
    const int nxs = 2 * g.nl + g.own1;                    // x positions of this slab (incl. the trailing plane)
    auto xpos = [&](int i) { return i < 2 * g.nl ? 2 * g.a0 + i : g.X - 1; };
    if (g.oz) {
        for (int q = tid; q < nxs * g.Y; q += NT) {
            const int i = q % nxs, y = q / nxs;
            out[((size_t)(g.Z - 1) * g.Y + (size_t)y) * g.X + (size_t)xpos(i)] = (T)0;
        }
    }
    if (g.oy) {
        const int nz2 = g.Z - g.oz;
        for (int q = tid; q < nxs * nz2; q += NT) {
            const int i = q % nxs, z = q / nxs;
            out[((size_t)z * g.Y + (size_t)(g.Y - 1)) * g.X + (size_t)xpos(i)] = (T)0;
        }
    }
    if (g.own1) {
        const int ny2 = g.Y - g.oy, nz2 = g.Z - g.oz;
        for (int q = tid; q < ny2 * nz2; q += NT) {
            const int y = q % ny2, z = q / ny2;
            out[((size_t)z * g.Y + (size_t)y) * g.X + (size_t)(g.X - 1)] = (T)0;
        }
    }
}

// Work item = (unit, x-slab r of S); one CTA per item, items handed out through a global counter.
template <int NT, int CW, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_xs_decompress(const DecUnitDev* __restrict__ dec, const InvUnitDev* __restrict__ inv, const int* __restrict__ unit_list,
                int n_list, int S, int* __restrict__ err, int* __restrict__ work_counter) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* const    C      = reinterpret_cast<float*>(smem);
    uint32_t* const s_wt   = reinterpret_cast<uint32_t*>(smem + CW * 4);          // [2][32]
    int* const      s_item = reinterpret_cast<int*>(smem + CW * 4 + 256);
    const char** const s_pfp = reinterpret_cast<const char**>(smem + CW * 4 + 264);   // pair list of the item gridDim.x ahead
    int* const      s_pfk  = reinterpret_cast<int*>(smem + CW * 4 + 272);             // [2]: its K, its slab
    const int tid = threadIdx.x;
    const int n_items = n_list * S;
    bool bad = false;
    for (int k = 0;; ++k) {
        if (tid == 0) *s_item = work_counter ? atomicAdd(work_counter, 1) : (int)blockIdx.x + k * (int)gridDim.x;
        __syncthreads();
        const int item = *s_item;
        if (item >= n_items) break;
        const int uid = unit_list[item / S], rank = item % S;
        if (tid == 32 % NT) {
            // Items are handed out in order, so the item gridDim.x further on starts when this one ends, on this CTA or a
            // neighbour: its share of its unit's pair list (1/S of the lines: the S items of a unit run at about the same
            // time) goes into L2 while this item is inverted and stored.  Read here, consumed behind the decode's barrier.
            const int ahead = item + (int)gridDim.x;
            int k2 = 0;
            if (ahead < n_items) {
                const DecUnitDev* d2 = dec + unit_list[ahead / S];
                k2 = d2->npairs_dev ? *d2->npairs_dev : d2->npairs;
                k2 = max(0, min(k2, d2->total));
                *s_pfp = reinterpret_cast<const char*>(d2->pairs);
                s_pfk[1] = ahead % S;
            }
            s_pfk[0] = k2;
        }
        const DecUnitDev du = dec[uid];
        const InvUnitDev iu = inv[uid];
        XGeom g;
        g.init(iu.nx, iu.ny, iu.nz, S, rank);
        if (g.npl > 0) {
            // the slab's low and high planes start out zero (src/decompressor.cpp:16)
            {
                float4* c4 = reinterpret_cast<float4*>(C);
                for (int i = tid; i < (2 * g.nl * g.PS + 3) / 4; i += NT) c4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncthreads();
            const int2* pairs = reinterpret_cast<const int2*>(du.pairs);
            int K = du.npairs_dev ? *du.npairs_dev : du.npairs;
            K = max(0, min(K, du.total));
            int tile = 0;
            if (g.nl > 0) {
                if (S == 1) {
                    xs_decode_range<NT>(pairs, 0, K, 0u, 0u, (uint32_t)(2 * g.hx * g.YZ), 0, g, C, s_wt, tile, bad);
                } else {
                    const int2* tab = reinterpret_cast<const int2*>(du.coef);
                    const int2 l0 = tab[g.a0], l1 = tab[g.a0 + g.nl];
                    const int2 h0 = tab[g.hx + g.a0], h1 = tab[g.hx + g.a0 + g.nl];
                    xs_decode_range<NT>(pairs, max(0, l0.x), min(K, l1.x), (uint32_t)(l0.y + 1), (uint32_t)(g.a0 * g.YZ),
                                    (uint32_t)((g.a0 + g.nl) * g.YZ), 0, g, C, s_wt, tile, bad);
                    xs_decode_range<NT>(pairs, max(0, h0.x), min(K, h1.x), (uint32_t)(h0.y + 1), (uint32_t)((g.hx + g.a0) * g.YZ),
                                    (uint32_t)((g.hx + g.a0 + g.nl) * g.YZ), g.nl, g, C, s_wt, tile, bad);
                }
            }
            __syncthreads();
            if (const int k2 = s_pfk[0]) {
                const int lines = (int)(((size_t)k2 * 8 + 127) >> 7), per = (lines + S - 1) / S;
                const int l0 = s_pfk[1] * per, l1 = min(lines, l0 + per);
                const char* pb2 = *s_pfp;
                for (int i = l0 + tid; i < l1; i += NT)
                    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(pb2 + (size_t)i * 128));
            }
            if (iu.dtype == WC_F64) xs_inverse_store<double, NT>(g, C, static_cast<double*>(iu.out));
            else                    xs_inverse_store<float, NT>(g, C, static_cast<float*>(iu.out));
        }
        __syncthreads();                       // C and s_item are rewritten by the next item
    }
    if (bad) atomicOr(err, 1);
}

// ---- launchers -----------------------------------------------------------------------------------------------------
int xs_class_slabs(int fused_cls) {
    switch (fused_cls) {
    case FUSED_CLS_XS1S:
    case FUSED_CLS_XS1: return 1;
    case FUSED_CLS_XS2: return 2;
    case FUSED_CLS_XS4: return 4;
    case FUSED_CLS_XS8: return 8;
    }
    return 0;
}
int xs_class_of(int nx, int ny, int nz) {
    switch (xs_slabs(nx, ny, nz)) {
    case 1: return xs_fits_small(nx, ny, nz) ? FUSED_CLS_XS1S : FUSED_CLS_XS1;
    case 2: return FUSED_CLS_XS2;
    case 4: return FUSED_CLS_XS4;
    case 8: return FUSED_CLS_XS8;
    }
    return FUSED_CLS_NONE;
}

constexpr int XS_NT = 512, XS_NT_S = 128, XS_MINB_S = 5;
constexpr int XS_DSMEM   = XS_CWORDS * 4 + 512;        // decompress: C + scan scratch
constexpr int XS_DSMEM_S = XS_CWORDS_S * 4 + 512;

template <int NT, int CW, int MINB>
static cudaError_t launch_xs_c(int kid, int S, int mode, const UnitDev* units, UnitState* states, const int* unit_list,
                               int n_list, double one_minus_keep, const u64* global_key, int sm_count, cudaStream_t st,
                               LaunchStats* ls) {
    auto kern = k_xs_compress<NT, CW, MINB>;
    constexpr int smem = XSmem<CW>::TOTAL;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim         = dim3(NT);
    cfg.dynamicSmemBytes = smem;
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
        e = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
        if (e != cudaSuccess) return e;
        if (nc < 1) return cudaErrorLaunchOutOfResources;
        max_clusters = nc;
    }
    const int nc = max_clusters < n_list ? max_clusters : n_list;
    cfg.gridDim = dim3(nc * S);
    ls->begin(kid, st);
    e = cudaLaunchKernelEx(&cfg, kern, units, states, unit_list, n_list, one_minus_keep, global_key, mode);
    ls->end(st);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

cudaError_t launch_xs_compress(int fused_cls, int mode, const UnitDev* units, UnitState* states, const int* unit_list,
                               int n_list, double one_minus_keep, const u64* global_key, int sm_count, cudaStream_t st,
                               LaunchStats* ls) {
    const int S = xs_class_slabs(fused_cls);
    if (S == 0) return cudaErrorInvalidValue;
    if (n_list <= 0) return cudaSuccess;
    if (fused_cls == FUSED_CLS_XS1S)
        return launch_xs_c<XS_NT_S, XS_CWORDS_S, XS_MINB_S>(KID_XS_C1S, 1, mode, units, states, unit_list, n_list,
                                                            one_minus_keep, global_key, sm_count, st, ls);
    const int kid = S == 1 ? KID_XS_C1 : S == 2 ? KID_XS_C2 : S == 4 ? KID_XS_C4 : KID_XS_C8;
    return launch_xs_c<XS_NT, XS_CWORDS, 1>(kid, S, mode, units, states, unit_list, n_list, one_minus_keep, global_key,
                                            sm_count, st, ls);
}

template <int NT, int CW, int MINB>
static cudaError_t launch_xs_d(int kid, int S, int smem, const DecUnitDev* dec, const InvUnitDev* inv, const int* unit_list,
                               int n_list, int* err, int sm_count, cudaStream_t st, LaunchStats* ls, int* work_counter) {
    auto kern = k_xs_decompress<NT, CW, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int& per_sm = ls->occ[kid];                // resident CTAs per SM, cached per ctx
    if (per_sm == 0) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    }
    const long long items = (long long)n_list * S, slots = (long long)per_sm * sm_count;
    const int nc = (int)(items < slots ? items : slots);
    ls->begin(kid, st);
    kern<<<nc, NT, smem, st>>>(dec, inv, unit_list, n_list, S, err, work_counter);
    ls->end(st);
    return cudaGetLastError();
}

cudaError_t launch_xs_decompress(int fused_cls, const DecUnitDev* dec, const InvUnitDev* inv, const int* unit_list,
                                 int n_list, int* err, int sm_count, cudaStream_t st, LaunchStats* ls, int* work_counter) {
    const int S = xs_class_slabs(fused_cls);
    if (S == 0) return cudaErrorInvalidValue;
    if (n_list <= 0) return cudaSuccess;
    if (fused_cls == FUSED_CLS_XS1S)
        return launch_xs_d<XS_NT_S, XS_CWORDS_S, XS_MINB_S>(KID_XS_DS, 1, XS_DSMEM_S, dec, inv, unit_list, n_list, err,
                                                            sm_count, st, ls, work_counter);
    return launch_xs_d<XS_NT, XS_CWORDS, 1>(KID_XS_D, S, XS_DSMEM, dec, inv, unit_list, n_list, err, sm_count, st, ls,
                                            work_counter);
}

} // namespace wc
