// Fused on-chip compress kernel for sm_100a: one unit's coefficients never leave the SM(s).
//
//   k_fused_compress<R>: a unit (box, component) is owned by a cluster of R CTAs (R = 1 or 8).
//   CTA r takes the block-rows b in [r*nb, (r+1)*nb) (a y-slab), so
//     - its input is, per z-plane, ONE contiguous piece of 2*nb rows  -> TMA bulk copies
//       (cp.async.bulk.shared::cluster.global.mbarrier) issued by a dedicated producer warp into a
//       4-stage shared-memory ring, running ahead of the consumers across unit boundaries;
//     - its coefficients are the rows j' in {b, hy+b} of every i': 2*X "segments" of nb*Z values that
//       are contiguous both in the CTA's shared-memory array C and in the global f order, so the
//       ordered (run,value) packing only needs per-segment counts from the other CTAs (DSMEM push
//       + cluster-scope mbarrier), never their coefficients.
//   16 consumer warps: (A) 2x2x2 Haar blocks from the staged float64/float32 rows -> C (f order,
//   padded against bank conflicts) + running max of +c and -c; (B) threshold of
//   src/compressor.cpp:212-216; (C1) per-segment count / last-kept; (C2) ballot-ranked emission of
//   (run, value) pairs straight to the unit's slot in HBM.
//   HBM traffic per unit = 8N (or 4N) in + 8K out: the algorithmic minimum of SURVEY.md §8d.
#include <cstdio>

#include "wc_common.cuh"
#include "wc_fused.h"

namespace wc {

// ---- compile-time geometry ---------------------------------------------------------------------
constexpr int F_CWARPS      = 16;                 // consumer warps
constexpr int F_CONSUMERS   = F_CWARPS * 32;      // 512
constexpr int F_THREADS     = F_CONSUMERS + 32;   // + producer warp
constexpr int F_NGROUPS     = 4;                  // consumer groups; chunk k is consumed by group k % 4
constexpr int F_GROUP       = F_CONSUMERS / F_NGROUPS;   // 128 threads = one warp per SM sub-partition
constexpr int F_CAP         = 32768;              // coefficients per CTA
constexpr int F_PAD         = 4;                  // padding words per i' slab of C (keeps float4/float2
                                                  // alignment, makes a-lanes hit banks 4a + ...)
constexpr int F_CPAD        = 256;                // total padding words of C (F_PAD * X, X <= 64)
constexpr int F_STAGE       = 16384;              // payload bytes per stage
constexpr int F_PADP        = 16;                 // bytes between plane pieces in a stage (LDS.128 banks)
constexpr int F_STAGE_ALLOC = F_STAGE + 64 * F_PADP;
constexpr int F_NSTAGES     = 5;
constexpr int F_MAXSEG      = 128;                // segments (2*X) per CTA
constexpr int F_MAXG        = 1024;               // gathered segment entries (2*X*R)

constexpr int SM_C      = 0;
constexpr int SM_STAGE  = SM_C + (F_CAP + F_CPAD) * 4;
constexpr int SM_G      = SM_STAGE + F_NSTAGES * F_STAGE_ALLOC;   // [2][F_MAXG] u32: cnt << 16 | last
constexpr int SM_BASE   = SM_G + 2 * F_MAXG * 4;                  // [F_MAXSEG] int
constexpr int SM_PREV   = SM_BASE + F_MAXSEG * 4;                 // [F_MAXSEG] int
constexpr int SM_RED    = SM_PREV + F_MAXSEG * 4;                 // scratch: 64 x 8 bytes
constexpr int SM_XS1    = SM_RED + 64 * 8;                        // [2][8] u64 exchange slots
constexpr int SM_XS2    = SM_XS1 + 16 * 8;                        // [2][8] u64
constexpr int SM_BARS   = SM_XS2 + 16 * 8;                        // full[5] empty[5] x1 x2 x3
constexpr int SM_TOTAL  = SM_BARS + 20 * 8;                       // + gen[5] u32
static_assert(SM_TOTAL <= 232448, "exceeds the 227 KB per-CTA shared memory of sm_100");
static_assert(SM_STAGE % 128 == 0 && F_STAGE_ALLOC % 128 == 0, "stage alignment");

struct FGeom {
    int X, Y, Z, hx, hy, hz, es;
    int nb;        // block-rows (y) per CTA
    int CB, CZ;    // chunk extents in blocks (CZ even: a thread transforms two c-adjacent blocks at once)
    int ncb, ncz;  // chunks along b and c
    int seglen;    // nb * Z
    int nseg;      // 2 * X
    int nlocal;    // X * 2 * nb * Z
    int slab;      // padded words of C per i'
};

__host__ __device__ inline bool fused_geom(int X, int Y, int Z, int dtype, int R, FGeom& g) {
    if (X < 2 || Y < 2 || Z < 4 || (X & 1) || (Y & 1) || (Z & 3)) return false;
    g.X = X; g.Y = Y; g.Z = Z;
    g.hx = X / 2; g.hy = Y / 2; g.hz = Z / 2;
    g.es = dtype == WC_F64 ? 8 : 4;
    if ((X * g.es) % 16) return false;
    if (2 * X > F_MAXSEG) return false;
    long long n = (long long)X * Y * Z;
    if (g.hy % R) return false;
    if (n / R > F_CAP) return false;
    g.nb = g.hy / R;
    int cz0 = (g.hz % 4 == 0) ? 4 : 2;
    int row_pair = 2 * X * g.es;                 // bytes of one block-row (2 y rows) of one z-plane
    if (2 * cz0 * row_pair > F_STAGE) cz0 = 2;
    int per_b = 2 * cz0 * row_pair;              // bytes of one block-row for cz0 block-planes
    if (per_b > F_STAGE) return false;
    int cb = F_STAGE / per_b;
    if (cb >= g.nb) {
        g.CB = g.nb;
        int cz = F_STAGE / (2 * g.nb * row_pair);
        cz -= cz % cz0;
        if (cz > g.hz) cz = g.hz;
        if (cz > 32) cz = 32;
        g.CZ = cz;
    } else {
        g.CB = cb;
        g.CZ = cz0;
    }
    g.ncb  = (g.nb + g.CB - 1) / g.CB;
    g.ncz  = (g.hz + g.CZ - 1) / g.CZ;
    g.seglen = g.nb * Z;
    g.nseg   = 2 * X;
    g.nlocal = g.nseg * g.seglen;
    g.slab   = 2 * g.nb * Z + F_PAD;
    return true;
}

int fused_class(int nx, int ny, int nz, int dtype, const void* ptr) {
    if (reinterpret_cast<uintptr_t>(ptr) & 15u) return 0;
    FGeom g;
    if (fused_geom(nx, ny, nz, dtype, 1, g)) return 1;
    if (fused_geom(nx, ny, nz, dtype, 8, g)) return 8;
    return 0;
}
bool fused_decode_available() { return false; }

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait_cluster(bar, parity)) { }
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// Hint: pull [src, src+bytes) into L2 (no destination).  Used one unit ahead of the TMA ring so the
// ring refills at L2 latency instead of HBM latency.
__device__ __forceinline__ void l2_prefetch(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t ld_volatile_shared_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_shared_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_u64(uint32_t addr, u64 v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t nclusters_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(F_CONSUMERS) : "memory"); }

// exact q / d for q < 65536, d < 65536 (m = ceil(2^32 / d); d == 1 handled by the caller's m == 0)
__device__ __forceinline__ uint32_t fdiv(uint32_t q, uint32_t m) { return m ? __umulhi(q, m) : q; }
__device__ __forceinline__ uint32_t fdiv_magic(uint32_t d) { return d <= 1 ? 0u : (0xffffffffu / d) + 1u; }

// ---- the kernel -----------------------------------------------------------------------------------
// Per-CTA phase cycle counters (clock64 by consumer thread 0 at the phase boundaries): a built-in
// light-weight profile, read back with wc_debug_phase_cycles().  [cta][phase], phases: 0 = A (transform,
// includes waiting for TMA data), 1 = B (threshold), 2 = C1 (count), 3 = scan (+ cluster exchange),
// 4 = C2 (emit), 5 = units processed.
__device__ unsigned long long g_phase_cycles[1024][6];
__device__ unsigned long long g_a_cycles[1024][4];   // debug: gen spin, full wait, transform, chunks

__device__ __forceinline__ void st_pair_pred(bool p, int2* addr, int run, float val) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.u32 q, %0, 0;\n\t"
        "@q st.global.v2.b32 [%1], {%2, %3};\n\t}"
        ::"r"((uint32_t)p), "l"(addr), "r"(run), "r"(__float_as_int(val)) : "memory");
}

// One 1-D Haar step on two independent blocks at once (packed f32x2): lo = (lo+hi)*0.5, hi = (lo-hi)*0.5,
// each rounded exactly like the scalar __fadd_rn / __fsub_rn / __fmul_rn sequence: hi*(-1)+lo is the
// correctly rounded difference, and the *0.5 is a separate rounding step as in the reference.
__device__ __forceinline__ void haar_pair2(float2& lo, float2& hi) {
    const float2 half = make_float2(0.5f, 0.5f), neg1 = make_float2(-1.f, -1.f);
    float2 s = __fadd2_rn(lo, hi);
    float2 d = __ffma2_rn(hi, neg1, lo);
    lo = __fmul2_rn(s, half);
    hi = __fmul2_rn(d, half);
}

// Transforms the two c-adjacent blocks whose 4 z-planes start at `p0` in the stage and stores the 8 x 2
// coefficients into C.  v[zi*4+yi*2+xi] = (block c, block c+1).
template <int ES>
__device__ __forceinline__ float transform_pair(const unsigned char* p0, uint32_t pstride, uint32_t row_bytes,
                                                float* cdst, int o1, int o2, int o3, float& bp, float& bn) {
    float2 v[8];
#pragma unroll
    for (int zi = 0; zi < 2; ++zi)
#pragma unroll
        for (int yi = 0; yi < 2; ++yi) {
            const unsigned char* pa = p0 + zi * pstride + yi * row_bytes;
            const unsigned char* pb = pa + 2 * pstride;
            if (ES == 8) {
                double2 da = *reinterpret_cast<const double2*>(pa);
                double2 db = *reinterpret_cast<const double2*>(pb);
                v[zi * 4 + yi * 2]     = make_float2(__double2float_rn(da.x), __double2float_rn(db.x)); // src/preprocess.cpp:78
                v[zi * 4 + yi * 2 + 1] = make_float2(__double2float_rn(da.y), __double2float_rn(db.y));
            } else {
                float2 fa = *reinterpret_cast<const float2*>(pa);
                float2 fb = *reinterpret_cast<const float2*>(pb);
                v[zi * 4 + yi * 2]     = make_float2(fa.x, fb.x);
                v[zi * 4 + yi * 2 + 1] = make_float2(fa.y, fb.y);
            }
        }
    // Z, then Y, then X (src/compressor.cpp:98-175)
#pragma unroll
    for (int q = 0; q < 4; ++q) haar_pair2(v[q], v[4 + q]);
#pragma unroll
    for (int zi = 0; zi < 2; ++zi)
#pragma unroll
        for (int xi = 0; xi < 2; ++xi) haar_pair2(v[zi * 4 + xi], v[zi * 4 + 2 + xi]);
#pragma unroll
    for (int q = 0; q < 4; ++q) haar_pair2(v[2 * q], v[2 * q + 1]);
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        const int idx = (o & 1) * o1 + ((o >> 1) & 1) * o2 + (o >> 2) * o3;
        *reinterpret_cast<float2*>(cdst + idx) = v[o];
        bp = fmaxf(fmaxf(bp, v[o].x), v[o].y);
        bn = fmaxf(fmaxf(bn, -v[o].x), -v[o].y);
    }
    return v[0].x;
}

template <int R>
__global__ void __launch_bounds__(F_THREADS, 1)
k_fused_compress(const UnitDev* __restrict__ units, UnitState* __restrict__ states,
                 const int* __restrict__ unit_list, int n_list, double one_minus_keep,
                 const u64* __restrict__ global_key, int mode) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* const    C      = reinterpret_cast<float*>(smem + SM_C);
    uint32_t* const g_pk   = reinterpret_cast<uint32_t*>(smem + SM_G);
    int* const      s_base = reinterpret_cast<int*>(smem + SM_BASE);
    int* const      s_prev = reinterpret_cast<int*>(smem + SM_PREV);
    u64* const      s_red  = reinterpret_cast<u64*>(smem + SM_RED);
    u64* const      xs1    = reinterpret_cast<u64*>(smem + SM_XS1);
    u64* const      xs2    = reinterpret_cast<u64*>(smem + SM_XS2);
    const uint32_t bars  = smem_u32(smem + SM_BARS);
    const uint32_t full0 = bars, empty0 = bars + 8 * F_NSTAGES;
    const uint32_t xb1 = bars + 8 * (2 * F_NSTAGES), xb2 = xb1 + 8, xb3 = xb2 + 8;
    const uint32_t gen0 = xb3 + 8;   // [F_NSTAGES] u32: chunk number each stage currently holds
    const uint32_t stage0 = smem_u32(smem + SM_STAGE);

    const int tid  = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = R > 1 ? cluster_ctarank() : 0u;
    const uint32_t cid  = R > 1 ? cluster_id_x() : blockIdx.x;
    const uint32_t ncl  = R > 1 ? nclusters_x() : gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < F_NSTAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, F_GROUP / 32);
        }
        mbar_init(xb1, R);
        mbar_init(xb2, R);
        mbar_init(xb3, R);
        for (int s = 0; s < F_NSTAGES; ++s) st_volatile_shared_u32(gen0 + 4 * s, 0xffffffffu);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (R > 1) cluster_sync_all();

    if (warp == F_CWARPS) {
        // =============================== producer warp ===============================
        uint32_t kg = 0;
        for (int ui = cid; ui < n_list; ui += ncl) {
            const UnitDev u = units[unit_list[ui]];
            FGeom g;
            fused_geom(u.nx, u.ny, u.nz, u.dtype, R, g);
            const char* in  = static_cast<const char*>(u.in);
            const int   b0  = rank * g.nb;
            const size_t row_bytes   = (size_t)g.X * g.es;
            const size_t plane_bytes = row_bytes * g.Y;
            // L2 prefetch of this CTA's whole slab of the NEXT unit: one contiguous piece per z-plane
            if (ui + (int)ncl < n_list) {
                const UnitDev un = units[unit_list[ui + ncl]];
                FGeom gn;
                fused_geom(un.nx, un.ny, un.nz, un.dtype, R, gn);
                const size_t rb = (size_t)gn.X * gn.es, pb = rb * gn.Y;
                const char* base = static_cast<const char*>(un.in) + (size_t)(2 * rank * gn.nb) * rb;
                const uint32_t slab_bytes = (uint32_t)(2 * gn.nb * rb);
                if (R == 1) {
                    // the slab is the whole box: contiguous, prefetch in 16 KB pieces
                    const size_t total = pb * gn.Z;
                    for (size_t off = (size_t)lane * 16384; off < total; off += 32 * 16384)
                        l2_prefetch(base + off, (uint32_t)min((size_t)16384, total - off));
                } else {
                    for (int z = lane; z < gn.Z; z += 32) l2_prefetch(base + (size_t)z * pb, slab_bytes);
                }
            }
            for (int icb = 0; icb < g.ncb; ++icb) {
                const int bc0 = icb * g.CB;
                const int cbc = min(g.CB, g.nb - bc0);
                const uint32_t piece_bytes = (uint32_t)(2 * cbc * row_bytes);
                const uint32_t pstride     = piece_bytes + F_PADP;
                for (int icz = 0; icz < g.ncz; ++icz, ++kg) {
                    const int cc0 = icz * g.CZ;
                    const int czc = min(g.CZ, g.hz - cc0);
                    const int npieces = 2 * czc;
                    const uint32_t s = kg % F_NSTAGES;
                    mbar_wait(empty0 + 8 * s, ((kg / F_NSTAGES) & 1) ^ 1);
                    if (lane == 0) {
                        st_volatile_shared_u32(gen0 + 4 * s, kg);   // stage s now belongs to chunk kg
                        mbar_arrive_expect_tx(full0 + 8 * s, piece_bytes * npieces);
                    }
                    __syncwarp();
                    const char* src0 = in + (size_t)(2 * cc0) * plane_bytes + (size_t)(2 * (b0 + bc0)) * row_bytes;
                    for (int p = lane; p < npieces; p += 32)
                        tma_load_1d(stage0 + s * F_STAGE_ALLOC + p * pstride, src0 + (size_t)p * plane_bytes,
                                    piece_bytes, full0 + 8 * s);
                }
            }
        }
    } else {
        // =============================== consumer warps ===============================
        const int group = warp >> 2;              // chunk k belongs to group k % 4
        const int tig   = tid & (F_GROUP - 1);
        const uint32_t lt = lanemask_lt();
        uint32_t kg = 0, xph1 = 0, xph2 = 0, xph3 = 0;
        for (int ui = cid; ui < n_list; ui += ncl) {
            const int     uid = unit_list[ui];
            const UnitDev u   = units[uid];
            FGeom g;
            fused_geom(u.nx, u.ny, u.nz, u.dtype, R, g);
            const int b0 = rank * g.nb;
            const uint32_t m_hx = fdiv_magic(g.hx);
            const uint32_t row_bytes = (uint32_t)g.X * g.es;
            const int o1 = g.hx * g.slab, o2 = g.nb * g.Z, o3 = g.hz;
            float bp = 0.f, bn = 0.f;             // running max of +c and of -c
            bool  nan0 = false;
            long long t0 = clock64();

            // ---------------- phase A: transform the staged rows into C ----------------
            int c_czc = -1, c_cbc = -1, c_npairs = 0;     // cached decomposition of pair index `tig`
            uint32_t c_mnpc = 0, c_src = 0;
            int c_crel = 0;
            bool c_first = false;
            for (int icb = 0; icb < g.ncb; ++icb) {
                const int bc0 = icb * g.CB;
                const int cbc = min(g.CB, g.nb - bc0);
                const uint32_t pstride = 2 * cbc * row_bytes + F_PADP;
                for (int icz = 0; icz < g.ncz; ++icz, ++kg) {
                    if ((int)(kg & (F_NGROUPS - 1)) != group) continue;
                    const int cc0 = icz * g.CZ;
                    const int czc = min(g.CZ, g.hz - cc0);
                    const int npc = czc >> 1;                    // c-pairs in this chunk
                    if (czc != c_czc || cbc != c_cbc) {
                        c_czc = czc; c_cbc = cbc;
                        c_npairs = npc * cbc * g.hx;
                        c_mnpc   = fdiv_magic(npc);
                        // pair q -> (cp fastest, a, bl)
                        uint32_t t1 = fdiv(tig, c_mnpc), cp = tig - t1 * npc;
                        uint32_t bl = fdiv(t1, m_hx), a = t1 - bl * g.hx;
                        c_src   = (4 * cp) * pstride + (2 * bl) * row_bytes + a * 2 * g.es;
                        c_crel  = a * g.slab + bl * g.Z + 2 * cp;
                        c_first = (a == 0 && bl == 0 && cp == 0);
                    }
                    const uint32_t s = kg % F_NSTAGES;
                    // The groups run independently, so this group may get here before the PREVIOUS use of
                    // stage s (chunk kg - NSTAGES, another group's) has even landed.
                    // A parity wait alone is ambiguous here: it cannot tell "the previous use of the stage has
                    // not landed yet" from "this use has landed".  The producer therefore publishes the chunk
                    // number it is filling stage s with (after the stage was released), and the group waits
                    // for that first.
                    long long ta = clock64();
                    while (ld_volatile_shared_u32(gen0 + 4 * s) != kg) { }
                    long long tb = clock64();
                    mbar_wait(full0 + 8 * s, (kg / F_NSTAGES) & 1);
                    long long tc = clock64();
                    const unsigned char* st = smem + SM_STAGE + s * F_STAGE_ALLOC;
                    float* const cbase = C + bc0 * g.Z + cc0;
                    for (int q = tig; q < c_npairs; q += F_GROUP) {
                        uint32_t src = c_src;
                        int crel = c_crel;
                        bool first = c_first;
                        if (q != tig) {
                            uint32_t t1 = fdiv(q, c_mnpc), cp = q - t1 * npc;
                            uint32_t bl = fdiv(t1, m_hx), a = t1 - bl * g.hx;
                            src   = (4 * cp) * pstride + (2 * bl) * row_bytes + a * 2 * g.es;
                            crel  = a * g.slab + bl * g.Z + 2 * cp;
                            first = false;
                        }
                        float v0 = g.es == 8
                            ? transform_pair<8>(st + src, pstride, row_bytes, cbase + crel, o1, o2, o3, bp, bn)
                            : transform_pair<4>(st + src, pstride, row_bytes, cbase + crel, o1, o2, o3, bp, bn);
                        if (first && rank == 0 && bc0 == 0 && cc0 == 0) nan0 = isnan(v0);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty0 + 8 * s);
                    if (tid == 0 && blockIdx.x < 1024) {
                        long long td = clock64();
                        unsigned long long* pa = g_a_cycles[blockIdx.x];
                        pa[0] += tb - ta; pa[1] += tc - tb; pa[2] += td - tc; pa[3] += 1;
                    }
                }
            }

            // ---------------- phase B: the threshold ----------------
            long long t1 = clock64();
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                bp = fmaxf(bp, __shfl_xor_sync(0xffffffffu, bp, o));
                bn = fmaxf(bn, __shfl_xor_sync(0xffffffffu, bn, o));
            }
            const bool any_nan0 = __any_sync(0xffffffffu, nan0);
            if (lane == 0) {
                // bn >= 0, so its sign bit is free: it carries "the coefficient at f = 0 is NaN"
                s_red[warp] = ((u64)__float_as_uint(bp) << 32) |
                              (u64)((__float_as_uint(bn) & 0x7fffffffu) | (any_nan0 ? 0x80000000u : 0u));
            }
            consumer_bar();
            float Mp = 0.f, Mn = 0.f;
            bool  first_nan = false;
            {
                u64 x = s_red[lane & (F_CWARPS - 1)];
                float p = __uint_as_float((uint32_t)(x >> 32));
                float n = __uint_as_float((uint32_t)x & 0x7fffffffu);
                first_nan = __any_sync(0xffffffffu, ((uint32_t)x >> 31) != 0);
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {
                    p = fmaxf(p, __shfl_xor_sync(0xffffffffu, p, o));
                    n = fmaxf(n, __shfl_xor_sync(0xffffffffu, n, o));
                }
                Mp = p; Mn = n;
            }
            if (R > 1) {
                // all-gather (Mp | Mn) over the cluster
                const uint32_t par = xph1 & 1;
                if (tid < R) {
                    u64 pay = ((u64)__float_as_uint(Mp) << 32) |
                              (u64)((__float_as_uint(Mn) & 0x7fffffffu) | (first_nan ? 0x80000000u : 0u));
                    st_cluster_u64(mapa(smem_u32(&xs1[par * 8 + rank]), tid), pay);
                    mbar_arrive_remote(mapa(xb1, tid));
                }
                mbar_wait_cluster(xb1, par);
                ++xph1;
                float p = 0.f, n = 0.f;
                bool fn = false;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    u64 x = xs1[par * 8 + r];
                    fn = fn || (((uint32_t)x >> 31) != 0);
                    p  = fmaxf(p, __uint_as_float((uint32_t)(x >> 32)));
                    n  = fmaxf(n, __uint_as_float((uint32_t)x & 0x7fffffffu));
                }
                Mp = p; Mn = n; first_nan = fn;
            }
            float M = fmaxf(Mp, Mn);
            uint32_t sign = Mn > Mp ? 1u : 0u;
            if (Mp == Mn && M != 0.f && !first_nan && mode != FUSED_GIVEN_THRESH) {
                // +M and -M tie: the FIRST one in f order decides (std::max_element) -> find min f
                u64 best = ~0ull;
                const uint32_t m_sl = fdiv_magic(g.seglen);
                for (int l = tid; l < g.nlocal; l += F_CONSUMERS) {
                    uint32_t sg = fdiv(l, m_sl), w = l - sg * g.seglen;
                    float c = C[l + F_PAD * (sg >> 1)];
                    if (fabsf(c) == M) {
                        uint32_t f = ((sg >> 1) * g.Y + (sg & 1) * g.hy + b0) * g.Z + w;
                        u64 cand = ((u64)f << 1) | (u64)(__float_as_uint(c) >> 31);
                        best = cand < best ? cand : best;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    u64 x = __shfl_xor_sync(0xffffffffu, best, o);
                    best = x < best ? x : best;
                }
                consumer_bar();   // s_red reuse
                if (lane == 0) s_red[warp] = best;
                consumer_bar();
                best = s_red[lane & (F_CWARPS - 1)];
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {
                    u64 x = __shfl_xor_sync(0xffffffffu, best, o);
                    best = x < best ? x : best;
                }
                if (R > 1) {
                    const uint32_t par = xph3 & 1;
                    if (tid < R) {
                        st_cluster_u64(mapa(smem_u32(&xs2[par * 8 + rank]), tid), best);
                        mbar_arrive_remote(mapa(xb3, tid));
                    }
                    mbar_wait_cluster(xb3, par);
                    ++xph3;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        u64 x = xs2[par * 8 + r];
                        best = x < best ? x : best;
                    }
                }
                sign = (uint32_t)(best & 1ull);
            }
            float tf;
            {
                u64 key = ((u64)__float_as_uint(M) << 32) | 2ull | (u64)sign;
                if (mode == FUSED_GIVEN_THRESH) {
                    u64 gk = *global_key;
                    tf = threshold_float(gk & ~(1ull << 63), (gk >> 63) != 0, one_minus_keep);
                } else {
                    tf = threshold_float(key, first_nan, one_minus_keep);
                }
                if (tid == 0 && rank == 0) {
                    states[uid].key      = key;
                    states[uid].flags    = first_nan ? 1 : 0;
                    states[uid].thresh_f = tf;
                }
            }
            if (mode == FUSED_KEYS_ONLY) {
                consumer_bar();   // C is rewritten by the next unit
                continue;
            }

            // ---------------- phase C1: per-segment count and last kept ----------------
            long long t2 = clock64();
            const int gpar = R > 1 ? (int)(xph2 & 1) : 0;
            uint32_t* const my_pk = g_pk + gpar * F_MAXG;
            for (int sg = warp; sg < g.nseg; sg += F_CWARPS) {
                const float* cs = C + sg * g.seglen + F_PAD * (sg >> 1);   // 16-byte aligned
                int cnt = 0, lb = 0;
                uint32_t lm = 0;
                for (int w = lane * 4; w < g.seglen; w += 128) {           // seglen % 4 == 0
                    const float4 c = *reinterpret_cast<const float4*>(cs + w);
                    uint32_t m = (keep_coef(c.x, tf) ? 1u : 0u) | (keep_coef(c.y, tf) ? 2u : 0u) |
                                 (keep_coef(c.z, tf) ? 4u : 0u) | (keep_coef(c.w, tf) ? 8u : 0u);
                    cnt += __popc(m);
                    if (m) { lb = w; lm = m; }
                }
                int last = lm ? lb + 31 - __clz(lm) : -1;
                cnt  = __reduce_add_sync(0xffffffffu, cnt);
                last = __reduce_max_sync(0xffffffffu, last);
                const uint32_t pk = ((uint32_t)cnt << 16) | ((uint32_t)last & 0xffffu);
                if (R == 1) {
                    if (lane == 0) my_pk[sg] = pk;
                } else if (lane < R) {
                    st_cluster_u32(mapa(smem_u32(&my_pk[sg * R + rank]), lane), pk);
                }
            }
            long long t3 = clock64();
            if (R > 1) {
                fence_cluster();
                consumer_bar();
                if (tid < R) mbar_arrive_remote(mapa(xb2, tid));
                mbar_wait_cluster(xb2, gpar);
                ++xph2;
            } else {
                consumer_bar();
            }

            // ---------------- scan over the segments in global order ----------------
            {
                const int NG = g.nseg * R;   // <= 1024, entry e = sg * R + r
                int cv[2], lv[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int e = tid * 2 + j;
                    cv[j] = 0; lv[j] = -1;
                    if (e < NG) {
                        const uint32_t pk = my_pk[e];
                        const int sg = e / R, r = e % R;
                        cv[j] = (int)(pk >> 16);
                        if ((pk & 0xffffu) != 0xffffu)
                            lv[j] = ((sg >> 1) * g.Y + (sg & 1) * g.hy + r * g.nb) * g.Z + (int)(pk & 0xffffu);
                    }
                }
                int isum = cv[0] + cv[1], imax = max(lv[0], lv[1]);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int ps = __shfl_up_sync(0xffffffffu, isum, o);
                    int pm = __shfl_up_sync(0xffffffffu, imax, o);
                    if (lane >= o) { isum += ps; imax = max(imax, pm); }
                }
                int* sr = reinterpret_cast<int*>(s_red);
                if (lane == 31) { sr[warp] = isum; sr[32 + warp] = imax; }
                consumer_bar();
                int wsum = 0, wmax = -1, total = 0;
#pragma unroll
                for (int i = 0; i < F_CWARPS; ++i) {
                    int xs = sr[i], xm = sr[32 + i];
                    if (i < warp) { wsum += xs; wmax = max(wmax, xm); }
                    total += xs;
                }
                int es = __shfl_up_sync(0xffffffffu, isum, 1);
                int em = __shfl_up_sync(0xffffffffu, imax, 1);
                if (lane == 0) { es = 0; em = -1; }
                es += wsum;
                em = max(em, wmax);
                const int e0 = tid * 2;
                if (e0 < NG && (R == 1 || (e0 % R) == (int)rank)) { s_base[e0 / R] = es; s_prev[e0 / R] = em; }
                if (e0 + 1 < NG && (R == 1 || ((e0 + 1) % R) == (int)rank)) {
                    s_base[(e0 + 1) / R] = es + cv[0];
                    s_prev[(e0 + 1) / R] = max(em, lv[0]);
                }
                if (tid == 0 && rank == 0) states[uid].npairs = total;
                consumer_bar();
            }

            // ---------------- phase C2: emit (run, value) pairs ----------------
            long long t4 = clock64();
            for (int sg = warp; sg < g.nseg; sg += F_CWARPS) {
                const int scnt = (int)(my_pk[sg * R + rank] >> 16);
                if (scnt == 0) continue;                                   // nothing kept in this segment
                const float* cs = C + sg * g.seglen + F_PAD * (sg >> 1);
                const int fstart = ((sg >> 1) * g.Y + (sg & 1) * g.hy + b0) * g.Z;
                uint32_t pos = (uint32_t)s_base[sg];
                int prev = s_prev[sg];
                int2* const out = reinterpret_cast<int2*>(u.out);
                if (scnt == g.seglen) {
                    // every coefficient kept (e.g. a negative max, SURVEY.md D3'): runs are 0, ranks are w
                    for (int w = lane; w < g.seglen; w += 32)
                        out[pos + (uint32_t)w] = make_int2(w == 0 ? fstart - prev - 1 : 0, __float_as_int(cs[w]));
                    continue;
                }
                for (int w0 = 0; w0 < g.seglen; w0 += 128) {
                    float    c[4];
                    uint32_t bal[4];
                    bool     kf[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int w = w0 + 32 * j + lane;
                        const bool ok = w < g.seglen;
                        c[j]   = ok ? cs[w] : 0.f;
                        kf[j]  = ok && keep_coef(c[j], tf);
                        bal[j] = __ballot_sync(0xffffffffu, kf[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (bal[j] == 0u) continue;                      // warp-uniform
                        const int f0 = fstart + w0 + 32 * j;
                        const uint32_t lower = bal[j] & lt;
                        const int pf = lower ? f0 + 31 - __clz(lower) : prev;
                        st_pair_pred(kf[j], out + (pos + __popc(lower)), f0 + lane - pf - 1, c[j]);
                        pos += __popc(bal[j]);
                        prev = f0 + 31 - __clz(bal[j]);
                    }
                }
            }
            consumer_bar();   // C and the segment arrays are rewritten by the next unit
            if (tid == 0 && blockIdx.x < 1024) {
                long long t5 = clock64();
                unsigned long long* pc = g_phase_cycles[blockIdx.x];
                pc[0] += t1 - t0; pc[1] += t2 - t1; pc[2] += t3 - t2; pc[3] += t4 - t3; pc[4] += t5 - t4; pc[5] += 1;
            }
        }
    }
    if (R > 1) cluster_sync_all();   // no CTA may exit while peers can still write into its smem
}

// ---- launchers --------------------------------------------------------------------------------------
template <int R>
static cudaError_t launch_fc(int mode, const UnitDev* units, UnitState* states, const int* list, int n,
                             double omk, const u64* gkey, int sm_count, cudaStream_t st, LaunchStats* ls) {
    static int n_clusters_cached = 0;
    auto kern = k_fused_compress<R>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim         = dim3(F_THREADS);
    cfg.dynamicSmemBytes = SM_TOTAL;
    cfg.stream           = st;
    cudaLaunchAttribute attr[1];
    int grid;
    if (R > 1) {
        attr[0].id               = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = R;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs    = attr;
        cfg.numAttrs = 1;
        if (n_clusters_cached == 0) {
            cfg.gridDim = dim3(R * (sm_count / R));
            int nc = 0;
            e = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
            if (e != cudaSuccess) return e;
            if (nc < 1) return cudaErrorLaunchOutOfResources;
            n_clusters_cached = nc;
        }
        int nc = n_clusters_cached < n ? n_clusters_cached : n;
        grid = nc * R;
    } else {
        grid = sm_count < n ? sm_count : n;
    }
    cfg.gridDim = dim3(grid);
    ls->begin(R == 1 ? KID_FUSED_C1 : KID_FUSED_C8, st);
    e = cudaLaunchKernelEx(&cfg, kern, units, states, list, n, omk, gkey, mode);
    ls->end(st);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

cudaError_t launch_fused_compress(int cluster, int mode, const UnitDev* units, UnitState* states,
                                  const int* unit_list, int n_list, double one_minus_keep,
                                  const u64* global_key, int sm_count, cudaStream_t st,
                                  LaunchStats* ls) {
    if (n_list <= 0) return cudaSuccess;
    if (cluster == 1)
        return launch_fc<1>(mode, units, states, unit_list, n_list, one_minus_keep, global_key, sm_count, st, ls);
    if (cluster == 8)
        return launch_fc<8>(mode, units, states, unit_list, n_list, one_minus_keep, global_key, sm_count, st, ls);
    return cudaErrorInvalidValue;
}

// debug: sums over CTAs of the phase cycle counters; reset = zero them afterwards
cudaError_t debug_phase_cycles(unsigned long long out[6], bool reset) {
    static unsigned long long h[1024][6];
    cudaError_t e = cudaMemcpyFromSymbol(h, g_phase_cycles, sizeof(h));
    if (e != cudaSuccess) return e;
    for (int p = 0; p < 6; ++p) out[p] = 0;
    for (int c = 0; c < 1024; ++c)
        for (int p = 0; p < 6; ++p) out[p] += h[c][p];
    {
        static unsigned long long ha[1024][4];
        cudaMemcpyFromSymbol(ha, g_a_cycles, sizeof(ha));
        unsigned long long t[4] = {0, 0, 0, 0};
        for (int c = 0; c < 1024; ++c) for (int p = 0; p < 4; ++p) t[p] += ha[c][p];
        if (t[3]) fprintf(stderr, "[phaseA warp0] per chunk: gen spin %.0f, full wait %.0f, transform %.0f cycles (%llu chunks)\n",
                          (double)t[0] / t[3], (double)t[1] / t[3], (double)t[2] / t[3], t[3]);
        static unsigned long long za[1024][4];
        if (reset) cudaMemcpyToSymbol(g_a_cycles, za, sizeof(za));
    }
    if (reset) {
        static unsigned long long z[1024][6];
        e = cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
    }
    return e;
}

cudaError_t launch_fused_decompress(int, const DecUnitDev*, const InvUnitDev*, const int*, int, int*,
                                    int, cudaStream_t, LaunchStats*) {
    return cudaSuccess;
}

} // namespace wc
