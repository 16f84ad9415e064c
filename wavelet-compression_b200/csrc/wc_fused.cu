// Fused on-chip kernels for sm_100a: one unit's coefficients never leave the SM(s).
//
// COMPRESS  k_fused_compress<R, CAP, NT, STATIC>: a unit (box, component) is owned by a cluster of R CTAs
//   (R = 1 or 8), each holding CAP coefficients in shared memory.  CTA r takes the block-rows b in
//   [r*nb, (r+1)*nb) (a y-slab), so
//     - its input is, per z-plane, ONE contiguous piece of 2*nb rows, read once with coalesced 16-byte
//       streaming loads (L2 evict-first), while the same slab of the CTA's NEXT unit is prefetched into
//       L2 (evict-last) so that HBM keeps streaming during the packing phases;
//     - its coefficients are the rows j' in {b, hy+b} of every i': 2*X "segments" of nb*Z values that
//       are contiguous both in the CTA's shared-memory array C and in the global f order, so the
//       ordered (run,value) packing only needs per-segment counts from the other CTAs (DSMEM push
//       + cluster-scope mbarrier), never their coefficients.
//   Phases per unit: (A) every thread narrows + transforms TWO c-adjacent 2x2x2 blocks at once (packed
//   f32x2 add/fma/mul) and stores the 16 coefficients into C in f order (padded: conflict-free 8-byte
//   stores), keeping a running max of +c and -c; (B) threshold of src/compressor.cpp:212-216;
//   (C1) per-segment count / last-kept with float4 reads; scan of the segment table (cluster-wide for
//   R = 8), also written out as the decoder's segment table; (C2) ballot-ranked emission of (run, value)
//   pairs straight to the unit's slot in HBM.
//   HBM traffic per unit = 8N (or 4N) in + 8K out: the algorithmic minimum of SURVEY.md §8d.
//   The per-unit body (fc_unit) is a template over the geometry: FGeom (registers, 512 threads) or an
//   SGeom<...> of literals for the 32^3 / 64^3 cubes (STATIC kernels: < 64 registers, 1024 threads).
//   Unit descriptors are staged two units ahead through shared memory (FLookahead); single-CTA kernels
//   take their units from a global counter (dynamic hand-out).
//
// DECOMPRESS  k_fused_decompress<S, NT, STATIC> (+ k_seg_index): one CTA per (unit, y-slab), no clusters;
//   see the comment block in front of FastDiv below.
//
//   History: a warp-specialised TMA ring (cp.async.bulk + mbarriers) fed phase A at
//   first; with C taking 128 KB of the SM's shared memory the ring was too shallow to cover the TMA round
//   trip and it measured slower than direct loads + L2 prefetch.  The first decompress used 8-CTA clusters
//   with a DSMEM scatter; the segment-table design replaced it.
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "wc_common.cuh"
#include "wc_fused.h"

namespace wc {

// ---- compile-time geometry ---------------------------------------------------------------------
// Kernel variants (template parameters of k_fused_compress):
//   <R = 1, CAP = 32768, NT = 512>  units up to 32768 cells (32^3): one CTA per unit, 1 CTA per SM
//   <R = 8, CAP = 32768, NT = 512>  units up to 262144 cells (64^3): 8-CTA cluster, 1 CTA per SM
constexpr int F_PAD    = 4;      // padding words per i' slab of C (keeps float4/float2 alignment and makes
                                 // the a-lanes of a warp hit banks 4a + 2cp + {0,1})
constexpr int F_CPAD   = 256;    // total padding words of C (F_PAD * X, X <= 64)
constexpr int F_MAXSEG = 128;    // segments (2*X) per CTA
constexpr int F_MAXSEG_BIG = 256;   // ... of the runtime-slab-count decompress class (nx <= 128)
constexpr int F_CPAD_BIG   = 512;

template <int R, int CAP>
struct FSmem {
    static constexpr int MAXG  = F_MAXSEG * R;                       // gathered segment entries (2*X*R)
    static constexpr int C     = 0;
    static constexpr int G     = C + (CAP + F_CPAD) * 4;             // [2][MAXG] u32: cnt << 16 | last
    static constexpr int BASE  = G + 2 * MAXG * 4;                   // [F_MAXSEG] int
    static constexpr int PREV  = BASE + F_MAXSEG * 4;                // [F_MAXSEG] int
    static constexpr int RED   = PREV + F_MAXSEG * 4;                // scratch: 64 x 8 bytes
    static constexpr int XS1   = RED + 64 * 8;                       // [2][8] u64 exchange slots
    static constexpr int XS2   = XS1 + 16 * 8;                       // [2][8] u64
    static constexpr int BARS  = XS2 + 16 * 8;                       // x1 x2 x3
    static constexpr int DESC  = BARS + 4 * 8;                       // [2] FDesc: unit descriptors, one unit ahead
    static constexpr int TOTAL = DESC + 2 * 64;
};
static_assert(FSmem<8, 32768>::TOTAL <= 232448, "exceeds the 227 KB per-CTA shared memory of sm_100");

struct FGeom {
    static constexpr bool is_static = false;
    int X, Y, Z, hx, hy, hz, es;
    int nb;        // block-rows (y) per CTA
    int ncq;       // c-quads: groups of two c-adjacent block pairs
    int npairs;    // pair slots per CTA: 2 * ncq * hx * nb (a slot past hz/2 pairs is empty)
    int seglen;    // nb * Z
    int nseg;      // 2 * X
    int nlocal;    // X * 2 * nb * Z
    int slab;      // padded words of C per i'
};

// The same quantities as literals, for the shapes AMR codes produce most (cubes of 32 and 64).
template <int X_, int Y_, int Z_, int ES_, int R_>
struct SGeom {
    static constexpr bool is_static = true;
    static constexpr int X = X_, Y = Y_, Z = Z_, hx = X_ / 2, hy = Y_ / 2, hz = Z_ / 2, es = ES_;
    static constexpr int nb     = hy / R_;
    static constexpr int ncq    = (hz / 2 + 1) / 2;
    static constexpr int npairs = 2 * ncq * hx * nb;
    static constexpr int seglen = nb * Z_;
    static constexpr int nseg   = 2 * X_;
    static constexpr int nlocal = nseg * seglen;
    static constexpr int slab   = 2 * nb * Z_ + F_PAD;
};

__host__ __device__ inline bool fused_geom(int X, int Y, int Z, int dtype, int R, int cap, FGeom& g, int maxseg = F_MAXSEG) {
    if (X < 2 || Y < 2 || Z < 4 || (X & 1) || (Y & 1) || (Z & 3) || R < 1) return false;
    g.X = X; g.Y = Y; g.Z = Z;
    g.hx = X / 2; g.hy = Y / 2; g.hz = Z / 2;
    g.es = dtype == WC_F64 ? 8 : 4;
    if ((X * g.es) % 16) return false;       // 16-byte vector loads of (x, x+1) pairs, row-aligned
    if (2 * X > maxseg) return false;
    long long n = (long long)X * Y * Z;
    if (g.hy % R) return false;
    if (n / R > cap) return false;
    g.nb     = g.hy / R;
    g.ncq    = (g.hz / 2 + 1) / 2;
    g.npairs = 2 * g.ncq * g.hx * g.nb;
    g.seglen = g.nb * Z;
    g.nseg   = 2 * X;
    g.nlocal = g.nseg * g.seglen;
    g.slab   = 2 * g.nb * Z + F_PAD;
    return true;
}

// Slab count of a box the cluster classes do not take (more than 262144 cells, or a half-height no cluster of 2 / 4 / 8
// divides): the fewest y-slabs of at most 32768 cells, 0 if the box has no such decomposition (odd dimension, nz not a
// multiple of 4, nx > 128, a single block-row above 32768 cells).  Decompress only needs independent slab items
// (FUSED_CLS_RBIG); compress has no cluster that large.
__host__ __device__ inline int big_slabs(int nx, int ny, int nz) {
    if (nx < 2 || ny < 2 || nz < 4 || (nx & 1) || (ny & 1) || (nz & 3) || 2 * nx > F_MAXSEG_BIG) return 0;
    const long long row = 2ll * nx * nz;                // cells of one block-row
    if (row > 32768) return 0;
    const int hy = ny / 2;
    int nb = (int)(32768 / row);
    if (nb > hy) nb = hy;
    while (hy % nb) --nb;
    const int S = hy / nb;
    return (S >= 2 && S <= 1024) ? S : 0;
}

int big_run_key(int nx, int ny, int nz) {
    const int S = big_slabs(nx, ny, nz);
    return S | ((nx == 128 && ny == 128 && nz == 128) ? (int)BIG_CUBE128 : 0);
}

int fused_class(int nx, int ny, int nz, int dtype, const void* ptr) {
    // the x-slab classes (wc_xslab.cu) only need element alignment: they take what the y-slab classes refuse
    const bool elem_aligned = (reinterpret_cast<uintptr_t>(ptr) & (dtype == WC_F64 ? 7u : 3u)) == 0;
    if (reinterpret_cast<uintptr_t>(ptr) & 15u) return elem_aligned ? xs_class_of(nx, ny, nz) : FUSED_CLS_NONE;
    FGeom g;
    if (nx == 32 && ny == 32 && nz == 32) return FUSED_CLS_CUBE32;
    if (nx == 64 && ny == 64 && nz == 64) return FUSED_CLS_CUBE64;
    if (nx == 16 && ny == 16 && nz == 16) return FUSED_CLS_CUBE16;
    if (nx == 8 && ny == 8 && nz == 8) return FUSED_CLS_CUBE8;
    if (fused_geom(nx, ny, nz, dtype, 1, 4096, g)) return FUSED_CLS_R1S;
    if (fused_geom(nx, ny, nz, dtype, 1, 32768, g)) return FUSED_CLS_R1;
    if (fused_geom(nx, ny, nz, dtype, 2, 32768, g)) return FUSED_CLS_R2;
    if (fused_geom(nx, ny, nz, dtype, 4, 32768, g)) return FUSED_CLS_R4;
    if (fused_geom(nx, ny, nz, dtype, 8, 32768, g)) return FUSED_CLS_R8;
    if (big_forward_slabs(nx, ny, nz, dtype, ptr)) return FUSED_CLS_NONE;     // two passes over a scratch (k_big_forward)
    return xs_class_of(nx, ny, nz);
}

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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
__device__ __forceinline__ bool mbar_try_wait_cta(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cta(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait_cta(bar, parity)) { }
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

// exact q / d for q < 65536, d < 65536 (m = ceil(2^32 / d); d == 1 is encoded as m == 0)
__device__ __forceinline__ uint32_t fdiv(uint32_t q, uint32_t m) { return m ? __umulhi(q, m) : q; }
// index of the most significant set bit, 0xffffffff for 0
__device__ __forceinline__ uint32_t bfind_u32(uint32_t x) {
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
}
// the same through the float exponent: I2FP + shift + add on the ALU/FMA pipes instead of one FLO on the
// (16 lanes/clk) XU pipe, which the emission phase otherwise loads with four ops per 32 coefficients
__device__ __forceinline__ int bfind_alu(uint32_t x) {
    return (int)(__float_as_uint(__uint2float_rz(x)) >> 23) - 127;     // x == 0 -> -127 (callers test x first)
}
__device__ __forceinline__ uint32_t fdiv_magic(uint32_t d) { return d <= 1 ? 0u : (0xffffffffu / d) + 1u; }

// Streaming accesses carry an L2 evict-first policy: the input is read once and the pairs are written
// once, so neither should push the NEXT unit's prefetched lines (evict-last) out of L2.
__device__ __forceinline__ u64 l2_policy_evict_first() {
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double2 ldg_stream_f64x2(const void* p, u64 pol) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float2 ldg_stream_f32x2(const void* p, u64 pol) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;"
                 : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
    return v;
}
template <class T> __device__ __forceinline__ T ldg_stream(const void* p, u64 pol);
template <> __device__ __forceinline__ double2 ldg_stream<double2>(const void* p, u64 pol) { return ldg_stream_f64x2(p, pol); }
template <> __device__ __forceinline__ float2  ldg_stream<float2>(const void* p, u64 pol) { return ldg_stream_f32x2(p, pol); }
__device__ __forceinline__ void st_pair_pred(bool p, int2* addr, int run, float val, u64 pol) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.u32 q, %0, 0;\n\t"
        "@q st.global.L2::cache_hint.v2.b32 [%1], {%2, %3}, %4;\n\t}"
        ::"r"((uint32_t)p), "l"(addr), "r"(run), "r"(__float_as_int(val)), "l"(pol) : "memory");
}

// ---- the kernel -----------------------------------------------------------------------------------
// Per-CTA phase cycle counters (clock64 by thread 0 at the phase boundaries): a built-in light-weight
// profile, read back with wc_debug_phase_cycles().  [cta][phase], phases: 0 = A (load + transform),
// 1 = B (threshold, includes waiting for the slowest warp of A), 2 = C1 (count), 3 = scan (+ cluster
// exchange), 4 = C2 (emit), 5 = units processed.
// Only in builds with -DWC_PHASE_PROFILE (make PHASE_PROFILE=1): the read-modify-writes cost thread 0 some
// 600 cycles per unit.
#ifdef WC_HANG_DEBUG
// Developer aid (never in the product build): per-warp progress markers in mapped host memory, readable while a kernel hangs.
__device__ int* g_dbg_ptr;
#define WC_MARK(s) do { if (g_dbg_ptr && (threadIdx.x & 31) == 0) { ((volatile int*)g_dbg_ptr)[blockIdx.x * 32 + (threadIdx.x >> 5)] = (s); __threadfence_system(); } } while (0)
extern "C" __attribute__((visibility("default"))) int wc_debug_set_marker(void* p) {
    return (int)cudaMemcpyToSymbol(g_dbg_ptr, &p, sizeof(p));
}
#else
#define WC_MARK(s) do { } while (0)
#endif
#ifdef WC_PHASE_PROFILE
__device__ unsigned long long g_phase_cycles[1024][8];
#define WC_PHASE_CLOCK(t) long long t = clock64()
#else
#define WC_PHASE_CLOCK(t) do { } while (0)
#endif

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

// Phase A works on the two c-adjacent 2x2x2 blocks whose first cell is at `p0`: 4 z-planes x 2 rows x one
// (x,x+1) pair each, straight from global memory / L2 with 16-byte (f64) or 8-byte (f32) vector loads.
// It is split in two so that the loop can issue the loads of slot q+NT right after narrowing slot q
// (the raw registers are free from that point on) and run the transform of slot q under their latency.
template <int ES> struct RawPair;
template <> struct RawPair<8> { double2 d[8]; };
template <> struct RawPair<4> { float2 d[8]; };

template <int ES>
__device__ __forceinline__ void load_pair(RawPair<ES>& r, const char* p0, size_t plane_bytes, size_t row_bytes,
                                          u64 pol) {
#pragma unroll
    for (int pl = 0; pl < 4; ++pl)
#pragma unroll
        for (int yi = 0; yi < 2; ++yi) {
            if (ES == 8) r.d[pl * 2 + yi] = ldg_stream<typename std::remove_reference<decltype(r.d[0])>::type>(p0 + pl * plane_bytes + yi * row_bytes, pol);
            else         r.d[pl * 2 + yi] = ldg_stream<typename std::remove_reference<decltype(r.d[0])>::type>(p0 + pl * plane_bytes + yi * row_bytes, pol);
        }
}
// v[zi*4+yi*2+xi] = (block c, block c+1), narrowed as src/preprocess.cpp:78 does
__device__ __forceinline__ void narrow_pair(const RawPair<8>& r, float2 v[8]) {
#pragma unroll
    for (int zi = 0; zi < 2; ++zi)
#pragma unroll
        for (int yi = 0; yi < 2; ++yi) {
            const double2 da = r.d[zi * 2 + yi], db = r.d[(zi + 2) * 2 + yi];
            v[zi * 4 + yi * 2]     = make_float2(__double2float_rn(da.x), __double2float_rn(db.x));
            v[zi * 4 + yi * 2 + 1] = make_float2(__double2float_rn(da.y), __double2float_rn(db.y));
        }
}
__device__ __forceinline__ void narrow_pair(const RawPair<4>& r, float2 v[8]) {
#pragma unroll
    for (int zi = 0; zi < 2; ++zi)
#pragma unroll
        for (int yi = 0; yi < 2; ++yi) {
            const float2 fa = r.d[zi * 2 + yi], fb = r.d[(zi + 2) * 2 + yi];
            v[zi * 4 + yi * 2]     = make_float2(fa.x, fb.x);
            v[zi * 4 + yi * 2 + 1] = make_float2(fa.y, fb.y);
        }
}
// Transforms both blocks at once and stores the 8 x 2 coefficients into C.
__device__ __forceinline__ void finish_pair(float2 v[8], float* cdst, int o1, int o2, int o3, float& bp, float& bn) {
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
}

// Phase A of one CTA: every thread takes the slots q = tid, tid + NT, ... of its y-slab.
//   slot q -> (cp2 fastest, a, cq, bl): the lanes of a warp are 2 c-pairs x 16 a  ->  coalesced rows, and
//   conflict-free 8-byte stores into C (banks 4a + 2cp2 + {0,1}).
// With literal geometry the trip count is a constant: the loop is fully unrolled and the loads of slot
// i+1 are issued between the narrowing and the transform of slot i (software pipeline without extra
// registers, and without a branch ptxas could hoist the math over).
// MM: also track min / max of the narrowed INPUT values (ingest statistics, src/preprocess.cpp:82-88; fminf /
// fmaxf skip NaNs exactly like the reference's `value < min` / `value > max` updates).
__device__ __forceinline__ void minmax_pair(const float2 v[8], float& mn, float& mx) {
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        mn = fminf(fminf(mn, v[o].x), v[o].y);
        mx = fmaxf(fmaxf(mx, v[o].x), v[o].y);
    }
}
template <int NT, int ES, bool MM, class G>
__device__ __forceinline__ void phase_a(const G& g, const char* in0, float* C, uint32_t rank, u64 pol,
                                        float& bp, float& bn, bool& nan0, float& vmn, float& vmx) {
    const int tid = threadIdx.x;
    const size_t row_bytes = (size_t)g.X * g.es, plane_bytes = row_bytes * g.Y;
    const int o1 = g.hx * g.slab, o2 = g.nb * g.Z, o3 = g.hz;
    const uint32_t m_hx = G::is_static ? 0u : fdiv_magic(g.hx), m_cq = G::is_static ? 0u : fdiv_magic(g.ncq);
    const int npc = g.hz >> 1;                         // c-pairs per (a, b)
    auto decode = [&](int q, const char*& p0, float*& cdst) -> bool {
        const uint32_t cp2 = q & 1, t1 = q >> 1;
        const uint32_t t2 = G::is_static ? t1 / (uint32_t)g.hx : fdiv(t1, m_hx), a = t1 - t2 * g.hx;
        const uint32_t bl = G::is_static ? t2 / (uint32_t)g.ncq : fdiv(t2, m_cq), cq = t2 - bl * g.ncq;
        const int cpi = 2 * cq + cp2;                  // c-pair index: blocks c = 2cpi, 2cpi+1
        p0   = in0 + (size_t)(4 * cpi) * plane_bytes + (size_t)(2 * bl) * row_bytes + (size_t)a * 2 * g.es;
        cdst = C + a * g.slab + bl * g.Z + 2 * cpi;
        return !(npc & 1) || cpi < npc;                // hz/2 odd: the last quad has one pair
    };
    RawPair<ES> raw;
    const char* p0;
    float*      cdst;
    if constexpr (G::is_static) {
        static_assert(G::npairs % NT == 0 && (G::hz / 2) % 2 == 0, "literal geometries fill every slot");
        constexpr int NIT = G::npairs / NT;
        decode(tid, p0, cdst);
        load_pair<ES>(raw, p0, plane_bytes, row_bytes, pol);
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            float2 v[8];
            narrow_pair(raw, v);
            if (MM) minmax_pair(v, vmn, vmx);
            float* const cd = cdst;
            if (it + 1 < NIT) {
                decode(tid + (it + 1) * NT, p0, cdst);
                load_pair<ES>(raw, p0, plane_bytes, row_bytes, pol);
            }
            finish_pair(v, cd, o1, o2, o3, bp, bn);
            if (it == 0 && tid == 0 && rank == 0) nan0 = isnan(v[0].x);   // block (0,0,0): coefficient f = 0
        }
    } else {
#pragma unroll 1
        for (int q = tid; q < g.npairs; q += NT) {
            if (!decode(q, p0, cdst)) continue;
            float2 v[8];
            load_pair<ES>(raw, p0, plane_bytes, row_bytes, pol);
            narrow_pair(raw, v);
            if (MM) minmax_pair(v, vmn, vmx);
            finish_pair(v, cdst, o1, o2, o3, bp, bn);
            if (q == 0 && rank == 0) nan0 = isnan(v[0].x);
        }
    }
}

// Unit descriptors travel through shared memory one unit ahead: while unit k is processed, thread 0
// fetches (index ->) unit id -> UnitDev of unit k+2 in stages placed at the phase boundaries, so that the
// dependent global loads (and the work-counter atomic of the dynamic hand-out) never sit on the critical
// path at the top of a unit.  Slot k&1 holds unit k; unit k+1 (slot (k+1)&1) is the L2-prefetch target.
struct __align__(8) FDesc {
    UnitDev u;      // 56 bytes
    int     uid;    // index into units[] / states[]
    int     ui;     // position in the work list; >= n_list: no more work
};
static_assert(sizeof(FDesc) == 64, "FDesc layout");
struct FLookahead {
    const UnitDev* units;
    const int*     unit_list;
    int*           work_counter;   // dynamic hand-out, or nullptr: static stride
    int            n_list, stride;
    FDesc*         slot;           // where unit k+2 goes (= the slot of unit k)
    int            ui_prev;        // static: list position of unit k+1
    int            idx, uid;       // thread 0 only
    int            batch, batch_next, batch_left;
    __device__ __forceinline__ void stage1() {          // top of the unit
        if (!work_counter) { idx = ui_prev + stride; return; }
        // dynamic hand-out, `batch` consecutive units per atomic: hundreds of CTAs hitting one address cost
        // ~15-30 cycles per atomic chip-wide, which would bound the small-unit kernels
        if (batch_left == 0) { batch_next = atomicAdd(work_counter, batch); batch_left = batch; }
        idx = batch_next++;
        --batch_left;
    }
    __device__ __forceinline__ void stage2() {          // after the first barrier: slot k&1 is free now
        uid = -1;          // volatile: issued HERE (the compiler would otherwise sink the load to its first use)
        if (idx < n_list) asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(uid) : "l"(unit_list + idx));
    }
    __device__ __forceinline__ void stage3() {
        slot->ui  = idx;
        slot->uid = uid;
        if (uid >= 0) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&slot->u);
            const char*    src = reinterpret_cast<const char*>(units + uid);
#pragma unroll
            for (int b = 0; b < 56; b += 8)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + b), "l"(src + b) : "memory");
        }
    }
    __device__ __forceinline__ void stage4() {          // before the unit's last barrier
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
};

// Everything the per-unit body needs from the kernel frame (all scalarised after inlining).
struct FShared {
    float*    C;
    uint32_t* g_pk;
    int*      s_base;
    int*      s_prev;
    u64*      s_red;
    u64*      xs1;
    u64*      xs2;
    int*      s_next;
    uint32_t  xb1, xb2, xb3;
};
// This CTA's slab of its NEXT unit, for the L2 prefetch: `nplanes` pieces of `piece_lines` 128-byte lines,
// `pitch` bytes apart (one piece of the whole unit when the CTA owns complete z-planes, R = 1).
struct FPrefetch {
    const char* base;
    uint32_t    piece_lines, nplanes;
    size_t      pitch;
};

// One unit, start to finish.  G is FGeom (geometry in registers, any admissible shape) or an SGeom<...>
// (the common cubes: every stride, trip count and divisor is a literal).
// order-preserving float -> uint32 (larger float = larger code; every code is > 0, so a zeroed word = unset)
__device__ __forceinline__ uint32_t float_order_code(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
template <int R, int CAP, int NT, bool MM, class G>
__device__ __forceinline__ void fc_unit(const G& g, const UnitDev& u, const int uid, const FShared& S,
                                        const FPrefetch& pf, FLookahead& la, const uint32_t rank, uint32_t& xph1,
                                        uint32_t& xph2, uint32_t& xph3, UnitState* __restrict__ states,
                                        const double one_minus_keep, const u64* __restrict__ global_key,
                                        const int mode, const u64 pol, const uint32_t lt) {
    typedef FSmem<R, CAP> SM;
    constexpr int NW = NT / 32;
    float* const C = S.C;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b0 = rank * g.nb;
    const size_t row_bytes = (size_t)g.X * g.es;
    float bp = 0.f, bn = 0.f;                 // running max of +c and of -c
    float vmn = __int_as_float(0x7f800000), vmx = __int_as_float(0xff800000);   // MM: min / max of the inputs
    bool  nan0 = false;
    WC_PHASE_CLOCK(t0);
#ifdef WC_PHASE_PROFILE
    long long wait_b = 0, wait_s = 0;      // cycles in the two cluster-scope mbarrier waits
#endif
    if (tid == 0) la.stage1();
    __syncwarp();      // see the note at the C2 loop: thread 0 must be back before the next warp collective

    // ---------------- phase A: load, narrow, transform two blocks per thread, store into C -------
    {
        const char* in0 = static_cast<const char*>(u.in) + (size_t)(2 * b0) * row_bytes;
        if (g.es == 8) phase_a<NT, 8, MM>(g, in0, C, rank, pol, bp, bn, nan0, vmn, vmx);
        else           phase_a<NT, 4, MM>(g, in0, C, rank, pol, bp, bn, nan0, vmn, vmx);
    }

    // L2 prefetch of this CTA's slab of its NEXT unit, issued after this unit's own loads: the HBM reads
    // of unit u+1 overlap the packing phases of unit u.
    if (pf.base) {
        if (R == 1) {
            for (uint32_t i = tid; i < pf.piece_lines; i += NT)
                asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(pf.base + (size_t)i * 128));
        } else {
            const uint32_t nlines = pf.piece_lines * pf.nplanes;
            for (uint32_t i = tid; i < nlines; i += NT) {
                const uint32_t z = i / pf.piece_lines, l = i - z * pf.piece_lines;
                asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(pf.base + (size_t)z * pf.pitch + (size_t)l * 128));
            }
        }
    }

    // ---------------- phase B: the threshold ----------------
    WC_PHASE_CLOCK(t1);
    WC_MARK(1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bp = fmaxf(bp, __shfl_xor_sync(0xffffffffu, bp, o));
        bn = fmaxf(bn, __shfl_xor_sync(0xffffffffu, bn, o));
    }
    const bool any_nan0 = __any_sync(0xffffffffu, nan0);
    if (MM) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vmn = fminf(vmn, __shfl_xor_sync(0xffffffffu, vmn, o));
            vmx = fmaxf(vmx, __shfl_xor_sync(0xffffffffu, vmx, o));
        }
        if (lane == 0) S.s_red[32 + warp] = ((u64)__float_as_uint(vmn) << 32) | (u64)__float_as_uint(vmx);
    }
    if (lane == 0) {
        // bn >= 0, so its sign bit is free: it carries "the coefficient at f = 0 is NaN"
        S.s_red[warp] = ((u64)__float_as_uint(bp) << 32) |
                        (u64)((__float_as_uint(bn) & 0x7fffffffu) | (any_nan0 ? 0x80000000u : 0u));
    }
    __syncthreads();
    WC_MARK(2);
    if (tid == 0) la.stage2();
    __syncwarp();
    if (MM && warp == 0) {
        // one atomic pair per CTA and unit: vmax holds the order code of the max, vmin the INVERTED code of
        // the min (both reduce with atomicMax over a zeroed word; wc_plan_unit_stats decodes them)
        const u64 x = lane < NW ? S.s_red[32 + lane] : ((u64)0x7f800000u << 32) | 0xff800000u;
        float a = __uint_as_float((uint32_t)(x >> 32)), b = __uint_as_float((uint32_t)x);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
            b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
        }
        if (lane == 0) {
            if (a <= b) {     // at least one comparable value
                atomicMax(reinterpret_cast<unsigned int*>(&states[uid].vmin), ~float_order_code(a));
                atomicMax(reinterpret_cast<unsigned int*>(&states[uid].vmax), float_order_code(b));
            }
        }
    }
    float Mp = 0.f, Mn = 0.f;
    bool  first_nan = false;
    {
        u64 x = lane < NW ? S.s_red[lane] : 0ull;
        float p = __uint_as_float((uint32_t)(x >> 32));
        float n = __uint_as_float((uint32_t)x & 0x7fffffffu);
        first_nan = __any_sync(0xffffffffu, ((uint32_t)x >> 31) != 0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
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
            st_cluster_u64(mapa(smem_u32(&S.xs1[par * 8 + rank]), tid), pay);
            mbar_arrive_remote(mapa(S.xb1, tid));
        }
        WC_PHASE_CLOCK(tw0);
        mbar_wait_cluster(S.xb1, par);
        WC_PHASE_CLOCK(tw1);
#ifdef WC_PHASE_PROFILE
        wait_b = tw1 - tw0;
#endif
        ++xph1;
    WC_MARK(3);
        float p = 0.f, n = 0.f;
        bool fn = false;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            u64 x = S.xs1[par * 8 + r];
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
#pragma unroll 1
        for (int l = tid; l < g.nlocal; l += NT) {
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
        __syncthreads();   // s_red reuse
        if (lane == 0) S.s_red[warp] = best;
        __syncthreads();
        best = lane < NW ? S.s_red[lane] : ~0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            u64 x = __shfl_xor_sync(0xffffffffu, best, o);
            best = x < best ? x : best;
        }
        if (R > 1) {
            const uint32_t par = xph3 & 1;
            if (tid < R) {
                st_cluster_u64(mapa(smem_u32(&S.xs2[par * 8 + rank]), tid), best);
                mbar_arrive_remote(mapa(S.xb3, tid));
            }
            mbar_wait_cluster(S.xb3, par);
            ++xph3;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                u64 x = S.xs2[par * 8 + r];
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
        if (tid == 0) { la.stage3(); la.stage4(); }
        __syncthreads();   // C is rewritten by the next unit
        return;
    }

    // ---------------- phase C1: per-segment count and last kept ----------------
    WC_PHASE_CLOCK(t2);
    WC_MARK(4);
    const int gpar = R > 1 ? (int)(xph2 & 1) : 0;
    uint32_t* const my_pk = S.g_pk + gpar * SM::MAXG;
#pragma unroll 1
    for (int sg = warp; sg < g.nseg; sg += NW) {
        const float* cs = C + sg * g.seglen + F_PAD * (sg >> 1);   // 16-byte aligned
        // 512 coefficients per round: four float4 per lane, one 16-bit keep mask, one POPC
        const bool whole512 = (g.seglen & 511) == 0;
        int cnt = 0, lw = 0;
        uint32_t lm = 0;
#pragma unroll 1
        for (int w0 = lane * 4; w0 < g.seglen; w0 += 512) {
            uint32_t m = 0;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int w = w0 + 128 * it;
                if (whole512 || w < g.seglen) {                      // seglen % 4 == 0
                    const float4 c = *reinterpret_cast<const float4*>(cs + w);
                    m |= ((keep_coef(c.x, tf) ? 1u : 0u) | (keep_coef(c.y, tf) ? 2u : 0u) |
                          (keep_coef(c.z, tf) ? 4u : 0u) | (keep_coef(c.w, tf) ? 8u : 0u)) << (4 * it);
                }
            }
            cnt += __popc(m);
            if (m) { lw = w0; lm = m; }
        }
        int last = -1;
        if (lm) {
            const int hb = (int)bfind_u32(lm);                       // bit 4*it + k  ->  w = lw + 128*it + k
            last = lw + 128 * (hb >> 2) + (hb & 3);
        }
        cnt  = __reduce_add_sync(0xffffffffu, cnt);
        last = __reduce_max_sync(0xffffffffu, last);
        const uint32_t pk = ((uint32_t)cnt << 16) | ((uint32_t)last & 0xffffu);
        if (R == 1) {
            if (lane == 0) my_pk[sg] = pk;
        } else if (lane < R) {
            st_cluster_u32(mapa(smem_u32(&my_pk[sg * R + rank]), lane), pk);
        }
    }
    WC_PHASE_CLOCK(t3);
    WC_MARK(5);
    if (R > 1) {
#ifndef WC_C1_FENCE
#define WC_C1_FENCE 1
#endif
#if WC_C1_FENCE
        fence_cluster();
#endif
        __syncthreads();
        if (tid < R) mbar_arrive_remote(mapa(S.xb2, tid));
        WC_PHASE_CLOCK(tw2);
        mbar_wait_cluster(S.xb2, gpar);
        WC_PHASE_CLOCK(tw3);
#ifdef WC_PHASE_PROFILE
        wait_s = tw3 - tw2;
#endif
        ++xph2;
    WC_MARK(6);
    } else {
        __syncthreads();
    }

    // ---------------- scan over the segments in global order ----------------
    {
        // entries per thread: a thread of a cluster kernel owns ONE segment with the entries of all R CTAs (they are
        // adjacent in the global order), so at most four warps scan and every thread reads 2 * NWA warp totals
        // instead of 2 * NW (one entry per thread kept all 32 warps of the 64^3 kernel busy with 64 shared-memory
        // reads each: ~2 k cycles per slab)
        // entries per thread: a thread of a cluster kernel owns ONE segment with the entries of all R CTAs (they are
        // adjacent in the global order), so at most four warps scan and every thread reads 2 * NWA warp totals
        // instead of 2 * NW (one entry per thread kept all 32 warps of the 64^3 kernel busy with 64 shared-memory
        // reads each).  Nothing but (isum, imax) lives across the barrier: the entries are read again behind it —
        // the kernel sits at its 64-register cap, and a version that kept them spilled to local memory.
        constexpr int EPT = R > 1 ? R : (SM::MAXG + NT - 1) / NT;
        constexpr int NWA = (SM::MAXG / EPT + 31) / 32 < NW ? (SM::MAXG / EPT + 31) / 32 : NW;   // warps that own entries
        const int NG = g.nseg * R;                      // <= MAXG, entry e = sg * R + r
        int* sr = reinterpret_cast<int*>(S.s_red);
        // only the warps that own entries work; the others (all but the first for R = 1) just meet the barriers
        const bool active = warp * 32 * EPT < NG;
        // entry e -> (kept count, flat index of its last kept coefficient or -1)
        auto entry = [&](int e, int& cnt, int& last) {
            cnt = 0; last = -1;
            if (e < NG) {
                const uint32_t pk = my_pk[e];
                const int sg = e / R, r = e % R;
                cnt = (int)(pk >> 16);
                if ((pk & 0xffffu) != 0xffffu)
                    last = ((sg >> 1) * g.Y + (sg & 1) * g.hy + r * g.nb) * g.Z + (int)(pk & 0xffffu);
            }
        };
        int isum = 0, imax = -1;
        if (active) {
#pragma unroll
            for (int j = 0; j < EPT; ++j) {
                int c, l;
                entry(tid * EPT + j, c, l);
                isum += c;
                imax = max(imax, l);
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int ps = __shfl_up_sync(0xffffffffu, isum, o);
                int pm = __shfl_up_sync(0xffffffffu, imax, o);
                if (lane >= o) { isum += ps; imax = max(imax, pm); }
            }
        }
        if (lane == 31 && warp < NWA) { sr[warp] = isum; sr[32 + warp] = imax; }
        __syncthreads();
    WC_MARK(7);
        if (active) {
            int wsum = 0, wmax = -1, total = 0;
#pragma unroll
            for (int i = 0; i < NWA; ++i) {
                int xs = sr[i], xm = sr[32 + i];
                if (i < warp) { wsum += xs; wmax = max(wmax, xm); }
                total += xs;
            }
            int es = __shfl_up_sync(0xffffffffu, isum, 1);
            int em = __shfl_up_sync(0xffffffffu, imax, 1);
            if (lane == 0) { es = 0; em = -1; }
            es += wsum;
            em = max(em, wmax);
#pragma unroll
            for (int j = 0; j < EPT; ++j) {                 // exclusive prefix in front of entry e
                const int e = tid * EPT + j;
                if (e < NG && (R == 1 || (e % R) == (int)rank)) {
                    S.s_base[e / R] = es; S.s_prev[e / R] = em;
                    // decode-side segment table (k_seg_index): first pair of the segment, last kept before it
                    if (u.coef) reinterpret_cast<int2*>(u.coef)[e] = make_int2(es, em);
                }
                if (j + 1 < EPT) {
                    int c, l;
                    entry(e, c, l);
                    es += c;
                    em = max(em, l);
                }
            }
            if (tid == 0 && rank == 0) {
                states[uid].npairs = total;
                states[uid].flags  = (first_nan ? UNIT_FLAG_NAN0 : 0) |
                                     ((total > 0 && unit_need32(M, tf)) ? UNIT_FLAG_NEED32 : 0);
                if (u.coef) reinterpret_cast<int2*>(u.coef)[NG] = make_int2(total, -1);   // sentinel
            }
            if (tid == 0) *S.s_next = 0;
        }
        __syncthreads();
    }

    // ---------------- phase C2: emit (run, value) pairs ----------------
    // Segments are handed out dynamically (shared-memory counter): their cost ranges from nothing
    // (detail bands below the threshold) to a full copy, and a static round-robin left half of the
    // warps idle at the closing barrier.
    WC_PHASE_CLOCK(t4);
    WC_MARK(8);
    if (tid == 0) la.stage3();
    // Thread 0 has to be back in its warp before the loop's first __shfl_sync.  ptxas 12.9 does not always put a
    // reconvergence point behind this one-thread region (it sits inside a larger region that spans CTA barriers), and
    // it issues the loop's SHFL without a WARPSYNC: lanes 1..31 of warp 0 then read the segment number from an absent
    // lane 0, never see it reach nseg and spin forever — the hang of DESIGN.md §4.5.  Which builds are hit depends on
    // register allocation and block layout around the scan, not on the scan's arithmetic.
    __syncwarp();
    int2* const out = reinterpret_cast<int2*>(u.out);
    for (;;) {
        int sg = 0;
        if (lane == 0) sg = atomicAdd(S.s_next, 1);
        sg = __shfl_sync(0xffffffffu, sg, 0);
        if (sg >= g.nseg) break;
        const int scnt = (int)(my_pk[sg * R + rank] >> 16);
        if (scnt == 0) continue;                                   // nothing kept in this segment
        const float* cs = C + sg * g.seglen + F_PAD * (sg >> 1);
        const int fstart = ((sg >> 1) * g.Y + (sg & 1) * g.hy + b0) * g.Z;
        uint32_t pos = (uint32_t)S.s_base[sg];
        int prev = S.s_prev[sg];
        if (scnt == g.seglen) {
            // every coefficient kept (e.g. a negative max, SURVEY.md D3'): runs are 0, ranks are w
            for (int w = lane; w < g.seglen; w += 32)
                st_pair_pred(true, out + (pos + (uint32_t)w), w == 0 ? fstart - prev - 1 : 0, cs[w], pol);
            continue;
        }
        const bool whole = (g.seglen & 127) == 0;                  // no ragged last group
#pragma unroll 1
        for (int w0 = 0; w0 < g.seglen; w0 += 128) {
            float    c[4];
            uint32_t bal[4];
            bool     kf[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int w = w0 + 32 * j + lane;
                const bool ok = whole || w < g.seglen;
                c[j]   = ok ? cs[w] : 0.f;
                kf[j]  = ok && keep_coef(c[j], tf);
                bal[j] = __ballot_sync(0xffffffffu, kf[j]);
            }
            // prel = (flat index of the previous kept coefficient) - (flat index of this group's lane 0)
            int prel = prev - (fstart + w0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (bal[j] != 0u) {                                    // warp-uniform
                    const uint32_t lower = bal[j] & lt;
                    const int pl = lower ? bfind_alu(lower) : prel;        // previous kept, lane units
                    st_pair_pred(kf[j], out + (pos + __popc(lower)), lane - pl - 1, c[j], pol);
                    pos += __popc(bal[j]);
                    prel = bfind_alu(bal[j]);
                }
                prel -= 32;
            }
            prev = prel + fstart + w0 + 128;
        }
    }
    WC_MARK(9);
    if (tid == 0) la.stage4();
    __syncthreads();   // C and the segment arrays are rewritten by the next unit
    WC_MARK(10);
#ifdef WC_PHASE_PROFILE
    if (tid == 0 && blockIdx.x < 1024) {
        long long t5 = clock64();
        unsigned long long* pc = g_phase_cycles[blockIdx.x];
        pc[0] += t1 - t0; pc[1] += t2 - t1; pc[2] += t3 - t2; pc[3] += t4 - t3; pc[4] += t5 - t4; pc[5] += 1;
        pc[6] += wait_b; pc[7] += wait_s;
    }
#endif
}

// STATIC: every unit of the list is the cube this variant is specialised for (32^3 for R = 1, 64^3 for R = 8).
template <int R, int CAP, int NT, bool STATIC, bool MM>
__global__ void __launch_bounds__(NT, (CAP <= 512 ? 32 : CAP <= 4096 ? 4 : 1))
k_fused_compress(const UnitDev* __restrict__ units, UnitState* __restrict__ states,
                 const int* __restrict__ unit_list, int n_list, double one_minus_keep,
                 const u64* __restrict__ global_key, int mode, int* __restrict__ work_counter) {
    extern __shared__ __align__(128) unsigned char smem[];
    typedef FSmem<R, CAP> SM;
    FShared S;
    S.C      = reinterpret_cast<float*>(smem + SM::C);
    S.g_pk   = reinterpret_cast<uint32_t*>(smem + SM::G);
    S.s_base = reinterpret_cast<int*>(smem + SM::BASE);
    S.s_prev = reinterpret_cast<int*>(smem + SM::PREV);
    S.s_red  = reinterpret_cast<u64*>(smem + SM::RED);
    S.xs1    = reinterpret_cast<u64*>(smem + SM::XS1);
    S.xs2    = reinterpret_cast<u64*>(smem + SM::XS2);
    S.s_next = reinterpret_cast<int*>(smem + SM::RED + 48 * 8);   // C2 work counter
    const uint32_t bars = smem_u32(smem + SM::BARS);
    S.xb1 = bars; S.xb2 = bars + 8; S.xb3 = bars + 16;

    const int tid = threadIdx.x;
    const uint32_t rank = R > 1 ? cluster_ctarank() : 0u;
    const uint32_t cid  = R > 1 ? cluster_id_x() : blockIdx.x;
    const uint32_t ncl  = R > 1 ? nclusters_x() : gridDim.x;

    if (R > 1) {
        if (tid == 0) {
            mbar_init(S.xb1, R);
            mbar_init(S.xb2, R);
            mbar_init(S.xb3, R);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        cluster_sync_all();
    }

    WC_MARK(20);
    const uint32_t lt = lanemask_lt();
    const u64 pol = l2_policy_evict_first();
    uint32_t xph1 = 0, xph2 = 0, xph3 = 0;
    // Unit hand-out: static round-robin over the clusters, or (R = 1, work_counter given) a global atomic
    // counter, so that CTAs that become resident late — e.g. on SMs the 8-CTA-cluster kernel of the same
    // step is still using — simply take fewer units.  Descriptors are staged two units ahead (FLookahead).
    const bool dynamic = (R == 1) && work_counter != nullptr;
    FDesc* const s_desc = reinterpret_cast<FDesc*>(smem + SM::DESC);
    FLookahead la;
    la.units = units; la.unit_list = unit_list; la.work_counter = dynamic ? work_counter : nullptr;
    la.n_list = n_list; la.stride = (int)ncl;
    la.idx = 0; la.uid = -1;
    la.batch = CAP <= 4096 ? 16 : 1; la.batch_next = 0; la.batch_left = 0;
    if (tid == 0) {
        // prologue: units 0 and 1 of this CTA
        la.ui_prev = (int)cid - (int)ncl;
        for (int k = 0; k < 2; ++k) {
            la.slot = &s_desc[k];
            la.stage1(); la.stage2(); la.stage3();
            la.ui_prev = la.idx;
        }
        la.stage4();
    }
    __syncthreads();
    for (int k = 0;; ++k) {
        const FDesc& d  = s_desc[k & 1];
        const FDesc& dn = s_desc[(k + 1) & 1];
        if (d.ui >= n_list) break;
        const int     uid = d.uid;
        const UnitDev u   = d.u;
        la.slot    = &s_desc[k & 1];
        la.ui_prev = dn.ui;
        FPrefetch pf = {nullptr, 0u, 0u, 0};
        if (dn.ui < n_list) {
            const size_t rb = (size_t)dn.u.nx * (dn.u.dtype == WC_F64 ? 8 : 4), pb = rb * dn.u.ny;
            if (R == 1) {
                pf.base = static_cast<const char*>(dn.u.in);
                pf.piece_lines = (uint32_t)((pb * dn.u.nz + 127) / 128);
                pf.nplanes = 1;
            } else {
                const int nbn = (dn.u.ny / 2) / R;
                pf.base = static_cast<const char*>(dn.u.in) + (size_t)(2 * rank * nbn) * rb;
                pf.piece_lines = (uint32_t)((2 * nbn * rb + 127) / 128);
                pf.nplanes = dn.u.nz;
                pf.pitch = pb;
            }
        }
#define WC_FC_UNIT(GEOM) fc_unit<R, CAP, NT, MM>(GEOM, u, uid, S, pf, la, rank, xph1, xph2, xph3, states, \
                                             one_minus_keep, global_key, mode, pol, lt)
        if constexpr (STATIC) {
            constexpr int CUBE = R == 1 ? (CAP <= 512 ? 8 : CAP <= 4096 ? 16 : 32) : 64;
            if (u.dtype == WC_F64) WC_FC_UNIT((SGeom<CUBE, CUBE, CUBE, 8, R>()));
            else                   WC_FC_UNIT((SGeom<CUBE, CUBE, CUBE, 4, R>()));
        } else {
            FGeom g;
            fused_geom(u.nx, u.ny, u.nz, u.dtype, R, CAP, g);
            WC_FC_UNIT(g);
        }
#undef WC_FC_UNIT
    }
    WC_MARK(21);
    if (R > 1) cluster_sync_all();   // no CTA may exit while peers can still write into its smem
    WC_MARK(22);
}

// ---- launchers --------------------------------------------------------------------------------------
template <int R, int CAP, int NT, bool STATIC, bool MM>
static cudaError_t launch_fc1(int kid, int mode, const UnitDev* units, UnitState* states, const int* list,
                             int n, double omk, const u64* gkey, int sm_count, cudaStream_t st,
                             LaunchStats* ls, int* work_counter) {
    int& max_clusters = MM ? ls->occ_mm[kid] : ls->occ[kid];   // resident clusters (CTAs for R = 1) on this ctx's device
    auto kern = k_fused_compress<R, CAP, NT, STATIC, MM>;
    constexpr int smem = FSmem<R, CAP>::TOTAL;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim         = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream           = st;
    cudaLaunchAttribute attr[1];
    if (R > 1) {
        attr[0].id               = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = R;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs    = attr;
        cfg.numAttrs = 1;
    }
    if (max_clusters == 0) {
        if (R > 1) {
            cfg.gridDim = dim3(R * sm_count);
            int nc = 0;
            e = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
            if (e != cudaSuccess) return e;
            max_clusters = nc;
        } else {
            int per_sm = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
            if (e != cudaSuccess) return e;
            max_clusters = per_sm * sm_count;
        }
        if (max_clusters < 1) return cudaErrorLaunchOutOfResources;
    }
    const int nc = max_clusters < n ? max_clusters : n;
    cfg.gridDim = dim3(nc * R);
    ls->begin(kid, st);
    e = cudaLaunchKernelEx(&cfg, kern, units, states, list, n, omk, gkey, mode, work_counter);
    ls->end(st);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

// MM (ingest statistics) is a separate instantiation: the min / max tracking costs two live registers in
// phase A, which the 64-register literal kernels do not have to spare.
template <int R, int CAP, int NT, bool STATIC>
static cudaError_t launch_fc(int kid, int mode, const UnitDev* units, UnitState* states, const int* list,
                             int n, double omk, const u64* gkey, int sm_count, cudaStream_t st,
                             LaunchStats* ls, int* work_counter) {
    if (mode & FUSED_MINMAX)
        return launch_fc1<R, CAP, NT, STATIC, true>(kid, mode & ~FUSED_MINMAX, units, states, list, n, omk, gkey,
                                                    sm_count, st, ls, work_counter);
    return launch_fc1<R, CAP, NT, STATIC, false>(kid, mode, units, states, list, n, omk, gkey, sm_count, st, ls,
                                                 work_counter);
}

cudaError_t launch_fused_compress(int fused_cls, int mode, const UnitDev* units, UnitState* states,
                                  const int* unit_list, int n_list, double one_minus_keep,
                                  const u64* global_key, int sm_count, cudaStream_t st,
                                  LaunchStats* ls, int* work_counter) {
    if (n_list <= 0) return cudaSuccess;
    if (xs_class_slabs(fused_cls))
        return launch_xs_compress(fused_cls, mode, units, states, unit_list, n_list, one_minus_keep, global_key, sm_count,
                                  st, ls);
    switch (fused_cls) {
    case FUSED_CLS_R1:
        return launch_fc<1, 32768, 512, false>(KID_FUSED_C1, mode, units, states, unit_list, n_list,
                                               one_minus_keep, global_key, sm_count, st, ls, work_counter);
    case FUSED_CLS_R8:
        return launch_fc<8, 32768, 512, false>(KID_FUSED_C8, mode, units, states, unit_list, n_list,
                                               one_minus_keep, global_key, sm_count, st, ls, nullptr);
    case FUSED_CLS_R4:
        return launch_fc<4, 32768, 512, false>(KID_FUSED_C4, mode, units, states, unit_list, n_list,
                                               one_minus_keep, global_key, sm_count, st, ls, nullptr);
    case FUSED_CLS_R2:
        return launch_fc<2, 32768, 512, false>(KID_FUSED_C2, mode, units, states, unit_list, n_list,
                                               one_minus_keep, global_key, sm_count, st, ls, nullptr);
    case FUSED_CLS_CUBE32:
        return launch_fc<1, 32768, 1024, true>(KID_FUSED_C1S, mode, units, states, unit_list, n_list,
                                               one_minus_keep, global_key, sm_count, st, ls, work_counter);
    case FUSED_CLS_R1S:
        return launch_fc<1, 4096, 128, false>(KID_FUSED_C1T, mode, units, states, unit_list, n_list,
                                              one_minus_keep, global_key, sm_count, st, ls, work_counter);
    case FUSED_CLS_CUBE16:
        return launch_fc<1, 4096, 256, true>(KID_FUSED_C16, mode, units, states, unit_list, n_list,
                                             one_minus_keep, global_key, sm_count, st, ls, work_counter);
    case FUSED_CLS_CUBE8:    // one warp per CTA, up to 32 CTAs per SM
        return launch_fc<1, 512, 32, true>(KID_FUSED_C8C, mode, units, states, unit_list, n_list,
                                           one_minus_keep, global_key, sm_count, st, ls, work_counter);
    case FUSED_CLS_CUBE64:
        return launch_fc<8, 32768, 1024, true>(KID_FUSED_C8S, mode, units, states, unit_list, n_list,
                                               one_minus_keep, global_key, sm_count, st, ls, nullptr);
    }
    return cudaErrorInvalidValue;
}

#ifdef WC_PHASE_PROFILE
// debug: sums over CTAs of the phase cycle counters; reset = zero them afterwards
cudaError_t debug_phase_cycles(unsigned long long out[8], bool reset) {
    static unsigned long long h[1024][8];
    cudaError_t e = cudaMemcpyFromSymbol(h, g_phase_cycles, sizeof(h));
    if (e != cudaSuccess) return e;
    for (int p = 0; p < 8; ++p) out[p] = 0;
    for (int c = 0; c < 1024; ++c)
        for (int p = 0; p < 8; ++p) out[p] += h[c][p];
    if (reset) {
        static unsigned long long z[1024][8];
        e = cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
    }
    return e;
}
#endif

// ---- big boxes: forward transform of one y-slab per CTA, coefficients to the unit's scratch -----------------------
// Boxes no cluster holds (more than 262144 cells: 128^3 ..., or a half-height no cluster size divides) take two passes:
// this kernel (transform + arg-max key), then threshold and ordered packing from the coefficient scratch.  An item is a
// (unit, y-slab) of at most 32768 cells, exactly the slab of the cluster kernels: phase A of k_fused_compress fills the
// shared-memory array C from coalesced 16-byte row loads, and C leaves in whole segments (nb * Z consecutive
// coefficients of the flat order) with 16-byte stores — 8N in, 4N out, where the tiled transform of the generic path
// (shared-memory transposes, 4-byte scattered stores) ran at 2.3 TB/s.
template <int NT>
__global__ void __launch_bounds__(NT, 1)
k_big_forward(const UnitDev* __restrict__ units, UnitState* __restrict__ states, const int* __restrict__ unit_list,
              int n_list, int s_rt, int* __restrict__ work_counter) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* const C     = reinterpret_cast<float*>(smem);
    u64* const   s_red = reinterpret_cast<u64*>(smem + (32768 + F_CPAD_BIG) * 4);          // [32]
    int* const   s_it  = reinterpret_cast<int*>(smem + (32768 + F_CPAD_BIG) * 4 + 256);    // [2]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u64 pol = l2_policy_evict_first();
    const int n_items = n_list * s_rt;
    if (tid == 0) s_it[0] = atomicAdd(work_counter, 1);
    __syncthreads();
    for (int k = 0;; ++k) {
        const int item = s_it[k & 1];
        if (item >= n_items) break;
        int next = 0;
        if (tid == 0) next = atomicAdd(work_counter, 1);       // consumed at the end of this item
        const int uid = __ldg(unit_list + item / s_rt);
        const uint32_t rank = (uint32_t)(item % s_rt);
        const UnitDev u = units[uid];
        // Items are handed out in order, so the item gridDim.x further on starts when this one ends, on this CTA or on a
        // neighbour: its input slab goes into L2 while this item's coefficients are written out (L2 is shared, so it does
        // not matter who gets it).  The descriptor loads are issued here and consumed after phase A.
        const int ahead = item + (int)gridDim.x;
        const void* pf_in = nullptr;
        int pf_nx = 0, pf_ny = 0, pf_nz = 0, pf_dt = 0;
        if (ahead < n_items) {
            const UnitDev* u2 = units + __ldg(unit_list + ahead / s_rt);
            pf_in = u2->in; pf_nx = u2->nx; pf_ny = u2->ny; pf_nz = u2->nz; pf_dt = u2->dtype;
        }
        FGeom g;
        fused_geom(u.nx, u.ny, u.nz, u.dtype, s_rt, 32768, g, F_MAXSEG_BIG);
        const int b0 = rank * g.nb;
        float bp = 0.f, bn = 0.f, vmn = 0.f, vmx = 0.f;
        bool nan0 = false;
        {
            const char* in0 = static_cast<const char*>(u.in) + (size_t)(2 * b0) * ((size_t)g.X * g.es);
            if (g.es == 8) phase_a<NT, 8, false>(g, in0, C, rank, pol, bp, bn, nan0, vmn, vmx);
            else           phase_a<NT, 4, false>(g, in0, C, rank, pol, bp, bn, nan0, vmn, vmx);
        }
        if (pf_in) {
            const size_t row = (size_t)pf_nx * (pf_dt == WC_F64 ? 8 : 4);
            const int nb2 = (pf_ny / 2) / s_rt;
            const size_t piece = 2 * (size_t)nb2 * row;                       // one z-plane's share of the slab
            const int lpp = (int)((piece + 127) >> 7);
            const char* base = static_cast<const char*>(pf_in) + (size_t)(2 * (ahead % s_rt) * nb2) * row;
            for (int i = tid; i < lpp * pf_nz; i += NT) {
                const int z = i / lpp, l = i - z * lpp;
                asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(base + (size_t)z * row * pf_ny + (size_t)l * 128));
            }
        }
        __syncthreads();
        // C -> scratch, segment by segment (a warp per segment at a time), and the arg-max key of the slab: largest
        // |c|, lowest flat index among equals, NaNs skipped (make_key) — the rule of std::max_element on the flat array
        u64 key = 0ull;
        const int q4 = g.seglen >> 2;                              // float4 per segment (Z % 4 == 0)
#pragma unroll 1
        for (int sg = warp; sg < g.nseg; sg += NT / 32) {
            const float4* src = reinterpret_cast<const float4*>(C + sg * g.seglen + F_PAD * (sg >> 1));
            const uint32_t f0 = (uint32_t)(((sg >> 1) * g.Y + (sg & 1) * g.hy + b0) * g.Z);
            float4* dst = reinterpret_cast<float4*>(u.coef + f0);
#pragma unroll 2
            for (int i = lane; i < q4; i += 32) {
                const float4 v = src[i];
                __stcs(dst + i, v);
                const uint32_t f = f0 + 4u * (uint32_t)i;
                key = max_u64(key, max_u64(max_u64(make_key(v.x, f), make_key(v.y, f + 1)),
                                           max_u64(make_key(v.z, f + 2), make_key(v.w, f + 3))));
            }
        }
        if (rank == 0 && tid == 0 && isnan(C[0])) atomicOr(&states[uid].flags, UNIT_FLAG_NAN0);
        key = warp_max_u64(key);
        if (lane == 0) s_red[warp] = key;
        if (tid == 0) s_it[(k + 1) & 1] = next;
        __syncthreads();                                           // C and s_red are rewritten by the next item
        if (warp == 0) {
            u64 x = lane < NT / 32 ? s_red[lane] : 0ull;
            x = warp_max_u64(x);
            if (lane == 0 && x != 0ull) atomicMax(&states[uid].key, x);
        }
    }
}

cudaError_t launch_big_forward(const UnitDev* units, UnitState* states, const int* unit_list, int n_list, int s_rt,
                               int* work_counter, int sm_count, cudaStream_t st, LaunchStats* ls) {
    if (n_list <= 0) return cudaSuccess;
    s_rt &= BIG_SLAB_MASK;      // launch key -> slab count
    if (s_rt < 2) return cudaErrorInvalidValue;
    constexpr int NT = 512, smem = (32768 + F_CPAD_BIG) * 4 + 512;
    cudaError_t e = cudaFuncSetAttribute(k_big_forward<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const long long items = (long long)n_list * s_rt;
    const int nc = (int)(items < sm_count ? items : sm_count);
    ls->begin(KID_BIG_FORWARD, st);
    k_big_forward<NT><<<nc, NT, smem, st>>>(units, states, unit_list, n_list, s_rt, work_counter);
    ls->end(st);
    return cudaGetLastError();
}
// ---- ordered packing from a flat coefficient array (second pass of the big-box / generic units) ------------------
// One pass over the unit's coefficient scratch instead of count + scan + emit (which read it twice and, at 0.8 SM
// cycles per coefficient, ran at a quarter of the fused kernels' packing rate).  An item is a chunk of PK_ELEMS
// consecutive coefficients of one unit: a single TMA bulk copy brings it into shared memory, phases C1 / C2 of the
// fused kernels count and emit it segment by segment (512 coefficients, a warp at a time), and the chunk's position in
// the unit's pair list — pairs kept before it, flat index of the last one — comes from its predecessors through a
// status word per item (single-pass "decoupled look-back": items are handed out in order, so a predecessor is always
// running or done).  status = flag << 62 | count << 31 | (last kept flat index + 1); flag 1: this chunk alone,
// 2: inclusive prefix of the unit up to and including this chunk; zeroed before the launch.  Flag and payload share the
// word, so relaxed 8-byte accesses are enough: with release / acquire the fence in front of every status store waited
// for the warp's own pair stores of the previous item to drain (33 k cycles per item, measured).
constexpr int PK_NT = 256, PK_ELEMS = 8192, PK_SEG = 512;
constexpr int PK_SMEM = 2 * PK_ELEMS * 4 + 1024;
__device__ __forceinline__ u64 pk_pack(u64 flag, uint32_t cnt, int last) {
    return (flag << 62) | ((u64)cnt << 31) | (u64)(uint32_t)(last + 1);
}
struct PkDesc {          // one item, fetched by thread 0 while the previous item is emitted
    const float* coef;   // the chunk's first coefficient
    int2*        out;
    int          item, uid, chunk, f0, ne, n;
    float        tf;
    int          c_f0, c_ne;   // the chunk counted ahead by this ticket
    float        c_tf;
    const float* c_coef;
};
// Two chunk buffers: while an item is emitted, thread 0 takes the next ticket, reads its descriptors and starts its
// bulk copy into the other buffer.  The ticket is taken as late as that (not an item ahead): a chunk that is taken but
// not yet counted holds up every later chunk of its unit in the look-back.
__global__ void __launch_bounds__(PK_NT, 3)
k_big_pack(const UnitDev* __restrict__ units, UnitState* __restrict__ states, const int2* __restrict__ items, int n_items,
           u64* __restrict__ status, int* __restrict__ work_counter, int ahead) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char*  tail    = smem + 2 * PK_ELEMS * 4;
    uint32_t* const s_pk    = reinterpret_cast<uint32_t*>(tail);             // [32] cnt << 16 | last (0xffff: none)
    int* const      s_base  = reinterpret_cast<int*>(tail + 128);            // [32] pairs of the chunk before the segment
    int* const      s_prev  = reinterpret_cast<int*>(tail + 256);            // [32] last kept local index before it, -1
    int* const      s_lb    = reinterpret_cast<int*>(tail + 384);            // base count, base last
    int* const      s_next  = reinterpret_cast<int*>(tail + 392);
    int* const      s_ca    = reinterpret_cast<int*>(tail + 640);            // [2][8] count-ahead partials
    PkDesc* const   s_desc  = reinterpret_cast<PkDesc*>(tail + 448);         // [2] x 64 bytes
    const uint32_t  bars    = smem_u32(tail + 576);                          // [2]
    static_assert(sizeof(PkDesc) == 64, "PkDesc layout");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const u64 pol = l2_policy_evict_first();
    u64 pol_last;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    auto fetch = [&](int slot) {                                             // thread 0
        PkDesc& d = s_desc[slot];
        const int item = atomicAdd(work_counter, 1);
        d.item = item;
        if (item >= n_items) return;
        const int2 it = __ldg(items + item);
        const UnitDev u = units[it.x];
        d.uid = it.x; d.chunk = it.y; d.n = u.n;
        d.f0 = it.y * PK_ELEMS;
        d.ne = min(PK_ELEMS, u.n - d.f0);                                    // coefficients of this chunk (> 0)
        d.coef = u.coef + d.f0;
        d.out = reinterpret_cast<int2*>(u.out);
        d.tf = states[it.x].thresh_f;
        // the chunk this CTA counts ahead of time (see "count ahead" below)
        d.c_ne = 0;
        if (item + ahead < n_items) {
            const int2 ic = __ldg(items + item + ahead);
            const UnitDev uc = units[ic.x];
            d.c_f0 = ic.y * PK_ELEMS;
            d.c_ne = min(PK_ELEMS, uc.n - d.c_f0);
            d.c_coef = uc.coef + d.c_f0;
            d.c_tf = states[ic.x].thresh_f;
        }
        // the scratch of a unit is padded to a multiple of 4 floats, and f0 * 4 is a multiple of 16
        const uint32_t bytes = (uint32_t)((d.ne + 3) & ~3) * 4u, bar = bars + 8 * slot;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(smem_u32(smem + slot * PK_ELEMS * 4)), "l"(d.coef), "r"(bytes), "r"(bar), "l"(pol) : "memory");
    };
    if (tid == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fetch(0);
    }
    __syncthreads();
    for (int k = 0;; ++k) {
        const int slot = k & 1;
        const PkDesc d = s_desc[slot];
        if (d.item >= n_items) break;
        WC_PHASE_CLOCK(t0);
        float* const C = reinterpret_cast<float*>(smem + slot * PK_ELEMS * 4);
        const float tf = d.tf;
        const int f0 = d.f0, ne = d.ne, item = d.item, chunk = d.chunk;
        const int nseg = (ne + PK_SEG - 1) / PK_SEG;
        if (tid == 0) *s_next = 0;
        // ---- count ahead: the aggregate (kept count, last kept index) of the chunk `ahead` tickets further on, straight
        // from global memory (its lines stay in L2 for the bulk copy that emits it).  `ahead` exceeds the number of
        // chunks in flight, so by the time a chunk is emitted the aggregates of all its predecessors have long been
        // published and the look-back below never waits — without this every chunk waited for the slowest of its ~30
        // nearest predecessors to load and count (10 k cycles per item, a convoy).  The loads are issued here and
        // consumed after C1, which covers their latency.
        constexpr int CA = PK_ELEMS / 4 / PK_NT;
        float4 ca[CA];
        const int c_n4 = (d.c_ne + 3) >> 2;
        if (d.c_ne > 0) {
            const float4* src = reinterpret_cast<const float4*>(d.c_coef);
#pragma unroll
            for (int r = 0; r < CA; ++r) {
                const int i4 = tid + r * PK_NT;
                ca[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i4 < c_n4)
                    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                                 : "=f"(ca[r].x), "=f"(ca[r].y), "=f"(ca[r].z), "=f"(ca[r].w) : "l"(src + i4), "l"(pol_last));
            }
        }
        mbar_wait_cta(bars + 8 * slot, (uint32_t)(k >> 1) & 1u);
        WC_PHASE_CLOCK(t1);
        // a ragged last segment: NaN is never kept (|NaN| > tf is false for every tf)
        for (int i = tid; i < nseg * PK_SEG - ne; i += PK_NT) C[ne + i] = __int_as_float(0x7fc00000);
        __syncthreads();

        // ---- C1: kept count and last kept index of every segment (fused kernels, phase C1)
        for (int sg = warp; sg < nseg; sg += PK_NT / 32) {
            const float4* cs = reinterpret_cast<const float4*>(C + sg * PK_SEG);
            uint32_t m = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float4 c = cs[lane + 32 * r];
                m |= ((keep_coef(c.x, tf) ? 1u : 0u) | (keep_coef(c.y, tf) ? 2u : 0u) |
                      (keep_coef(c.z, tf) ? 4u : 0u) | (keep_coef(c.w, tf) ? 8u : 0u)) << (4 * r);
            }
            int cnt = __popc(m), last = -1;
            if (m) {
                const int hb = (int)bfind_u32(m);               // bit 4 * r + j -> w = 128 * r + 4 * lane + j
                last = 128 * (hb >> 2) + 4 * lane + (hb & 3);
            }
            cnt  = __reduce_add_sync(0xffffffffu, cnt);
            last = __reduce_max_sync(0xffffffffu, last);
            if (lane == 0) s_pk[sg] = ((uint32_t)cnt << 16) | ((uint32_t)last & 0xffffu);
        }
        __syncthreads();
        WC_PHASE_CLOCK(t2);
        auto count_ahead = [&]() {                               // count ahead, second half (per warp)
            if (d.c_ne <= 0) return;
            const float ctf = d.c_tf;
            int cnt = 0, last = -1;
#pragma unroll
            for (int r = 0; r < CA; ++r) {
                const int e = 4 * (tid + r * PK_NT);
                const float4 c = ca[r];
                const uint32_t m = ((e     < d.c_ne && keep_coef(c.x, ctf)) ? 1u : 0u) | ((e + 1 < d.c_ne && keep_coef(c.y, ctf)) ? 2u : 0u) |
                                   ((e + 2 < d.c_ne && keep_coef(c.z, ctf)) ? 4u : 0u) | ((e + 3 < d.c_ne && keep_coef(c.w, ctf)) ? 8u : 0u);
                cnt += __popc(m);
                if (m) last = e + (int)bfind_u32(m);
            }
            cnt  = __reduce_add_sync(0xffffffffu, cnt);
            last = __reduce_max_sync(0xffffffffu, last);
            if (lane == 0) { s_ca[warp] = cnt; s_ca[8 + warp] = last; }
        };
        // warps 1.. count ahead while warp 0 scans the segments and looks back (it counts its own share afterwards)
        if (warp != 0) count_ahead();

        // ---- scan of the (at most 16) segments + look-back over the unit's earlier chunks, all in warp 0
        if (warp == 0) {
            const uint32_t pk = lane < nseg ? s_pk[lane] : 0xffffu;
            const int cnt = (int)(pk >> 16);
            const int lastl = (pk & 0xffffu) != 0xffffu ? lane * PK_SEG + (int)(pk & 0xffffu) : -1;
            int isum = cnt, imax = lastl;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ps = __shfl_up_sync(0xffffffffu, isum, o), pm = __shfl_up_sync(0xffffffffu, imax, o);
                if (lane >= o) { isum += ps; imax = max(imax, pm); }
            }
            const int total = __shfl_sync(0xffffffffu, isum, 31), lastc = __shfl_sync(0xffffffffu, imax, 31);
            int em = __shfl_up_sync(0xffffffffu, imax, 1);
            if (lane == 0) em = -1;
            s_base[lane] = isum - cnt;
            s_prev[lane] = em;
            const int lastg = lastc >= 0 ? f0 + lastc : -1;       // flat index of the chunk's last kept coefficient
            uint32_t bcnt = 0;
            int blast = -1;
            if (chunk > 0) {
                if (lane == 0 && item < ahead)       // later chunks were counted ahead of time
                    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(status + item), "l"(pk_pack(1, (uint32_t)total, lastg)) : "memory");
                int hi = item - 1;
                for (;;) {
                    const int j = hi - lane;                    // lane 0 = nearest predecessor
                    u64 v = 2ull << 62;                         // in front of the unit's first chunk: nothing kept
                    if (j >= item - chunk)
                        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(status + j) : "memory");
                    // only the words between this chunk and the nearest inclusive prefix have to be there
                    const uint32_t incl = __ballot_sync(0xffffffffu, (v >> 62) == 2);
                    const uint32_t none = __ballot_sync(0xffffffffu, (v >> 62) == 0);
                    const int first = incl ? __ffs(incl) - 1 : 31;
                    if (none & (0xffffffffu >> (31 - first))) continue;      // poll again
                    const bool use = lane <= first;
                    const uint32_t c31 = use ? (uint32_t)((v >> 31) & 0x7fffffffu) : 0u;
                    bcnt += __reduce_add_sync(0xffffffffu, c31);
                    const uint32_t l31 = use ? (uint32_t)(v & 0x7fffffffu) : 0u;
                    const uint32_t has = __ballot_sync(0xffffffffu, l31 != 0u);
                    if (blast < 0 && has) blast = (int)__shfl_sync(0xffffffffu, l31, __ffs(has) - 1) - 1;
                    if (incl) break;
                    hi -= 32;
                }
            }
            if (lane == 0) {
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(status + item),
                             "l"(pk_pack(2, bcnt + (uint32_t)total, lastg >= 0 ? lastg : blast)) : "memory");
                s_lb[0] = (int)bcnt;
                s_lb[1] = blast;
                if (f0 + ne >= d.n) {                           // the unit's last chunk: K and need32 (as k_scan_tiles)
                    const int K = (int)bcnt + total;
                    states[d.uid].npairs = K;
                    const UnitState st = states[d.uid];
                    if (K > 0 && unit_need32(fabsf(key_value(st.key)), st.thresh_f)) states[d.uid].flags = st.flags | UNIT_FLAG_NEED32;
                }
            }
        }
        if (warp == 0) count_ahead();
        __syncthreads();
        WC_PHASE_CLOCK(t3);
        if (tid == 0 && d.c_ne > 0) {
            int tc = 0, tl = -1;
#pragma unroll
            for (int w = 0; w < PK_NT / 32; ++w) { tc += s_ca[w]; tl = max(tl, s_ca[8 + w]); }
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(status + item + ahead),
                         "l"(pk_pack(1, (uint32_t)tc, tl >= 0 ? d.c_f0 + tl : -1)) : "memory");
        }
        if (tid == 0) fetch(slot ^ 1);     // the other buffer was emitted before the barrier at the end of the last item
        __syncwarp();      // thread 0 back in its warp before the loop's __shfl_sync (see k_fused_compress, phase C2)

        // ---- C2: emit (run, value) pairs, segments handed out dynamically (fused kernels, phase C2)
        const uint32_t bcnt = (uint32_t)s_lb[0];
        const int blast = s_lb[1];
        int2* const out = d.out;
        for (;;) {
            int sg = 0;
            if (lane == 0) sg = atomicAdd(s_next, 1);
            sg = __shfl_sync(0xffffffffu, sg, 0);
            if (sg >= nseg) break;
            const int scnt = (int)(s_pk[sg] >> 16);
            if (scnt == 0) continue;
            const float* cs = C + sg * PK_SEG;
            const int fstart = f0 + sg * PK_SEG;
            uint32_t pos = bcnt + (uint32_t)s_base[sg];
            int prev = s_prev[sg] >= 0 ? f0 + s_prev[sg] : blast;
            if (scnt == PK_SEG) {
                for (int w = lane; w < PK_SEG; w += 32)
                    st_pair_pred(true, out + (pos + (uint32_t)w), w == 0 ? fstart - prev - 1 : 0, cs[w], pol);
                continue;
            }
#pragma unroll 1
            for (int w0 = 0; w0 < PK_SEG; w0 += 128) {
                float    c[4];
                uint32_t bal[4];
                bool     kf[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    c[j]   = cs[w0 + 32 * j + lane];
                    kf[j]  = keep_coef(c[j], tf);
                    bal[j] = __ballot_sync(0xffffffffu, kf[j]);
                }
                int prel = prev - (fstart + w0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (bal[j] != 0u) {
                        const uint32_t lower = bal[j] & lt;
                        const int pl = lower ? bfind_alu(lower) : prel;
                        st_pair_pred(kf[j], out + (pos + __popc(lower)), lane - pl - 1, c[j], pol);
                        pos += __popc(bal[j]);
                        prel = bfind_alu(bal[j]);
                    }
                    prel -= 32;
                }
                prev = prel + fstart + w0 + 128;
            }
        }
        __syncthreads();       // this buffer, the segment arrays, s_next and the next descriptor change hands
#ifdef WC_PHASE_PROFILE
        if (tid == 0 && blockIdx.x < 1024) {   // 0 load, 1 C1, 2 scan + look-back, 3 C2
            long long t4 = clock64();
            unsigned long long* pc = g_phase_cycles[blockIdx.x];
            pc[0] += t1 - t0; pc[1] += t2 - t1; pc[2] += t3 - t2; pc[3] += t4 - t3; pc[5] += 1;
        }
#endif
    }
}

cudaError_t launch_big_pack(const UnitDev* units, UnitState* states, const int2* items, int n_items, u64* status,
                            int* work_counter, int sm_count, cudaStream_t st, LaunchStats* ls) {
    if (n_items <= 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_big_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM);
    if (e != cudaSuccess) return e;
    int& per_sm = ls->occ[KID_BIG_PACK];
    if (per_sm == 0) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_big_pack, PK_NT, PK_SMEM);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    }
    e = cudaMemsetAsync(status, 0, sizeof(u64) * (size_t)n_items, st);
    if (e != cudaSuccess) return e;
    const long long slots = (long long)per_sm * sm_count;
    const int nc = (int)(n_items < slots ? n_items : slots);
    ls->begin(KID_BIG_PACK, st);
    k_big_pack<<<nc, PK_NT, PK_SMEM, st>>>(units, states, items, n_items, status, work_counter, 2 * nc);
    ls->end(st);
    return cudaGetLastError();
}
int big_pack_chunks(long long n) { return (int)((n + PK_ELEMS - 1) / PK_ELEMS); }

// slab count of a big box for the compress side (0: the tiled generic transform takes it)
int big_forward_slabs(int nx, int ny, int nz, int dtype, const void* ptr) {
    if (reinterpret_cast<uintptr_t>(ptr) & 15u) return 0;
    FGeom g;
    const int S = big_slabs(nx, ny, nz);
    return (S && fused_geom(nx, ny, nz, dtype, S, 32768, g, F_MAXSEG_BIG)) ? S : 0;
}

// =====================================================================================================
// Fused decompress: rle_decode (src/decompressor.cpp:14-30) + inverse_wavelet_decompose (:79-159)
// =====================================================================================================
// Work item = (unit, y-slab r of S), one CTA per item, no clusters.  The CTA zero-fills its share C of the
// coefficient array in shared memory, decodes the pairs that land in its segments, then inverts two
// c-adjacent blocks per thread straight from C and writes its slab of the box with coalesced 8/16-byte
// stores.  No coefficient scratch in HBM.
//   S = 1 (<= 32768 cells): the item is the whole unit; its pair list is walked once with a block-wide
//          prefix sum of run+1 (flat index of every pair).  Traffic = 8K + 4N (or 8N).
//   S = 8 (<= 262144 cells): a slab's coefficients are 2*X segments of nb*Z consecutive flat indices.  A
//          small index kernel (k_seg_index) first walks each unit's list once and records, for every
//          segment boundary, the first pair at/after it and that pair's flat index; the 8 slab items of a
//          unit are then independent: a warp decodes one segment at a time from its own sub-range of the
//          list.  Traffic = 16K + 4N: the list is read twice, which costs less than the cluster-wide
//          barriers and the DSMEM scatter of the previous design (3x faster on 64^3 boxes).
struct FastDiv {     // q / d, exact: magic multiply when q_max * d < 2^32, plain division otherwise
    uint32_t d, m;
    __device__ __forceinline__ void init(uint32_t div, uint32_t qmax) {
        d = div;
        m = (div > 1 && (unsigned long long)qmax * div < (1ull << 32)) ? (0xffffffffu / div) + 1u : 0u;
    }
    __device__ __forceinline__ uint32_t div(uint32_t q) const { return d == 1 ? q : (m ? __umulhi(q, m) : q / d); }
};

__device__ __forceinline__ uint32_t sat_add(uint32_t a, uint32_t b) {   // saturating at 2^30 (> any ncoef)
    uint32_t s = a + b;
    return s > 0x40000000u ? 0x40000000u : s;
}

__device__ __forceinline__ void ihaar_pair2(float2& avg, float2& diff) {
    const float2 neg1 = make_float2(-1.f, -1.f);
    float2 p = __fadd2_rn(avg, diff);
    float2 m = __ffma2_rn(diff, neg1, avg);     // avg - diff, correctly rounded
    avg  = p;
    diff = m;
}

constexpr int FD_PPT = 8;        // pairs per thread per tile of the block-wide scan

// q-th segment of a warp (slab decode): round-robin, with the parity flipped every other round — odd
// segments are the high-j' (detail, sparse) halves, so a fixed parity would give half the warps all the work.
// NW is even and nseg = 2X is even, so q * NW + (warp ^ 1) stays below nseg whenever q * NW + warp does.
__device__ __forceinline__ int fd_seg_of(int q, int warp, int NW) { return NW > 1 ? q * NW + (warp ^ (q & 1)) : q; }

// Loads FD_PPT consecutive pairs starting at p (a multiple of FD_PPT); pairs past k1 read as (0, 0).
__device__ __forceinline__ void fd_load_tile(const int2* pairs, bool vec16, int p, int k1, int2 (&pr)[FD_PPT]) {
    if (vec16 && p + FD_PPT <= k1) {
#pragma unroll
        for (int j = 0; j < FD_PPT; j += 2) {
            const int4 v = __ldg(reinterpret_cast<const int4*>(pairs + p + j));
            pr[j] = make_int2(v.x, v.y); pr[j + 1] = make_int2(v.z, v.w);
        }
    } else {
#pragma unroll
        for (int j = 0; j < FD_PPT; ++j) pr[j] = (p + j < k1) ? __ldg(pairs + p + j) : make_int2(0, 0);
    }
}

// Block-wide exclusive prefix of run+1 over one tile (saturating; negative runs are flagged, count as 0 and
// are skipped by the callers).  wt = 32 words of shared memory, alternating between two buffers per tile so
// that one barrier per tile suffices.  Returns this thread's exclusive prefix (tile-relative) and the tile total.
template <int NT>
__device__ __forceinline__ uint32_t fd_tile_scan(const int2 (&pr)[FD_PPT], int nvalid, uint32_t* wt, bool& bad,
                                                 uint32_t& ttot) {
    constexpr int NW = NT / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < FD_PPT; ++j) {
        const bool in = j < nvalid;
        if (in && pr[j].x < 0) bad = true;
        s = sat_add(s, (in && pr[j].x >= 0) ? (uint32_t)pr[j].x + 1u : 0u);
    }
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = sat_add(inc, v);
    }
    if (lane == 31) wt[warp] = inc;
    __syncthreads();
    // prefix over the warp totals: one warp-scan instead of NW shared loads per thread
    uint32_t winc = lane < NW ? wt[lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc = sat_add(winc, v);
    }
    ttot = __shfl_sync(0xffffffffu, winc, 31);
    uint32_t wpre = __shfl_sync(0xffffffffu, winc, (warp + 31) & 31);   // inclusive up to warp-1
    if (warp == 0) wpre = 0;
    uint32_t exl = __shfl_up_sync(0xffffffffu, inc, 1);                 // exclusive prefix inside the warp
    if (lane == 0) exl = 0;
    return sat_add(wpre, exl);
}

// Geometry of a unit's table for the index kernels: `slabs` y-slabs (segments of nb * Z coefficients, 2 * X per slab), or,
// for the x-slab classes of wc_xslab.cu (slabs < 0), one entry per plane i' of the flat order (Y * Z coefficients, X planes).
__device__ __forceinline__ void index_geom(const InvUnitDev& iu, int slabs, uint32_t& seglen, int& nseg) {
    if (slabs < 0) {
        seglen = (uint32_t)(iu.ny * iu.nz);
        nseg   = iu.nx;
    } else {
        FGeom g;
        fused_geom(iu.nx, iu.ny, iu.nz, WC_F64, slabs, 32768, g, F_MAXSEG_BIG);
        seglen = (uint32_t)g.seglen;
        nseg   = g.nseg * slabs;                   // == total / seglen
    }
}

// ---- segment index of the slab-decoded units --------------------------------------------------------
// tab[m] = (first pair p whose flat index F_p >= m * seglen, flat index of the pair before it or -1),
// m = 0 .. nseg; pairs at or past `total` (and everything after them) are dropped as rle_decode does:
// tab[m >= first uncovered] = (Kend, .).  The fused compress kernels of the same classes write the same
// table for free (their segment scan holds exactly these two numbers), so a plan round trip skips this kernel.
// The table lives where the generic path keeps its coefficient scratch pointer (DecUnitDev::coef).
template <int NT>
__global__ void __launch_bounds__(NT, 2)
k_seg_index(const DecUnitDev* __restrict__ dec, const InvUnitDev* __restrict__ inv,
            const int* __restrict__ unit_list, int n_list, int* __restrict__ err, int slabs) {
    __shared__ uint32_t s_wt[2][32];
    __shared__ int s_kend, s_flast;
    const int tid = threadIdx.x;
    for (int ui = blockIdx.x; ui < n_list; ui += gridDim.x) {
        const int uid = unit_list[ui];
        const DecUnitDev du = dec[uid];
        const InvUnitDev iu = inv[uid];
        uint32_t seglen;
        int nseg;
        index_geom(iu, slabs, seglen, nseg);
        const uint32_t total = (uint32_t)du.total;
        int2* const tab = reinterpret_cast<int2*>(du.coef);
        const int K = du.npairs_dev ? *du.npairs_dev : du.npairs;
        const int2* pairs = reinterpret_cast<const int2*>(du.pairs);
        const bool vec16 = (reinterpret_cast<uintptr_t>(pairs) & 15u) == 0;
        FastDiv dsl;
        dsl.init(seglen, total);
        if (tid == 0) { s_kend = 0; s_flast = -1; }
        __syncthreads();
        bool bad = false;
        uint32_t carry = 0;
        int kend = 0, flast = -1;
        int tile = 0;
        int2 pr[FD_PPT], nxt[FD_PPT];
        fd_load_tile(pairs, vec16, tid * FD_PPT, K, nxt);
#pragma unroll 1
        for (int p0 = 0; p0 < K; p0 += NT * FD_PPT, ++tile) {
            const int p = p0 + tid * FD_PPT;
#pragma unroll
            for (int j = 0; j < FD_PPT; ++j) pr[j] = nxt[j];
            if (p0 + NT * FD_PPT < K) fd_load_tile(pairs, vec16, p + NT * FD_PPT, K, nxt);   // next tile in flight
            uint32_t ttot;
            uint32_t pre = sat_add(carry, fd_tile_scan<NT>(pr, K - p, s_wt[tile & 1], bad, ttot));
            // last pair of this thread's group that is still inside the box (flat indices only grow)
            int gl = -1, gk = 0;
            {
                uint32_t rp = pre;
#pragma unroll
                for (int j = 0; j < FD_PPT; ++j) {
                    if (p + j < K && pr[j].x >= 0) {
                        const uint32_t f = sat_add(rp, (uint32_t)pr[j].x);
                        if (f < total) { gl = (int)f; gk = p + j + 1; }
                        rp = sat_add(rp, (uint32_t)pr[j].x + 1u);
                    }
                }
            }
            if (gl >= 0) { flast = gl; kend = gk; }
            // the group covers the flat indices (pre - 1, gl]: most groups cross no segment boundary
            if (gl >= 0) {
                uint32_t m = pre == 0 ? 0u : dsl.div(pre - 1u) + 1u;     // next boundary to assign ...
                uint32_t fb = m * seglen;                                 // ... and its flat index
                if ((uint32_t)gl >= fb) {
#pragma unroll
                    for (int j = 0; j < FD_PPT; ++j) {
                        if (p + j < K && pr[j].x >= 0) {
                            const uint32_t f = sat_add(pre, (uint32_t)pr[j].x);
                            if (f < total)
                                for (; fb <= f; fb += seglen, ++m) tab[m] = make_int2(p + j, (int)pre - 1);
                            pre = sat_add(pre, (uint32_t)pr[j].x + 1u);
                        }
                    }
                }
            }
            carry = sat_add(carry, ttot);
        }
        if (bad) atomicOr(err, 1);
        if (kend > 0) { atomicMax(&s_kend, kend); atomicMax(&s_flast, flast); }
        __syncthreads();
        const int ke = s_kend, fl = s_flast;
        for (int m = (fl < 0 ? 0 : (int)dsl.div((uint32_t)fl) + 1) + tid; m <= nseg; m += NT)
            tab[m] = make_int2(ke, fl);
        __syncthreads();
    }
}

// ---- streamed segment index (packed streams that arrive without tables: files, wc_dplan) -------------
// Same table as k_seg_index, built at memory speed: a CTA walks one unit's list tile after tile, but the tiles
// arrive through a three-stage ring of shared-memory buffers that the TMA engine fills (cp.async.bulk, one
// mbarrier per stage) two to three tiles ahead, so the per-tile chain is shared-memory reads + one block scan and
// never waits for HBM (k_seg_index: loads -> scan -> barrier in sequence per tile, 2 TB/s; the chunk-parallel
// look-back version that came between them was bound by five dependent global round trips per 32 KB work item).
// Two CTAs per SM, units handed out dynamically.
//   A thread owns SI_PPT consecutive pairs of a tile and reads only their runs (odd SI_PPT: the lanes' 4-byte reads at a
//   stride of 2 * SI_PPT words are two-way conflicts at worst).  The kernel is bound by its instruction count, not by
//   memory: ~4 instructions per pair for the sums, one block scan per 3840 pairs, and a branch-free search for the
//   (usually single) segment boundary a thread's pairs cross.
//   Bulk copies need 16-byte aligned addresses and sizes: a list that starts on an odd pair is fetched from one
//   pair earlier, an odd count is rounded up — both stay inside the 16-byte granule of a valid pair, so they never
//   leave the allocation's pages; slot r + s0 of a stage holds the tile's r-th pair.
constexpr int SI_NT = 256, SI_PPT = 15, SI_TILE = SI_NT * SI_PPT, SI_STAGES = 3, SI_SLOTS = SI_TILE + 2;
constexpr uint32_t SI_CL = 1u << 19;   // clamp of one pair's run + 1 for units below 2^19 coefficients (the cluster and x-slab
                                       // classes): SI_TILE * SI_CL < 2^31 and the carry is clamped at 2^30: u32 sums never wrap.
constexpr uint32_t SI_CL_BIG = 1u << 26;   // units of 2^19 coefficients and more (FUSED_CLS_RBIG: up to 2^25): a run can exceed
                                       // 2^19 there, so the clamp has to sit above the unit's size; thread sums are then clamped
                                       // at 2^26 and warp totals at 2^28 as well — every clamped value stays above `total`
                                       // ("past the end"), every exact one is untouched, and no u32 sum below can wrap.
constexpr int SI_SMEM = SI_STAGES * SI_SLOTS * 8 + 64 + 2 * 32 * 4 + 32;

__global__ void __launch_bounds__(SI_NT, 2)
k_seg_index3(const DecUnitDev* __restrict__ dec, const InvUnitDev* __restrict__ inv, const int* __restrict__ unit_list,
             int n_list, int* __restrict__ err, int slabs, int* __restrict__ work_counter) {
    extern __shared__ __align__(128) unsigned char smem[];
    int2* const      stage  = reinterpret_cast<int2*>(smem);                                   // [SI_STAGES][SI_SLOTS]
    const uint32_t   bars   = smem_u32(smem + SI_STAGES * SI_SLOTS * 8);                       // [SI_STAGES] mbarriers
    uint32_t* const  s_wt   = reinterpret_cast<uint32_t*>(smem + SI_STAGES * SI_SLOTS * 8 + 64);   // [2][32]
    int* const       s_unit = reinterpret_cast<int*>(smem + SI_STAGES * SI_SLOTS * 8 + 64 + 256);
    u64* const       s_last = reinterpret_cast<u64*>(smem + SI_STAGES * SI_SLOTS * 8 + 64 + 256 + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u64 pol = l2_policy_evict_first();
    if (tid == 0) {
        for (int k = 0; k < SI_STAGES; ++k) mbar_init(bars + 8 * k, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t q0 = 0;                 // tiles this CTA has streamed so far: tile q uses stage q % 3, parity (q / 3) & 1
    bool bad = false;
    for (;;) {
        if (tid == 0) { *s_unit = atomicAdd(work_counter, 1); *s_last = 0ull; }
        __syncthreads();             // also: every stage of the previous unit has been consumed
        const int ui = *s_unit;
        if (ui >= n_list) break;
        const int uid = unit_list[ui];
        const DecUnitDev du = dec[uid];
        const InvUnitDev iu = inv[uid];
        uint32_t seglen;
        int nseg;
        index_geom(iu, slabs, seglen, nseg);
        const uint32_t total = (uint32_t)du.total;
        int2* const tab = reinterpret_cast<int2*>(du.coef);
        const int K = du.npairs_dev ? *du.npairs_dev : du.npairs;
        const int2* pairs = reinterpret_cast<const int2*>(du.pairs);
        const int s0 = (int)((reinterpret_cast<uintptr_t>(pairs) >> 3) & 1u);
        const int ntiles = (K + SI_TILE - 1) / SI_TILE;
        FastDiv dsl;
        dsl.init(seglen, total + seglen);
        const uint32_t cl = total < SI_CL ? SI_CL : SI_CL_BIG;
        auto issue = [&](int t) {                          // thread 0: tile t of this unit into its stage
            const uint32_t q = q0 + (uint32_t)t, st = q % SI_STAGES;
            const int cnt = min(SI_TILE, K - t * SI_TILE) + s0;
            const uint32_t bytes = (uint32_t)((cnt + 1) & ~1) * 8u;
            const uint32_t bar = bars + 8 * st;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(smem_u32(stage + st * SI_SLOTS)), "l"(pairs + (size_t)t * SI_TILE - s0), "r"(bytes), "r"(bar), "l"(pol)
                         : "memory");
        };
        if (tid == 0)
            for (int t = 0; t < ntiles && t < SI_STAGES; ++t) issue(t);
        __syncwarp();      // one-thread region in front of warp collectives: explicit reconvergence (k_fused_compress, phase C2)
        uint32_t carry = 0;
        int gk = 0, gl = -1;                               // this thread's last in-box pair: index + 1, flat index
#pragma unroll 1
        for (int t = 0; t < ntiles; ++t) {
            const uint32_t q = q0 + (uint32_t)t, st = q % SI_STAGES;
            mbar_wait_cta(bars + 8 * st, (q / SI_STAGES) & 1u);
            const int p = t * SI_TILE + tid * SI_PPT;      // first pair of this thread
            int run[SI_PPT];
            {
                const int* sp = reinterpret_cast<const int*>(stage + st * SI_SLOTS + tid * SI_PPT + s0);
#pragma unroll
                for (int j = 0; j < SI_PPT; ++j) run[j] = sp[2 * j];
            }
            if (p + SI_PPT > K) {                          // pairs past the end of the list are dead: run = -1
#pragma unroll
                for (int j = 0; j < SI_PPT; ++j) if (p + j >= K) run[j] = -1;
            }
            uint32_t inc[SI_PPT];
            int any = 0;
            uint32_t s = 0;
#pragma unroll
            for (int j = 0; j < SI_PPT; ++j) {
                any |= run[j];
                inc[j] = min((uint32_t)run[j] + 1u, cl);
            }
            if (any < 0) {
#pragma unroll
                for (int j = 0; j < SI_PPT; ++j)
                    if (run[j] < 0) { inc[j] = 0; if (p + j < K) bad = true; }
            }
#pragma unroll
            for (int j = 0; j < SI_PPT; ++j) s += inc[j];
            s = min(s, SI_CL_BIG);                         // no-op below 2^19 coefficients (s < 2^23)
            uint32_t w = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += v;
            }
            w = min(w, 1u << 28);                          // likewise (w <= 2^28 there)
            uint32_t* const wt = s_wt + (t & 1) * 32;
            if (lane == 31) wt[warp] = w;
            __syncthreads();                               // every thread has read its pairs: the stage is free
            if (tid == 0 && t + SI_STAGES < ntiles) issue(t + SI_STAGES);
            __syncwarp();
            uint32_t ws = lane < SI_NT / 32 ? wt[lane] : 0u;
#pragma unroll
            for (int o = 1; o < SI_NT / 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += v;
            }
            const uint32_t ttot = __shfl_sync(0xffffffffu, ws, SI_NT / 32 - 1);
            uint32_t wpre = __shfl_sync(0xffffffffu, ws, (warp + 31) & 31);
            if (warp == 0) wpre = 0;
            const uint32_t pre = carry + wpre + (w - s);   // flat index this thread's first run starts at
            carry = min(carry + ttot, 1u << 30);
            if (s != 0 && pre < total) {
                const uint32_t end1 = pre + s - 1u;        // flat index of the thread's last live pair
                const bool plain = any >= 0 && end1 < total;   // all pairs live and inside the box
                if (plain) { gk = p + SI_PPT; gl = (int)end1; }
                // segment boundaries m * seglen inside [pre, end1]: each belongs to the pair whose interval
                // [start of its run, its flat index] holds it
                const uint32_t mlo = dsl.div(pre + seglen - 1u), mhi = dsl.div(min(end1, total - 1u));
                if (plain && mlo == mhi) {
                    // one boundary (the usual case): count the pairs that end in front of it, without a branch
                    const uint32_t fb = mlo * seglen;
                    uint32_t rp = pre, start = pre;
                    int nb = 0;
#pragma unroll
                    for (int j = 0; j < SI_PPT; ++j) {
                        rp += inc[j];
                        const bool before = rp <= fb;
                        nb += before ? 1 : 0;
                        start = before ? rp : start;
                    }
                    tab[mlo] = make_int2(p + nb, (int)start - 1);
                } else if (!plain || mlo < mhi) {
                    uint32_t rp = pre, m = mlo, fb = mlo * seglen;
#pragma unroll
                    for (int j = 0; j < SI_PPT; ++j) {
                        if (inc[j]) {
                            const uint32_t f = rp + inc[j] - 1u;
                            if (f < total) {
                                gk = p + j + 1; gl = (int)f;
                                for (; fb <= f; fb += seglen, ++m) tab[m] = make_int2(p + j, (int)rp - 1);
                            }
                            rp += inc[j];
                        }
                    }
                }
            }
        }
        q0 += (uint32_t)ntiles;
        // entries of the segments that start after the unit's last in-box pair
        if (gk > 0) atomicMax(s_last, ((u64)(uint32_t)gk << 32) | (uint32_t)gl);
        __syncthreads();
        const u64 last = *s_last;
        const int ke = (int)(last >> 32), fl = ke > 0 ? (int)(uint32_t)last : -1;
        for (int m = (fl < 0 ? 0 : (int)dsl.div((uint32_t)fl) + 1) + tid; m <= nseg; m += SI_NT)
            tab[m] = make_int2(ke, fl);
        __syncthreads();             // s_last is reset at the top of the next unit
    }
    if (bad) atomicOr(err, 1);
}

// wc_dplan: the units' pairs arrive as ONE dense stream (unit after unit) plus the per-unit counts.  This kernel
// turns the counts into per-unit pointers (exclusive scan) inside the DecUnitDev table and validates them
// (0 <= K <= ncoef, else the corrupt flag) — nothing of this touches the host.
// One CTA per 1024 units, chained through `chain` (zeroed before the launch): [0] = tile ticket, [1 + t] = status of
// tile t: 0 | 1 << 62 | tile total | 2 << 62 | inclusive prefix (totals stay below 2^62).  Tiles are taken in ticket
// order, so a predecessor is always running or done.  (A single CTA took 40 us for 12800 units: the 40-byte stride of
// the table makes every access its own sector, and one SM retires one sector per cycle.)
__global__ void __launch_bounds__(1024)
k_dec_prepare(DecUnitDev* __restrict__ dec, int n_units, const wc_pair* __restrict__ dense,
              const int32_t* __restrict__ npairs, unsigned long long* __restrict__ chain, int* __restrict__ err) {
    __shared__ long long s_w[32];
    __shared__ long long s_base;
    __shared__ int s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = (int)atomicAdd(chain, 1ull);
    __syncthreads();
    const int tile = s_tile, i = tile * 1024 + tid;
    int k = 0;
    if (i < n_units) {
        k = __ldg(npairs + i);
        if (k < 0 || k > dec[i].total) { atomicOr(err, 1); k = 0; }
    }
    long long inc = k;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += x;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    long long w = s_w[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long x = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += x;
    }
    const long long total = __shfl_sync(0xffffffffu, w, 31);
    const long long wpre = warp ? __shfl_sync(0xffffffffu, w, warp - 1) : 0;
    if (tid == 0) {
        unsigned long long* const st = chain + 1;
        long long base = 0;
        if (tile > 0) {
            asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(st + tile), "l"((1ull << 62) | (unsigned long long)total) : "memory");
            for (int j = tile - 1; j >= 0; --j) {
                unsigned long long v;
                do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(st + j) : "memory"); } while ((v >> 62) == 0);
                base += (long long)(v & ((1ull << 62) - 1));
                if ((v >> 62) == 2) break;
            }
        }
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(st + tile), "l"((2ull << 62) | (unsigned long long)(base + total)) : "memory");
        s_base = base;
    }
    __syncthreads();
    if (i < n_units) {
        dec[i].pairs      = dense + (s_base + wpre + inc - k);
        dec[i].npairs     = k;
        dec[i].npairs_dev = nullptr;
    }
}

// Decodes the pairs [c0, min(c0 + 32 * NCH, e1x)) of one segment: pair p sits at flat index
// base + sum of (run + 1) over the pairs c0 .. p; `base` is advanced past the last one.
template <int NCH>
__device__ __forceinline__ void fd_decode_chunks(const int2* __restrict__ pairs, int c0, int e1x, int lane,
                                                 uint32_t& base, uint32_t fseg, uint32_t seglen, uint32_t total,
                                                 float* cseg) {
    int2     pv[NCH];
    uint32_t inc[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int p = c0 + 32 * c + lane;
        pv[c] = make_int2(-1, 0);
        if (p < e1x) pv[c] = __ldg(pairs + p);
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) inc[c] = pv[c].x >= 0 ? (uint32_t)pv[c].x + 1u : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc[c], o);
            if (lane >= o) inc[c] += v;
        }
    }
    uint32_t tot[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) tot[c] = __shfl_sync(0xffffffffu, inc[c], 31);
    uint32_t b = base;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const uint32_t f = b + inc[c];
        if (pv[c].x >= 0 && f - fseg < seglen && f < total) cseg[f - fseg] = __int_as_float(pv[c].y);
        b += tot[c];
    }
    base = b;
}

// Two segments at once, NCH chunks of 32 pairs each (both have at most 32 * NCH pairs): all 2 * NCH loads are in flight
// together and the scans are independent chains, so a warp pays ONE memory round trip for the two segments.  Short
// segments are the rule for slab items (64^3: 96 of a slab's 128 segments are detail bands with a few dozen pairs;
// 128^3: 256 segments of ~77 pairs, 16 per warp) and the decode was a chain of one L2 latency per segment.
struct FdSeg {
    int      c0, e1x;       // pair range
    uint32_t base, fseg;    // flat index of the pair before the first; flat index of the segment's first coefficient
    float*   cseg;
};
template <int NCH>
__device__ __forceinline__ void fd_decode_two(const int2* __restrict__ pairs, const FdSeg& a, const FdSeg& b, int lane,
                                              uint32_t seglen, uint32_t total) {
    int2     pv[2 * NCH];
    uint32_t inc[2 * NCH];
#pragma unroll
    for (int c = 0; c < 2 * NCH; ++c) {
        const FdSeg& sgm = c < NCH ? a : b;
        const int p = sgm.c0 + 32 * (c < NCH ? c : c - NCH) + lane;
        pv[c] = make_int2(-1, 0);
        if (p < sgm.e1x) pv[c] = __ldg(pairs + p);
    }
#pragma unroll
    for (int c = 0; c < 2 * NCH; ++c) inc[c] = pv[c].x >= 0 ? (uint32_t)pv[c].x + 1u : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int c = 0; c < 2 * NCH; ++c) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc[c], o);
            if (lane >= o) inc[c] += v;
        }
    }
    uint32_t tot[2 * NCH];
#pragma unroll
    for (int c = 0; c < 2 * NCH; ++c) tot[c] = __shfl_sync(0xffffffffu, inc[c], 31);
    uint32_t ba = a.base, bb = b.base;
#pragma unroll
    for (int c = 0; c < 2 * NCH; ++c) {
        const FdSeg& sgm = c < NCH ? a : b;
        uint32_t& bs = c < NCH ? ba : bb;
        const uint32_t f = bs + inc[c];
        if (pv[c].x >= 0 && f - sgm.fseg < seglen && f < total) sgm.cseg[f - sgm.fseg] = __int_as_float(pv[c].y);
        bs += tot[c];
    }
}

// ---- staged decode (S = 1, table-less packed streams) ------------------------------------------------
// The block-scan decode above is a chain of dependent latencies per tile (HBM load -> sum -> barrier -> scatter) that
// nothing hides with one 128 KB coefficient array per SM.  Here the pair list of the NEXT item is copied into a
// 96 KB staging area of shared memory by the TMA engine (one cp.async.bulk, completion on an mbarrier) while the
// current item is inverted and stored, so the decode phase reads its pairs from shared memory: its HBM time is
// hidden behind the store phase of the previous item and what is left is ~2 k cycles of scan + scatter.
//   ST slot p + sh holds pair p for p < nst = min(K, ST_PAIRS); sh = 2 for 16-byte aligned lists, 1 otherwise (the
//   bulk copy needs 16-byte aligned source, destination and size: a misaligned first pair and an odd last pair are
//   moved by two ordinary loads).  Pairs past nst (lists longer than the staging area) come from global memory /
//   L2 (prefetched), software-pipelined one tile ahead.
//   A thread owns FS_PPT CONSECUTIVE pairs of a tile (one shuffle scan per tile instead of one per 32 pairs); with an
//   odd FS_PPT the lanes' 8-byte shared-memory loads are conflict-free (stride FS_PPT * 8 bytes).
#ifndef WC_FS_PPT
#define WC_FS_PPT 11
#endif
constexpr int FS_PPT    = WC_FS_PPT;     // odd
constexpr int FS_SLOTS  = 12288;               // 96 KB
constexpr int FS_PAIRS  = FS_SLOTS - 2;
constexpr uint32_t FS_CL = 1u << 17;           // clamp of one pair's run + 1: > any ncoef of an S = 1 unit (32768), and
                                               // 1024 threads * FS_PPT * FS_CL < 2^32, so plain u32 adds never wrap
// Issued by ONE thread: stage the first pairs of a list; always completes exactly one phase of `bar`.
__device__ __forceinline__ void fs_issue(const wc_pair* pairs, int K, int2* ST, uint32_t bar, u64 pol) {
    const int nst = K < FS_PAIRS ? K : FS_PAIRS;
    const int s0  = (int)((reinterpret_cast<uintptr_t>(pairs) >> 3) & 1u);
    const int nb  = nst > s0 ? ((nst - s0) & ~1) : 0;           // pairs moved by the bulk copy
    const uint32_t bytes = (uint32_t)nb * 8u;
    if (bytes) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(smem_u32(ST + 2)), "l"(pairs + s0), "r"(bytes), "r"(bar), "l"(pol) : "memory");
    } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
    }
}
// The (at most two) pairs of the staged range the bulk copy cannot move; any thread(s) but in program order before
// the CTA barrier that precedes the decode.  which = 0: the misaligned first pair, 1: the odd last pair.
__device__ __forceinline__ void fs_patch(const wc_pair* pairs, int K, int2* ST, int which) {
    const int nst = K < FS_PAIRS ? K : FS_PAIRS;
    const int s0  = (int)((reinterpret_cast<uintptr_t>(pairs) >> 3) & 1u);
    const int nb  = nst > s0 ? ((nst - s0) & ~1) : 0;
    const int sh  = 2 - s0;
    const int2* gp = reinterpret_cast<const int2*>(pairs);
    if (which == 0) { if (s0 && nst > 0) ST[sh] = __ldg(gp); }
    else            { if (s0 + nb < nst) ST[s0 + nb + sh] = __ldg(gp + s0 + nb); }
}

template <int NT, class G>
__device__ __forceinline__ void fd_decode_staged(const G& g, const int2* __restrict__ pairs, const int K, const uint32_t total,
                                                 float* const C, const int2* const ST, uint32_t* const s_wt,
                                                 const uint32_t bar, const uint32_t parity, int* __restrict__ err) {
    constexpr int NW = NT / 32, TILE = NT * FS_PPT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nst = K < FS_PAIRS ? K : FS_PAIRS;
    const int sh  = 2 - (int)((reinterpret_cast<uintptr_t>(pairs) >> 3) & 1u);
    FastDiv dyz;
    if (!G::is_static) dyz.init((uint32_t)(g.Y * g.Z), total);
    int2 pr[FS_PPT];
#ifdef WC_PHASE_PROFILE
    const long long tw0 = clock64();
#endif
    mbar_wait_cta(bar, parity);
#ifdef WC_PHASE_PROFILE
    if (tid == 0 && blockIdx.x < 1024) g_phase_cycles[blockIdx.x][6] += clock64() - tw0;
#endif
    bool bad = false;
    if (K == (int)total && K > 0) {
        // Every coefficient kept (a negative max keeps everything, SURVEY.md D3': 10 % of the bench's 32^3 units and 30 % of
        // their pairs): all runs of a well-formed list are 0 and pair p sits at flat index p — a straight copy, no scan.
        // Any other run sends the unit through the general path below (pairs past the end are dropped there).
        int nz = 0;
#pragma unroll 4
        for (int q = tid; q < K; q += NT) {
            const int2 v = q < nst ? ST[q + sh] : __ldg(pairs + q);
            nz |= v.x;
            const uint32_t f = (uint32_t)q, ip = G::is_static ? f / (uint32_t)(g.Y * g.Z) : dyz.div(f);
            C[f + F_PAD * ip] = __int_as_float(v.y);
        }
        if (!__syncthreads_or(nz)) return;
        for (int q = tid; q < K; q += NT) {                      // undo, then decode properly
            const uint32_t f = (uint32_t)q, ip = G::is_static ? f / (uint32_t)(g.Y * g.Z) : dyz.div(f);
            C[f + F_PAD * ip] = 0.f;
        }
        __syncthreads();
    }
    uint32_t carry = 0;
    int tile = 0;
#pragma unroll 1
    for (int p0 = 0; p0 < K; p0 += TILE, ++tile) {
        const int p = p0 + tid * FS_PPT;
        if (p0 + warp * (32 * FS_PPT) >= K) {
            // the whole warp lies past the end of the list (now and in every later tile): it only keeps the barrier
            if (lane == 31) s_wt[(tile & 1) * 32 + warp] = 0u;
            __syncthreads();
            continue;
        }
        if (p + FS_PPT <= nst) {
            const int2* sp = ST + p + sh;
#pragma unroll
            for (int j = 0; j < FS_PPT; ++j) pr[j] = sp[j];
        } else {
            // past the staged part (L2 holds it: prefetched with the bulk copy); pairs past the end of the list read as
            // (-1, 0): run + 1 == 0 marks a dead pair.  (Requesting these one tile ahead into registers measured
            // slower: the 14 extra live registers spill in the 64-register kernel.)
#pragma unroll
            for (int j = 0; j < FS_PPT; ++j) pr[j] = (p + j < K) ? __ldg(pairs + p + j) : make_int2(-1, 0);
        }
        // run + 1 per pair, clamped (see FS_CL); dead pairs count 0.  Negative runs raise the corrupt flag and are
        // skipped; the common case (a thread whose pairs are all inside the list, no negative run) takes no branch
        // per pair
        uint32_t inc[FS_PPT];
        int any = 0;
#pragma unroll
        for (int j = 0; j < FS_PPT; ++j) {
            any |= pr[j].x;
            inc[j] = min((uint32_t)pr[j].x + 1u, FS_CL);
        }
        if (any < 0) {
#pragma unroll
            for (int j = 0; j < FS_PPT; ++j)
                if (pr[j].x < 0) { inc[j] = 0; if (p + j < K) bad = true; }
        }
        uint32_t s = 0;
#pragma unroll
        for (int j = 0; j < FS_PPT; ++j) s += inc[j];
        uint32_t w = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += v;
        }
        uint32_t* const wt = s_wt + (tile & 1) * 32;
        if (lane == 31) wt[warp] = w;
        __syncthreads();
        uint32_t ws = lane < NW ? wt[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= o) ws += v;
        }
        const uint32_t ttot = __shfl_sync(0xffffffffu, ws, 31);
        uint32_t wpre = __shfl_sync(0xffffffffu, ws, (warp + 31) & 31);
        if (warp == 0) wpre = 0;
        uint32_t pre = carry + wpre + (w - s);                 // flat index this thread's first run starts at
        carry = min(carry + ttot, 1u << 30);
        if (pre < total && s != 0) {
#pragma unroll
            for (int j = 0; j < FS_PPT; ++j) {
                pre += inc[j];
                const uint32_t f  = pre - 1u;                      // dead pair: the previous pair's index, masked below
                const uint32_t ip = G::is_static ? f / (uint32_t)(g.Y * g.Z) : dyz.div(f);
                if (inc[j] && f < total) C[f + F_PAD * ip] = __int_as_float(pr[j].y);
            }
        }
    }
    if (bad) atomicOr(err, 1);
}

// Unit descriptors are staged two items ahead through shared memory, like FLookahead of the compress
// kernels; K (which may live on the device after a plan round trip) is resolved one item ahead, in time
// for the L2 prefetch of the next unit's pair list (S = 1).
struct __align__(8) FDDesc {
    DecUnitDev du;      // 40 bytes
    InvUnitDev iu;      // 32 bytes
    int        K;       // resolved pair count
    int        uid, ui; // index into dec[] / inv[]; position in the item list (>= n_items: no more work)
    int        pad;
};
static_assert(sizeof(DecUnitDev) == 40 && sizeof(InvUnitDev) == 32 && sizeof(FDDesc) == 88, "FDDesc layout");
// Thread 0 only; every piece of state lives in shared memory (FDLookState) and every global access lands there
// through cp.async, so the hand-out costs the CTA no registers and no step waits for a value it has just requested:
//   stage3 (after the post-decode barrier of item k)  item k's slot receives the descriptors of item k+2 (cp.async);
//           the work-counter atomic for item k+4 is ISSUED — the one value that must travel in a register, across
//           the straight-line inverse phase only
//   stage4 (end of item k)  descriptors of k+2 have landed: its K is requested (plan round trips keep K on the
//           device); the atomic's result becomes the index of item k+4 and its unit id is requested
//   before the post-decode barrier of item k+1: cp.async.wait_all — K of k+2 and the unit id of k+4 are in place
struct FDLookState {
    int idx_a, uid_a;      // item k+2 (k+3 after stage3)
    int idx_b, uid_b;      // item k+3 (k+4 after stage4)
    int batch_next, batch_left;
};
template <int S>
struct FDLookahead {
    const DecUnitDev* dec;
    const InvUnitDev* inv;
    const int*        unit_list;
    int*              work_counter;
    int               n_items, stride;    // items = S per listed unit: item i -> unit_list[i / S], slab i % S
    int               s_rt;               // S == 0: the slab count is a launch argument (one value per launch)
    __device__ __forceinline__ int unit_of(int idx) const { return S ? idx / S : idx / s_rt; }
    FDLookState*      st;
    FDDesc*           slot;        // item k's slot: receives item k+2
    FDDesc*           next_slot;   // item k+1
    int               batch;
    int               raw;         // in flight between stage3 and stage4
    bool              refill;
    __device__ __forceinline__ static void cp4(void* smem_dst, const void* gsrc) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
    }
    __device__ __forceinline__ static void wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
    __device__ __forceinline__ void stage3() {
        slot->ui  = st->idx_a;
        const int uid = st->uid_a;
        slot->uid = uid;
        if (uid >= 0) {
            const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(&slot->du);
            const char*    s0 = reinterpret_cast<const char*>(dec + uid);
#pragma unroll
            for (int b = 0; b < 40; b += 8)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0 + b), "l"(s0 + b) : "memory");
            const uint32_t d1 = (uint32_t)__cvta_generic_to_shared(&slot->iu);
            const char*    s1 = reinterpret_cast<const char*>(inv + uid);
#pragma unroll
            for (int b = 0; b < 32; b += 8)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d1 + b), "l"(s1 + b) : "memory");
        }
        st->idx_a = st->idx_b; st->uid_a = st->uid_b;
        // dynamic hand-out, `batch` consecutive units per atomic: hundreds of CTAs hitting one address cost
        // ~15-30 cycles per atomic chip-wide, which would bound the small-unit kernels
        refill = false;
        if (work_counter && st->batch_left == 0) { raw = atomicAdd(work_counter, batch); refill = true; }
    }
    __device__ __forceinline__ void stage4() {
        wait_all();                                     // descriptors of item k+2
        if (slot->uid >= 0) {
            slot->K = slot->du.npairs;
            if (slot->du.npairs_dev) cp4(&slot->K, slot->du.npairs_dev);
        }
        int idx;
        if (!work_counter) idx = st->idx_a + stride;
        else {
            if (refill) { st->batch_next = raw; st->batch_left = batch; }
            idx = st->batch_next;
            st->batch_next = idx + 1;
            st->batch_left -= 1;
        }
        st->idx_b = idx;
        st->uid_b = -1;
        if (idx < n_items) cp4(&st->uid_b, unit_list + unit_of(idx));
    }
    // the first four items of the CTA: 0 and 1 into the descriptor slots, 2 and 3 into the state
    __device__ __forceinline__ void prologue(FDDesc* s_desc) {
        st->batch_next = 0; st->batch_left = 0;
        int prev = (int)blockIdx.x - stride;
        for (int k = 0; k < 4; ++k) {
            int idx;
            if (!work_counter) idx = prev + stride;
            else {
                if (st->batch_left == 0) { st->batch_next = atomicAdd(work_counter, batch); st->batch_left = batch; }
                idx = st->batch_next;
                st->batch_next = idx + 1;
                st->batch_left -= 1;
            }
            prev = idx;
            const int uid = idx < n_items ? __ldg(unit_list + unit_of(idx)) : -1;
            if (k < 2) {
                FDDesc& d = s_desc[k];
                d.ui = idx; d.uid = uid; d.K = 0;
                if (uid >= 0) {
                    d.du = dec[uid];
                    d.iu = inv[uid];
                    d.K  = d.du.npairs_dev ? *d.du.npairs_dev : d.du.npairs;
                }
            } else if (k == 2) { st->idx_a = idx; st->uid_a = uid; }
            else               { st->idx_b = idx; st->uid_b = uid; }
        }
    }
};

struct FDStage {          // STG kernels only: staging area, its mbarrier, staged items so far (= phase parity)
    int2*    ST;
    uint32_t bar;
    uint32_t n;
    bool     ignore_tab;  // decode every S = 1 unit from the staged list, also when a segment table came with it
    u64      pol;
};
template <int S, int NT, bool STG, class G>
__device__ __forceinline__ void fd_unit(const G& g, const DecUnitDev& du, const InvUnitDev& iu, const int K,
                                        float* const C, uint32_t* const s_wt, FDLookahead<S>& la,
                                        const uint32_t rank, int* __restrict__ err, const bool have_next, FDStage& stg,
                                        const int s_rt = 0) {
    constexpr int NW = NT / 32;
    const int Sx = S ? S : s_rt;            // slabs per unit
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b0 = rank * g.nb;
    const uint32_t total = (uint32_t)du.total;
    WC_PHASE_CLOCK(t0);

    const int2* pairs = reinterpret_cast<const int2*>(du.pairs);
    const bool vec16 = (reinterpret_cast<uintptr_t>(pairs) & 15u) == 0;
    int2 pr[FD_PPT];
    int2 te = make_int2(0, 0);
    // S = 1 units decode by segments too when a table came with them (plan round trip: the compress kernel
    // wrote it); without one they take the block-wide scan, which needs no second pass over the list
    const bool use_tab = S != 1 || (du.coef != nullptr && !(STG && stg.ignore_tab));
    if (!use_tab) {
        if (!STG) fd_load_tile(pairs, vec16, tid * FD_PPT, K, pr);
    } else {
        // segment table entries of this warp's first 16 segments: lane 2q + e <- tab[m(q) + e]
        const int sg = fd_seg_of(lane >> 1, warp, NW);
        if ((lane >> 1) * NW + warp < g.nseg)
            te = __ldg(reinterpret_cast<const int2*>(du.coef) + (sg >> 1) * (2 * Sx) + (sg & 1) * Sx + (int)rank + (lane & 1));
    }

    // 1. zero-fill C (rle_decode starts from zeros, src/decompressor.cpp:17).  The staged kernel instead zeroes C once at
    //    its start and has step 3 put a zero back behind every coefficient it reads ("clean as you go": one pass over
    //    C and one CTA barrier less per item, measured 0.603 -> 0.575 ms); for the table-driven items the separate pass
    //    measured slightly faster (its 16-byte stores overlap the first pair loads).
    if (!STG) {
        const int nwords = g.nlocal + F_PAD * g.X;
        float4* c4 = reinterpret_cast<float4*>(C);
#pragma unroll 4
        for (int i = tid; i < (nwords + 3) / 4; i += NT) c4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
    }
    WC_PHASE_CLOCK(t1);
    WC_PHASE_CLOCK(t2);

    // 2. decode the pairs that land in this item's segments
    if (STG && !use_tab) {
        fd_decode_staged<NT>(g, pairs, K, total, C, stg.ST, s_wt, stg.bar, stg.n & 1u, err);
        ++stg.n;
    } else if (!use_tab) {
        // block-wide scan over the whole list; flat index -> (i', j', k') -> padded index in C
        FastDiv dyz;
        if (!G::is_static) dyz.init((uint32_t)(g.Y * g.Z), total);
        bool bad = false;
        uint32_t carry = 0;
        int tile = 0;
#pragma unroll 1
        for (int p0 = 0; p0 < K; p0 += NT * FD_PPT, ++tile) {
            const int p = p0 + tid * FD_PPT;
            if (tile > 0) fd_load_tile(pairs, vec16, p, K, pr);
            uint32_t ttot;
            uint32_t pre = sat_add(carry, fd_tile_scan<NT>(pr, K - p, s_wt + (tile & 1) * 32, bad, ttot));
            carry = sat_add(carry, ttot);
            // carry < 2^30: nothing saturated up to and including this tile -> plain adds are exact
            const bool plain = carry < 0x40000000u;
#pragma unroll
            for (int j = 0; j < FD_PPT; ++j) {
                const int  run  = pr[j].x;
                const bool live = (p + j < K) && run >= 0;
                const uint32_t f = plain ? pre + (uint32_t)run : sat_add(pre, (uint32_t)max(run, 0));
                // S == 1: the CTA owns every segment; C is the flat array with F_PAD words after every i' slab
                const uint32_t ip  = G::is_static ? f / (uint32_t)(g.Y * g.Z) : dyz.div(f);
                if (live && f < total) C[f + F_PAD * ip] = __int_as_float(pr[j].y);
                const uint32_t nx = plain ? pre + (uint32_t)run + 1u : sat_add(pre, (uint32_t)max(run, 0) + 1u);
                pre = live ? nx : pre;
            }
        }
        if (bad) atomicOr(err, 1);
    } else {
        // one warp per segment: the pairs [tab[m].x, tab[m+1].x) start at flat index tab[m].y >= m * seglen.
        // Lanes 2q, 2q+1 hold the table entries of the warp's q-th segment (loaded before the zero-fill);
        // up to 8 chunks of 32 pairs are in flight per segment before the first one is decoded.
        const uint32_t seglen = (uint32_t)g.seglen;
        bool have_held = false;      // a short segment waiting for a partner (fd_decode_two)
        FdSeg held = {0, 0, 0u, 0u, nullptr};
#pragma unroll 1
        for (int q = 0; q * NW + warp < g.nseg; ++q) {
            const int sg = fd_seg_of(q, warp, NW);
            const int i = sg >> 1, half = sg & 1;
            const int m = i * (2 * Sx) + half * Sx + (int)rank;
            if (q && (q & 15) == 0) {
                // a warp holds the entries of 16 segments at a time: next round (more than 16 segments per
                // warp only happens with few warps and a long x axis, e.g. 48 x 4 x 8 boxes)
                const int qq = q + (lane >> 1), sq = fd_seg_of(qq, warp, NW);
                te = make_int2(0, 0);
                if (qq * NW + warp < g.nseg)
                    te = __ldg(reinterpret_cast<const int2*>(du.coef) + (sq >> 1) * (2 * Sx) + (sq & 1) * Sx + (int)rank + (lane & 1));
            }
            const int ql = q & 15;
            const int e0x = __shfl_sync(0xffffffffu, te.x, 2 * ql), e0y = __shfl_sync(0xffffffffu, te.y, 2 * ql);
            const int e1x = __shfl_sync(0xffffffffu, te.x, 2 * ql + 1);
            float* const cseg = C + i * g.slab + half * g.seglen;    // C index of flat index m * seglen
            const uint32_t fseg = (uint32_t)m * seglen;
            uint32_t base = (uint32_t)e0y;                           // flat index of the pair before the first (or -1)
            if (e1x <= e0x) continue;                                // nothing kept in this segment
            if (S != 1 && e1x - e0x <= 128) {
                // short segment: decoded together with the warp's next short one (one memory round trip for both)
                const FdSeg cur = {e0x, e1x, base, fseg, cseg};
                if (have_held) { fd_decode_two<4>(pairs, held, cur, lane, seglen, total); have_held = false; }
                else           { held = cur; have_held = true; }
                continue;
            }
#pragma unroll 1
            for (int c0 = e0x; c0 < e1x; c0 += 256) {
                // 4 or 8 chunks of 32 pairs at once: the loads are all in flight together, and the chunks'
                // shuffle scans are independent chains the scheduler interleaves (no branch between them)
                if (e1x - c0 <= 128) fd_decode_chunks<4>(pairs, c0, e1x, lane, base, fseg, seglen, total, cseg);
                else                 fd_decode_chunks<8>(pairs, c0, e1x, lane, base, fseg, seglen, total, cseg);
            }
        }
        if (have_held) fd_decode_chunks<4>(pairs, held.c0, held.e1x, lane, held.base, held.fseg, seglen, total, held.cseg);
    }
    WC_PHASE_CLOCK(t3);
    if (tid == 0) la.wait_all();          // K of the next item, unit id of the item after the staged ones
    __syncthreads();

    // the NEXT unit's pair list (S = 1) lands while this unit is inverted: in the staging area (TMA bulk copy; what
    // does not fit is prefetched into L2), or in L2 only
    if (S == 1 && have_next) {
        const int Kn = la.next_slot->K;
        const char* base = reinterpret_cast<const char*>(la.next_slot->du.pairs);
        int line0 = 0;
        if (STG && (stg.ignore_tab || la.next_slot->du.coef == nullptr)) {
            if (tid == 0)  fs_issue(la.next_slot->du.pairs, Kn, stg.ST, stg.bar, stg.pol);
            if (tid == 32) fs_patch(la.next_slot->du.pairs, Kn, stg.ST, 0);
            if (tid == 64) fs_patch(la.next_slot->du.pairs, Kn, stg.ST, 1);
            line0 = (FS_PAIRS * 8) / 128;
        }
        const int nlines = (Kn * 8 + 127) / 128;
        for (int i = line0 + tid; i < nlines; i += NT)
            asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(base + (size_t)i * 128));
    }
    if (S != 1 && have_next) {
        // slab items: the pair ranges of the NEXT item's segments (one per thread) go into L2 while this item is inverted.
        // With ~600-byte ranges and one segment in flight per warp the decode was bound by memory latency (128^3 boxes:
        // 12 GB/s per SM); the table entries are usually L2 hits (the index kernel has just written them).
        const FDDesc* nd = la.next_slot;
        const int2* ntab = reinterpret_cast<const int2*>(nd->du.coef);
        const int nseg_n = 2 * nd->iu.nx;
        if (ntab && tid < nseg_n) {
            const int rank_n = nd->ui % Sx;
            const int2* e = ntab + (tid >> 1) * (2 * Sx) + (tid & 1) * Sx + rank_n;
            const int p0 = __ldg(&e[0].x), p1 = __ldg(&e[1].x);
            const char* b0p = reinterpret_cast<const char*>(nd->du.pairs);
            const uintptr_t a0 = (reinterpret_cast<uintptr_t>(b0p) + (size_t)p0 * 8) & ~(uintptr_t)127;
            const uintptr_t a1 = reinterpret_cast<uintptr_t>(b0p) + (size_t)p1 * 8;
            for (uintptr_t a = a0; a < a1; a += 128)
                asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(a));
        }
    }
    WC_PHASE_CLOCK(t4);
    if (tid == 0) la.stage3();

    // 3. inverse transform (X, then Y, then Z), two c-adjacent blocks per thread, and store the slab
    {
        const uint32_t m_hx = G::is_static ? 0u : fdiv_magic(g.hx), m_cq = G::is_static ? 0u : fdiv_magic(g.ncq);
        const int o1 = g.hx * g.slab, o2 = g.nb * g.Z, o3 = g.hz;
        const int npc = g.hz >> 1;
        const size_t es = iu.dtype == WC_F64 ? 8 : 4;
        const size_t row_bytes = (size_t)g.X * es, plane_bytes = row_bytes * g.Y;
        char* out0 = static_cast<char*>(iu.out) + (size_t)(2 * b0) * row_bytes;
        auto one = [&](int q) {
            const uint32_t cp2 = q & 1, t1 = q >> 1;
            const uint32_t t2 = G::is_static ? t1 / (uint32_t)g.hx : fdiv(t1, m_hx), a = t1 - t2 * g.hx;
            const uint32_t bl = G::is_static ? t2 / (uint32_t)g.ncq : fdiv(t2, m_cq), cq = t2 - bl * g.ncq;
            const int cpi = 2 * cq + cp2;
            if ((npc & 1) && cpi >= npc) return;
            const float* csrc = C + a * g.slab + bl * g.Z + 2 * cpi;
            float2 v[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                float2* const pc = const_cast<float2*>(reinterpret_cast<const float2*>(csrc + (o & 1) * o1 + ((o >> 1) & 1) * o2 + (o >> 2) * o3));
                v[o] = *pc;
                if (STG) *pc = make_float2(0.f, 0.f);        // clean as you go: the next item finds C zeroed
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) ihaar_pair2(v[2 * k], v[2 * k + 1]);                 // X
#pragma unroll
            for (int zi = 0; zi < 2; ++zi)
#pragma unroll
                for (int xi = 0; xi < 2; ++xi) ihaar_pair2(v[zi * 4 + xi], v[zi * 4 + 2 + xi]);   // Y
#pragma unroll
            for (int k = 0; k < 4; ++k) ihaar_pair2(v[k], v[4 + k]);                         // Z
            char* p0 = out0 + (size_t)(4 * cpi) * plane_bytes + (size_t)(2 * bl) * row_bytes + (size_t)a * 2 * es;
#pragma unroll
            for (int zi = 0; zi < 2; ++zi)
#pragma unroll
                for (int yi = 0; yi < 2; ++yi) {
                    const float2 lo = v[zi * 4 + yi * 2], hi = v[zi * 4 + yi * 2 + 1];   // xi = 0, 1
                    char* pa = p0 + zi * plane_bytes + yi * row_bytes;        // block c:   planes 4cpi + zi
                    char* pb = pa + 2 * plane_bytes;                          // block c+1: planes 4cpi + 2 + zi
                    if (iu.dtype == WC_F64) {
                        __stcs(reinterpret_cast<double2*>(pa), make_double2((double)lo.x, (double)hi.x));
                        __stcs(reinterpret_cast<double2*>(pb), make_double2((double)lo.y, (double)hi.y));
                    } else {
                        __stcs(reinterpret_cast<float2*>(pa), make_float2(lo.x, hi.x));
                        __stcs(reinterpret_cast<float2*>(pb), make_float2(lo.y, hi.y));
                    }
                }
        };
        if constexpr (G::is_static) {
            static_assert(G::npairs % NT == 0, "literal geometries fill every slot");
#pragma unroll
            for (int it = 0; it < G::npairs / NT; ++it) one(tid + it * NT);
        } else {
#pragma unroll 1
            for (int q = tid; q < g.npairs; q += NT) one(q);
        }
    }
    if (tid == 0) la.stage4();
    // C is zero-filled again by the next item, and the staged descriptor becomes visible to the CTA
    __syncthreads();
#ifdef WC_PHASE_PROFILE
    if (tid == 0 && blockIdx.x < 1024) {   // 0 zero-fill, 1 -, 2 decode, 3 barrier + prefetch, 4 inverse + store
        long long t5 = clock64();
        unsigned long long* pc = g_phase_cycles[blockIdx.x];
        pc[0] += t1 - t0; pc[1] += t2 - t1; pc[2] += t3 - t2; pc[3] += t4 - t3; pc[4] += t5 - t4; pc[5] += 1;
    }
#endif
}

// padding words of C: F_PAD per i' slab, X <= 64 for the cluster-sized classes, <= 128 for the run-time / 64-slab ones
__host__ __device__ constexpr int fd_cpad(int S) { return (S >= 1 && S <= 8) ? F_CPAD : F_CPAD_BIG; }
// STATIC: every unit of the list is the cube this variant is specialised for (32^3 for S = 1, 64^3 for S = 8, 128^3 for
// S = 64).
template <int S, int CAP, int NT, bool STATIC, bool STG>
__global__ void __launch_bounds__(NT, (CAP <= 512 ? 32 : CAP <= 4096 ? 4 : 1))
k_fused_decompress(const DecUnitDev* __restrict__ dec, const InvUnitDev* __restrict__ inv,
                   const int* __restrict__ unit_list, int n_list, int* __restrict__ err,
                   int* __restrict__ work_counter, int ignore_tab, int s_rt) {
    static_assert(!STG || S == 1, "staged decode: whole-unit items only");
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int BASE = (CAP + fd_cpad(S)) * 4;
    const int Sx = S ? S : s_rt;
    float* const    C    = reinterpret_cast<float*>(smem);
    uint32_t* const s_wt = reinterpret_cast<uint32_t*>(smem + BASE);               // [2][32]
    FDDesc* const s_desc = reinterpret_cast<FDDesc*>(smem + BASE + 256);          // [2] x 88 bytes
    FDStage stg;
    stg.ST  = reinterpret_cast<int2*>(smem + BASE + 1024);                        // STG: [FS_SLOTS] pairs
    stg.bar = smem_u32(smem + BASE + 512);
    stg.n   = 0;
    stg.ignore_tab = (ignore_tab & 1) != 0;
#ifdef WC_PHASE_PROFILE
    if (const int sg = ignore_tab >> 8) {        // experiment: staggered CTA start
        const long long t0 = clock64(), w = (long long)(blockIdx.x & 3) * sg * 256;
        while (clock64() - t0 < w) { }
    }
#endif
    stg.pol = 0;
    const int tid = threadIdx.x;
    if (STG) {
        stg.pol = l2_policy_evict_first();
        if (tid == 0) {
            mbar_init(stg.bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    const int n_items = n_list * Sx;
    if (STG) {   // the coefficient array starts out zeroed; every item leaves it zeroed (fd_unit step 3)
        float4* c4 = reinterpret_cast<float4*>(C);
        for (int i = tid; i < (CAP + F_CPAD) / 4; i += NT) c4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // dynamic hand-out of the items through a global counter (or a static stride without one)
    FDLookahead<S> la;
    la.dec = dec; la.inv = inv; la.unit_list = unit_list;
    la.work_counter = work_counter;
    la.n_items = n_items; la.stride = (int)gridDim.x; la.s_rt = s_rt;
    la.st = reinterpret_cast<FDLookState*>(smem + BASE + 448);
    la.batch = CAP <= 4096 ? 16 : 1; la.raw = 0; la.refill = false;
    if (tid == 0) la.prologue(s_desc);
    __syncthreads();
    if (STG) {      // the first item's list goes into the staging area here, every later one during its predecessor
        const FDDesc& d = s_desc[0];
        if (d.ui < n_items && (stg.ignore_tab || d.du.coef == nullptr)) {
            if (tid == 0)  fs_issue(d.du.pairs, d.K, stg.ST, stg.bar, stg.pol);
            if (tid == 32) fs_patch(d.du.pairs, d.K, stg.ST, 0);
            if (tid == 64) fs_patch(d.du.pairs, d.K, stg.ST, 1);
        }
        __syncthreads();
    }
    for (int k = 0;; ++k) {
        const FDDesc& d = s_desc[k & 1];
        if (d.ui >= n_items) break;
        const DecUnitDev du = d.du;
        const InvUnitDev iu = d.iu;
        const int K = d.K;
        const uint32_t rank = (uint32_t)(d.ui % Sx);
        la.slot      = &s_desc[k & 1];
        la.next_slot = &s_desc[(k + 1) & 1];
        const bool have_next = la.next_slot->ui < n_items;
#define WC_FD_UNIT(GEOM) fd_unit<S, NT, STG>(GEOM, du, iu, K, C, s_wt, la, rank, err, have_next, stg, s_rt)
        if constexpr (STATIC) {
            constexpr int CUBE = S == 1 ? (CAP <= 512 ? 8 : CAP <= 4096 ? 16 : 32) : S == 64 ? 128 : 64;
            WC_FD_UNIT((SGeom<CUBE, CUBE, CUBE, 8, S>()));
        } else {
            FGeom g;
            fused_geom(iu.nx, iu.ny, iu.nz, WC_F64, Sx, CAP, g, S ? F_MAXSEG : F_MAXSEG_BIG);   // same rule as fused_decode_class
            WC_FD_UNIT(g);
        }
#undef WC_FD_UNIT
    }
}

template <int S, int CAP, int NT, bool STATIC, bool STG = false>
static cudaError_t launch_fd(int kid, const DecUnitDev* dec, const InvUnitDev* inv, const int* list, int n,
                             int* err, int sm_count, cudaStream_t st, LaunchStats* ls, int* work_counter,
                             bool build_tables, bool ignore_tab = false, int s_rt = 0) {
    auto kern = k_fused_decompress<S, CAP, NT, STATIC, STG>;
    constexpr int smem = (CAP + fd_cpad(S)) * 4 + 1024 + (STG ? FS_SLOTS * 8 : 0);
    const int Sx = S ? S : s_rt;
    static_assert(smem <= 232448, "exceeds the 227 KB per-CTA shared memory of sm_100");
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    if (S != 1 && build_tables) {
        // segment tables first: one CTA per unit, a few units per SM
        const int nb = n < 2 * sm_count ? n : 2 * sm_count;
        ls->begin(KID_SEG_INDEX, st);
        k_seg_index<512><<<nb, 512, 0, st>>>(dec, inv, list, n, err, Sx);
        ls->end(st);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    int& per_sm = ls->occ[kid];      // resident CTAs per SM (4 for the small-unit variants), cached per ctx
    if (per_sm == 0) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    }
    const long long items = (long long)n * Sx, slots = (long long)per_sm * sm_count;
    const int nc = (int)(slots < items ? slots : items);
    ls->begin(kid, st);
    int karg = ignore_tab ? 1 : 0;
#ifdef WC_PHASE_PROFILE
    if (const char* sgv = getenv("WCGPU_STAGGER")) karg |= atoi(sgv) << 8;
#endif
    kern<<<nc, NT, smem, st>>>(dec, inv, list, n, err, work_counter, karg, s_rt);
    ls->end(st);
    return cudaGetLastError();
}

int fused_decode_slabs(int fused_cls) {
    switch (fused_cls) {
    case FUSED_CLS_R8: case FUSED_CLS_CUBE64: return 8;
    case FUSED_CLS_R4: return 4;
    case FUSED_CLS_R2: return 2;
    }
    if (const int xs = xs_class_slabs(fused_cls)) return xs;
    return 1;
}

cudaError_t launch_seg_index3(int fused_cls, const DecUnitDev* dec, const InvUnitDev* inv, const int* list, int n,
                              int* work_counter, int* err, int sm_count, cudaStream_t st, LaunchStats* ls, int s_rt) {
    if (n <= 0 || !fused_decode_needs_table(fused_cls)) return cudaSuccess;
    const int slabs = fused_cls == FUSED_CLS_RBIG ? (s_rt & BIG_SLAB_MASK)
                    : xs_class_slabs(fused_cls) ? -1 /* plane table */ : fused_decode_slabs(fused_cls);
    if (slabs >= 0 && slabs < 2) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k_seg_index3, cudaFuncAttributeMaxDynamicSharedMemorySize, SI_SMEM);
    if (e != cudaSuccess) return e;
    const int nb = n < 2 * sm_count ? n : 2 * sm_count;
    ls->begin(KID_SEG_INDEX2, st);
    k_seg_index3<<<nb, SI_NT, SI_SMEM, st>>>(dec, inv, list, n, err, slabs, work_counter);
    ls->end(st);
    return cudaGetLastError();
}

cudaError_t launch_dec_prepare(DecUnitDev* dec, int n_units, const wc_pair* dense, const int32_t* npairs,
                               unsigned long long* chain, int* err, cudaStream_t st, LaunchStats* ls) {
    if (n_units <= 0) return cudaSuccess;
    const int tiles = (n_units + 1023) / 1024;
    cudaError_t e = cudaMemsetAsync(chain, 0, sizeof(unsigned long long) * (size_t)(tiles + 1), st);
    if (e != cudaSuccess) return e;
    ls->begin(KID_DEC_PREPARE, st);
    k_dec_prepare<<<tiles, 1024, 0, st>>>(dec, n_units, dense, npairs, chain, err);
    ls->end(st);
    return cudaGetLastError();
}

// Which fused-decompress class a unit belongs to (FUSED_CLS_*, 0 = generic): same geometry rules as compress,
// plus the output pointer alignment for the vector stores.
int fused_decode_class(int nx, int ny, int nz, int out_dtype, const void* out_ptr) {
    if (reinterpret_cast<uintptr_t>(out_ptr) & (out_dtype == WC_F64 ? 15u : 7u)) {
        // element-aligned is enough for the x-slab classes (scalar stores)
        const bool elem_aligned = (reinterpret_cast<uintptr_t>(out_ptr) & (out_dtype == WC_F64 ? 7u : 3u)) == 0;
        return elem_aligned ? xs_class_of(nx, ny, nz) : FUSED_CLS_NONE;
    }
    if (nx == 32 && ny == 32 && nz == 32) return FUSED_CLS_CUBE32;
    if (nx == 64 && ny == 64 && nz == 64) return FUSED_CLS_CUBE64;
    if (nx == 16 && ny == 16 && nz == 16) return FUSED_CLS_CUBE16;
    if (nx == 8 && ny == 8 && nz == 8) return FUSED_CLS_CUBE8;
    FGeom g;
    if (fused_geom(nx, ny, nz, WC_F64, 1, 4096, g)) return FUSED_CLS_R1S;
    if (fused_geom(nx, ny, nz, WC_F64, 1, 32768, g)) return FUSED_CLS_R1;   // WC_F64: keeps the X*es % 16 rule valid for both
    if (fused_geom(nx, ny, nz, WC_F64, 2, 32768, g)) return FUSED_CLS_R2;
    if (fused_geom(nx, ny, nz, WC_F64, 4, 32768, g)) return FUSED_CLS_R4;
    if (fused_geom(nx, ny, nz, WC_F64, 8, 32768, g)) return FUSED_CLS_R8;
    if (big_slabs(nx, ny, nz)) return FUSED_CLS_RBIG;
    return xs_class_of(nx, ny, nz);
}
int fused_decode_slabs_of(int fused_cls, int nx, int ny, int nz) {
    return fused_cls == FUSED_CLS_RBIG ? big_slabs(nx, ny, nz) : fused_decode_slabs(fused_cls);
}
// int2 entries of the segment table a slab-decoded unit needs (0 for the other classes)
size_t fused_decode_table_entries(int fused_cls, int nx, int ny, int nz) {
    if (fused_cls == FUSED_CLS_RBIG) return (size_t)(2 * nx * big_slabs(nx, ny, nz) + 1);
    if (fused_cls == FUSED_CLS_R8 || fused_cls == FUSED_CLS_CUBE64) return (size_t)(2 * nx * 8 + 1);
    if (fused_cls == FUSED_CLS_R4) return (size_t)(2 * nx * 4 + 1);
    if (fused_cls == FUSED_CLS_R2) return (size_t)(2 * nx * 2 + 1);
    if (fused_cls == FUSED_CLS_R1 || fused_cls == FUSED_CLS_CUBE32 || fused_cls == FUSED_CLS_R1S ||
        fused_cls == FUSED_CLS_CUBE16 || fused_cls == FUSED_CLS_CUBE8)
        return (size_t)(2 * nx + 1);
    if (xs_class_slabs(fused_cls) > 1) return (size_t)(nx + 2);      // one entry per plane i' + the end of the list
    return 0;
}
bool fused_decode_needs_table(int fused_cls) {
    return fused_cls == FUSED_CLS_R8 || fused_cls == FUSED_CLS_CUBE64 || fused_cls == FUSED_CLS_R4 || fused_cls == FUSED_CLS_R2 ||
           fused_cls == FUSED_CLS_RBIG || xs_class_slabs(fused_cls) > 1;
}

cudaError_t launch_fused_decompress(int fused_cls, const DecUnitDev* dec, const InvUnitDev* inv,
                                    const int* unit_list, int n_list, int* err, int sm_count,
                                    cudaStream_t st, LaunchStats* ls, int* work_counter, bool build_tables, int stage,
                                    int s_rt) {
    if (n_list <= 0) return cudaSuccess;
    if (const int xs = xs_class_slabs(fused_cls)) {
        if (xs > 1 && build_tables) {      // plane tables by the direct-load index kernel (WC_OPT_SEG_INDEX = 1, mixed lists)
            const int nb = n_list < 2 * sm_count ? n_list : 2 * sm_count;
            ls->begin(KID_SEG_INDEX, st);
            k_seg_index<512><<<nb, 512, 0, st>>>(dec, inv, unit_list, n_list, err, -1);
            ls->end(st);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
        return launch_xs_decompress(fused_cls, dec, inv, unit_list, n_list, err, sm_count, st, ls, work_counter);
    }
    if (fused_cls == FUSED_CLS_RBIG) {
        const int slabs = s_rt & BIG_SLAB_MASK;
        if (slabs < 2) return cudaErrorInvalidValue;
        if ((s_rt & BIG_CUBE128) && slabs == 64)      // the 128^3 cube: literal geometry, 1024 threads
            return launch_fd<64, 32768, 1024, true>(KID_FUSED_D128, dec, inv, unit_list, n_list, err, sm_count, st, ls,
                                                    work_counter, build_tables);
        return launch_fd<0, 32768, 512, false>(KID_FUSED_DBIG, dec, inv, unit_list, n_list, err, sm_count, st, ls,
                                               work_counter, build_tables, false, slabs);
    }
    if (stage && fused_cls == FUSED_CLS_CUBE32)
        return launch_fd<1, 32768, 1024, true, true>(KID_STAGED_D1S, dec, inv, unit_list, n_list, err, sm_count, st, ls,
                                                     work_counter, false, stage == 2);
    switch (fused_cls) {
    case FUSED_CLS_R1:
        return launch_fd<1, 32768, 512, false>(KID_FUSED_D1, dec, inv, unit_list, n_list, err, sm_count, st, ls, work_counter, false);
    case FUSED_CLS_R8:
        return launch_fd<8, 32768, 512, false>(KID_FUSED_D8, dec, inv, unit_list, n_list, err, sm_count, st, ls, work_counter, build_tables);
    case FUSED_CLS_R4:
        return launch_fd<4, 32768, 512, false>(KID_FUSED_D4, dec, inv, unit_list, n_list, err, sm_count, st, ls, work_counter, build_tables);
    case FUSED_CLS_R2:
        return launch_fd<2, 32768, 512, false>(KID_FUSED_D2, dec, inv, unit_list, n_list, err, sm_count, st, ls, work_counter, build_tables);
    case FUSED_CLS_CUBE32:
        return launch_fd<1, 32768, 1024, true>(KID_FUSED_D1S, dec, inv, unit_list, n_list, err, sm_count, st, ls, work_counter, false);
    case FUSED_CLS_CUBE64:
        return launch_fd<8, 32768, 1024, true>(KID_FUSED_D8S, dec, inv, unit_list, n_list, err, sm_count, st, ls, work_counter, build_tables);
    case FUSED_CLS_R1S:
        return launch_fd<1, 4096, 128, false>(KID_FUSED_D1T, dec, inv, unit_list, n_list, err, sm_count, st, ls, work_counter, false);
    case FUSED_CLS_CUBE16:
        return launch_fd<1, 4096, 256, true>(KID_FUSED_D16, dec, inv, unit_list, n_list, err, sm_count, st, ls, work_counter, false);
    case FUSED_CLS_CUBE8:
        return launch_fd<1, 512, 32, true>(KID_FUSED_D8C, dec, inv, unit_list, n_list, err, sm_count, st, ls, work_counter, false);
    }
    return cudaErrorInvalidValue;
}

} // namespace wc
