// positionbaseddynamics_b200/csrc/resident.cuh
//
// k_step_resident: one TimeStepController::step (Simulation/TimeStepController.cpp:75-241) with the particle positions
// RESIDENT IN THE DISTRIBUTED SHARED MEMORY of thread-block clusters for the whole step.
//
// Why.  Measured in round 1 (profiles/README.md): the per-bucket kernels are bound by L2 sector throughput on big scenes
// (every projection gathers and scatters 2-4 half-used 32-byte sectors through L2) and by the colour-phase dependency
// chain on small ones (cfg3: 3,700 dependent phases per step at 2.2-2.9 us each, however they are launched).  Both come
// from the same thing: positions live in L2 and every colour phase is a device-wide round trip.  Here
//   * the particles are partitioned spatially (host, recursive coordinate bisection) into G cluster regions x C tiles; CTA
//     (g, r) keeps its tile's positions in shared memory from the first integration to the last velocity update;
//   * a projection reaches a particle of a sibling tile of the same cluster through DSMEM (mapa + ld/st.shared::cluster),
//     so inside a cluster a colour phase ends with the hardware cluster barrier (barrier.cluster, ~0.2 us) instead of a
//     grid-wide atomic barrier or a kernel boundary;
//   * only particles touched by a constraint that spans two clusters stay in global memory ("global-homed").  The constraints
//     touching them ("X items") are run by dedicated warps of every CTA, which order themselves across clusters through one
//     monotone arrival counter: red.release.gpu after their X items of phase p, ld.acquire.gpu spin before their X items of
//     phase p+1.  The round trip overlaps the shared-memory work the other warps do in the meantime; those never touch the
//     counter.  A scene that fits one cluster (G = 1) has no X items.
//   Two regimes are used (engine.cu:choose_resident_shape, measured in profiles/README.md R2.3): ONE cluster of up to 16 CTAs for
//   scenes whose colour phases are small (cfg1 / cfg3 / cfg4), and G = #SMs clusters of ONE CTA for big scenes (cfg2 / cfg5: 148
//   independent CTAs, 6.9 % of the particles global-homed, 11 % X items; the SINGLE instantiations: plain LDS / STS, __syncthreads).
//   The order of the items inside a tile's run is chosen at flatten time so that the lanes of a quarter-warp hit distinct 16-byte
//   bank groups (engine.cu, "bank-group fill").
//
// Exactness.  Colour phases are kept: inside a colour no two constraints share a particle (checked at flatten time),
// across colours every constraint reads exactly what the reference's sequential sweep would have produced, whichever CTA
// executes it.  The projection code is the one the per-bucket kernels inline (project_streamed_acc), compiled with
// -fmad=false + explicit fmaf, so the result is bit-identical to the graph / launch modes (tested).
//
// Index encoding (flatten time): bit 31 set -> shared memory: bits 30..27 = CTA rank inside the executing cluster, bits
// 26..0 = slot inside that CTA's tile;  bit 31 clear -> device slot in the global float4 array.
// Constraint order inside a (colour,type) bucket: by executing tile, inside a tile the X items (touching a global-homed
// particle) first; tileOff[bucket][2t], [2t+1], [2t+2] delimit the two runs of tile t (relative to the bucket's `first`).
#pragma once
#include <type_traits>
#include <utility>
#include "kernels.cuh"

namespace pbdk {

constexpr unsigned kSmemFlag = 0x80000000u;
constexpr unsigned kRankShift = 27u;
constexpr unsigned kLocalMask = (1u << kRankShift) - 1u;
constexpr size_t kMaxDynamicSmem = 227u * 1024u;  // what a CTA may opt in to on sm_100
// Bank swizzle of a tile: slot s lives at tile[s ^ ((s >> 3) & 7)] (a permutation inside every aligned 64-slot block; tileCap is a
// multiple of 64).  One colour's constraints touch slots 2 or 3 apart, i.e. only half of the eight 16-byte bank groups per
// quarter-warp; the swizzle spreads such strides over all eight (ncu before: 2.2 / 4.6 bank conflicts per shared load / store request).
__host__ __device__ __forceinline__ unsigned tile_swizzle(unsigned s) { return s ^ ((s >> 3) & 7u); }
constexpr int kMaxClusterCtas = 16;   // non-portable cluster size (opt-in attribute), one GPC
// Dynamic shared memory of a CTA: [tile: tileCap float4][run table: nBuckets x RunEntry][xArrive: 1 + nColours][colourStart: nColours + 1]
// this CTA's two runs of one bucket (absolute indices into the type's arrays); rotR: where the run of shared-memory items starts in the
// colour's concatenated item sequence, modulo the number of threads that share it (the buckets of one colour go to different threads)
struct RunEntry { int type; unsigned firstX, nX, firstR, nR, rotR, pad0, pad1; };
__host__ __device__ inline size_t resident_smem_bytes(unsigned tileCap, unsigned nBuckets, unsigned nColours) {
    return (size_t)tileCap * sizeof(float4) + (size_t)nBuckets * sizeof(RunEntry) + (size_t)(2u * nColours + 2u) * sizeof(unsigned);
}

#ifdef PBD_NO_DEEP
constexpr bool kDeepPipeline = false;
#else
constexpr bool kDeepPipeline = true;
#endif
constexpr unsigned type_bit(int t) { return 1u << t; }
constexpr unsigned kMaskClothXPBD = type_bit(PBD_DISTANCE_XPBD) | type_bit(PBD_ISOBENDING_XPBD);
constexpr unsigned kMaskCloth = kMaskClothXPBD | type_bit(PBD_DISTANCE) | type_bit(PBD_ISOBENDING);
constexpr unsigned kMaskLight = type_bit(PBD_DISTANCE) | type_bit(PBD_DISTANCE_XPBD) | type_bit(PBD_DIHEDRAL) | type_bit(PBD_ISOBENDING) |
                                type_bit(PBD_ISOBENDING_XPBD) | type_bit(PBD_VOLUME) | type_bit(PBD_VOLUME_XPBD) | type_bit(PBD_FEMTRIANGLE);
// FEM solids (cfg3): tets with FEM / volume constraints; and cloth + FEM solids + the coupling joints (cfg4).  Lean instantiations
// matter: the everything-kernel is 20k instructions and 226 registers, and every colour phase jumps through it.
constexpr unsigned kMaskFem = type_bit(PBD_FEMTET) | type_bit(PBD_FEMTET_XPBD) | type_bit(PBD_VOLUME) | type_bit(PBD_VOLUME_XPBD);
constexpr unsigned kMaskSolid = kMaskCloth | kMaskFem | type_bit(PBD_FEMTRIANGLE) | type_bit(PBD_BALLJOINT) | type_bit(PBD_RB_PARTICLE_BALLJOINT);
constexpr unsigned kMaskAll = (1u << PBD_NUM_TYPES) - 1u;

// dispatch a statement on the runtime type, restricted to the compiled-in mask (T is a constant inside the statement)
#define PBD_CASE_TYPE(MASK, TT, ...) case TT: if constexpr ((MASK) & type_bit(TT)) { constexpr int T = TT; __VA_ARGS__ } break;
#define PBD_FOR_TYPE(MASK, type, ...)                                                                                              \
    switch (type) {                                                                                                                \
        PBD_CASE_TYPE(MASK, PBD_DISTANCE, __VA_ARGS__) PBD_CASE_TYPE(MASK, PBD_DISTANCE_XPBD, __VA_ARGS__)                         \
        PBD_CASE_TYPE(MASK, PBD_DIHEDRAL, __VA_ARGS__) PBD_CASE_TYPE(MASK, PBD_ISOBENDING, __VA_ARGS__)                            \
        PBD_CASE_TYPE(MASK, PBD_ISOBENDING_XPBD, __VA_ARGS__) PBD_CASE_TYPE(MASK, PBD_FEMTRIANGLE, __VA_ARGS__)                    \
        PBD_CASE_TYPE(MASK, PBD_STRAINTRIANGLE, __VA_ARGS__) PBD_CASE_TYPE(MASK, PBD_VOLUME, __VA_ARGS__)                          \
        PBD_CASE_TYPE(MASK, PBD_VOLUME_XPBD, __VA_ARGS__) PBD_CASE_TYPE(MASK, PBD_FEMTET, __VA_ARGS__)                             \
        PBD_CASE_TYPE(MASK, PBD_FEMTET_XPBD, __VA_ARGS__) PBD_CASE_TYPE(MASK, PBD_STRAINTET, __VA_ARGS__)                          \
        PBD_CASE_TYPE(MASK, PBD_SHAPEMATCHING, __VA_ARGS__) PBD_CASE_TYPE(MASK, PBD_BALLJOINT, __VA_ARGS__)                        \
        PBD_CASE_TYPE(MASK, PBD_RB_PARTICLE_BALLJOINT, __VA_ARGS__)                                                                \
    default: break;                                                                                                                \
    }

struct ResidentArgs {
    float4 *pos, *vel, *oldp, *lastp;
    const Bucket *buckets;        // colour after colour
    const unsigned *colourStart;  // [nColours + 1] bucket ranges of the colours that own buckets
    const unsigned *tileOff;      // [nBuckets][2 nTiles + 1], relative to the bucket's `first`: tile t = [2t] X items.. [2t+1] others.. [2t+2]
    const unsigned *tileStart;    // [nTiles + 1] device slots: tile t owns [tileStart[t], tileStart[t+1])
    const unsigned *tileSmem;     // [nTiles] leading slots of the tile that live in shared memory; the rest is global-homed
    const unsigned *xArrive;      // [1 + nColours] counter arrivals (CTAs that own X work) of the integration phase and of every colour
    unsigned nBuckets, nColours, nTiles, subSteps, maxIter;
    unsigned tileCap;             // float4 slots reserved for the tile in every CTA's shared memory (>= the largest tileSmem)
    unsigned xThreads;            // the last xThreads threads of every CTA run the X items (and nothing else); 0 when there is one cluster
    unsigned clusterSize;         // CTAs per cluster (1: the colour barrier is a plain __syncthreads)
    int l2Prefetch;               // prefetch the next colour's operand runs into L2 (scenes whose constraint stream exceeds L2)
    int relaxedPoll;              // X warps poll the counter with relaxed loads + back-off and fence once (else: acquire loads)
    float h, invH, twoInvH, gx, gy, gz;
    int secondOrder;
    unsigned long long *xCounter; // monotone arrival counter of the X items (global memory)
    unsigned long long xBase;     // its value when this launch starts
    RbState rb;                   // rigid bodies coupled through joints (global memory; single-cluster scenes only)
    unsigned long long *trace;    // development aid (PBD_B200_TRACE): globaltimer per (phase, CTA)
    unsigned tracePhases;
    TypeArrays types[PBD_NUM_TYPES];
};

// ---- cluster primitives --------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}

// Particle accessors of the resident kernel.
//   ClusterAcc: an X item -- each particle is in the shared memory of a CTA of the cluster (DSMEM) or global-homed;
//   SmemAcc   : every other item -- shared memory only (the flag bit of the index is known to be set).
__device__ __forceinline__ unsigned dsmem_addr(unsigned tileBase, unsigned idx) {
    unsigned a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(tileBase + ((idx & kLocalMask) << 4)), "r"((idx >> kRankShift) & 15u));
    return a;
}
__device__ __forceinline__ float4 dsmem_ld(unsigned a) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void dsmem_st(unsigned a, const float4 &v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
struct SmemAcc {
    typedef unsigned Handle;  // shared::cluster address
    unsigned tileBase;        // shared::cta address of this CTA's tile; the same offset in every CTA of the cluster
    __device__ __forceinline__ Handle handle(unsigned idx) const { return dsmem_addr(tileBase, idx); }
    __device__ __forceinline__ float4 ld(Handle h) const { return dsmem_ld(h); }
    __device__ __forceinline__ void st(Handle h, const float4 &v) const { if (v.w != 0.0f) dsmem_st(h, v); }  // static particles never move
};
struct ClusterAcc {
    typedef unsigned long long Handle;  // bit 63: shared::cluster address in the low word, else a global pointer
    float4 *pos;
    unsigned tileBase;
    __device__ __forceinline__ Handle handle(unsigned idx) const {
        return (idx & kSmemFlag) ? ((1ull << 63) | dsmem_addr(tileBase, idx)) : (unsigned long long)(pos + idx);
    }
    __device__ __forceinline__ float4 ld(Handle h) const { return (h >> 63) ? dsmem_ld((unsigned)h) : __ldcg((const float4 *)h); }
    __device__ __forceinline__ void st(Handle h, const float4 &v) const {
        if (v.w == 0.0f) return;
        if (h >> 63) dsmem_st((unsigned)h, v); else __stcg((float4 *)h, v);
    }
};

// One CTA per cluster (SINGLE instantiations: big scenes, one independent CTA per SM; tiny scenes, one CTA): every shared-memory
// particle is in this CTA's own tile -- plain LDS.128 / STS.128, no mapa, no cluster address space.
struct LocalAcc {
    typedef float4 *Handle;
    float4 *tile;
    __device__ __forceinline__ Handle handle(unsigned idx) const { return tile + (idx & kLocalMask); }
    __device__ __forceinline__ float4 ld(Handle h) const { return *h; }
    __device__ __forceinline__ void st(Handle h, const float4 &v) const { if (v.w != 0.0f) *h = v; }
};
struct LocalXAcc {
    typedef unsigned long long Handle;  // bit 63: offset into the tile in the low word, else a global pointer
    float4 *pos, *tile;
    __device__ __forceinline__ Handle handle(unsigned idx) const { return (idx & kSmemFlag) ? ((1ull << 63) | (idx & kLocalMask)) : (unsigned long long)(pos + idx); }
    __device__ __forceinline__ float4 ld(Handle h) const { return (h >> 63) ? tile[(unsigned)h] : __ldcg((const float4 *)h); }
    __device__ __forceinline__ void st(Handle h, const float4 &v) const {
        if (v.w == 0.0f) return;
        if (h >> 63) tile[(unsigned)h] = v; else __stcg((float4 *)h, v);
    }
};

__device__ __forceinline__ Bucket load_bucket(const Bucket *buckets, unsigned bi) {
    Bucket b;
    const int4 raw = __ldg(reinterpret_cast<const int4 *>(buckets) + bi);  // Bucket is 16 bytes
    b.type = raw.x; b.first = (unsigned)raw.y; b.count = (unsigned)raw.z; b.colour = (unsigned)raw.w;
    return b;
}

// L2 prefetch of a contiguous piece of a streamed array (cp.async.bulk.prefetch.L2: one instruction per <= 32 KB, no registers, no
// shared memory): issued one colour ahead, so that the operand loads of the next phase hit L2 instead of paying the DRAM latency on
// the dependency chain of every item (the per-sweep constraint stream of a big scene does not stay in L2 between sweeps).
__device__ __forceinline__ void l2_prefetch_bytes(const void *p, size_t bytes) {
    if (!p || !bytes) return;
    unsigned long long a = (unsigned long long)p, end = (a + bytes + 15ull) & ~15ull;
    a &= ~15ull;
    while (a < end) {
        const unsigned chunk = (unsigned)((end - a) < 32768ull ? (end - a) : 32768ull);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a), "r"(chunk) : "memory");
        a += chunk;
    }
}
// the streamed arrays of one type (mirror of load_streamed): array k of items [first, first + n), one array per calling lane
template <int T>
__device__ __forceinline__ void l2_prefetch_run(const TypeArrays &a, unsigned first, unsigned n, unsigned lane) {
    constexpr bool two = (T == PBD_DISTANCE || T == PBD_DISTANCE_XPBD || T == PBD_BALLJOINT || T == PBD_RB_PARTICLE_BALLJOINT);
    constexpr bool tri = (T == PBD_FEMTRIANGLE || T == PBD_STRAINTRIANGLE);
    const TypeShape sh = type_shape_dev(T);
    if (lane == 0) {
        if (two) l2_prefetch_bytes(a.idx2 + first, (size_t)n * 8u);
        else if (tri) l2_prefetch_bytes(a.idx3[0] + first, (size_t)n * 4u);
        else l2_prefetch_bytes(a.idx4 + first, (size_t)n * 16u);
    } else if (lane <= 2) {
        if (tri) l2_prefetch_bytes(a.idx3[lane] + first, (size_t)n * 4u);
    } else if (lane < 3 + (unsigned)kMaxGeoV) {
        const unsigned k = lane - 3u;
        if ((int)k < sh.nGeoV && a.gv[k]) l2_prefetch_bytes(a.gv[k] + first, (size_t)n * 16u);
    } else if (lane < 3 + (unsigned)kMaxGeoV + (unsigned)kMaxGeoS) {
        const unsigned k = lane - 3u - (unsigned)kMaxGeoV;
        if ((int)k < sh.nGeoS && a.gs[k]) l2_prefetch_bytes(a.gs[k] + first, (size_t)n * 4u);
    } else if (lane == 3 + (unsigned)kMaxGeoV + (unsigned)kMaxGeoS) {
        if (sh.xpbd && a.lambda) l2_prefetch_bytes(a.lambda + first, (size_t)n * 4u);
    }
}
constexpr unsigned kPrefetchLanes = 3u + (unsigned)kMaxGeoV + (unsigned)kMaxGeoS + 1u;

// What a thread fetched ahead of the colour barrier for its first item of the next phase: the streamed operands (immutable) and
// the XPBD multiplier (last written by this very thread one sweep ago), so that after the barrier only the gathers remain.
struct Prefetched { Streamed s; float lam; unsigned bucket; };

template <int T> __device__ __forceinline__ constexpr bool is_xpbd() { return T == PBD_DISTANCE_XPBD || T == PBD_VOLUME_XPBD || T == PBD_ISOBENDING_XPBD || T == PBD_FEMTET_XPBD; }
// types whose item loop keeps the next item's operands in flight (the register-heavy projections would spill with a second operand set)
template <int T> __device__ __forceinline__ constexpr bool is_pipelined() {
    return T == PBD_DISTANCE || T == PBD_DISTANCE_XPBD || T == PBD_ISOBENDING || T == PBD_ISOBENDING_XPBD || T == PBD_VOLUME || T == PBD_VOLUME_XPBD || T == PBD_FEMTRIANGLE || T == PBD_DIHEDRAL;
}
// ... and the ones light enough to keep TWO items ahead in flight (an L2 hit is ~300 cycles, one cloth projection ~60 issue slots)
template <int T> __device__ __forceinline__ constexpr bool is_pipelined2() { return T == PBD_DISTANCE_XPBD || T == PBD_ISOBENDING_XPBD || T == PBD_DISTANCE || T == PBD_ISOBENDING; }

template <unsigned MASK>
__device__ __forceinline__ void prefetch_first(const ResidentArgs &A, const RunEntry &r, unsigned bi, bool iterZero, unsigned tid, bool xRun, Prefetched &pre) {
    pre.bucket = 0xffffffffu;
    const unsigned first = xRun ? r.firstX : r.firstR, n = xRun ? r.nX : r.nR;
    if (tid >= n) return;
    PBD_FOR_TYPE(MASK, r.type,
        pre.s = load_streamed<T>(A.types[T], first + tid);
        if (is_xpbd<T>() && !iterZero) pre.lam = __ldcg(A.types[T].lambda + first + tid);
        pre.bucket = bi;)
}

// items [first, first + n) of a type's arrays: item tid, tid + stride, ... on this thread; with usePre the first one is in `pre`.
// Software pipelined: the streamed operands (and multiplier) of the next item are in flight while the current one is projected.
template <unsigned MASK, bool DEEP, class Acc>
__device__ __forceinline__ void run_items(const ResidentArgs &A, const Acc &acc, int type, unsigned first, unsigned n, bool iterZero, unsigned tid, unsigned stride,
                                          bool usePre, const Prefetched &pre) {
    PBD_FOR_TYPE(MASK, type,
        const TypeArrays &ta = A.types[T];
        unsigned i = tid;
        if (DEEP && is_pipelined2<T>()) {
            if (i < n) {
                auto fetch = [&](unsigned k, Streamed &sx, float &lx) { sx = load_streamed<T>(ta, first + k); if (is_xpbd<T>() && !iterZero) lx = __ldcg(ta.lambda + first + k); };
                Streamed s0, s1, s2; float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f;
                if (usePre) { s0 = pre.s; l0 = pre.lam; } else fetch(i, s0, l0);
                bool h1 = i + stride < n;
                if (h1) fetch(i + stride, s1, l1);
                _Pragma("unroll 1")
                for (;;) {
                    const unsigned j2 = i + 2u * stride;
                    const bool h2 = j2 < n;
                    if (h2) fetch(j2, s2, l2);
                    project_streamed_acc<T, Acc, 0>(acc, ta, first + i, s0, A.h, iterZero, true, l0);
                    if (!h1) break;
                    s0 = s1; l0 = l1; s1 = s2; l1 = l2; h1 = h2; i += stride;
                }
            }
        } else
        if (i < n) {
            Streamed cur; float lam = 0.0f;
            if (usePre) { cur = pre.s; lam = pre.lam; }
            else { cur = load_streamed<T>(ta, first + i); if (is_xpbd<T>() && !iterZero) lam = __ldcg(ta.lambda + first + i); }
            _Pragma("unroll 1")
            for (;;) {
                const unsigned j = i + stride;
                const bool more = j < n;
                Streamed nxt; float lamN = 0.0f;
                if (is_pipelined<T>() && more) { nxt = load_streamed<T>(ta, first + j); if (is_xpbd<T>() && !iterZero) lamN = __ldcg(ta.lambda + first + j); }
                project_streamed_acc<T, Acc, 0>(acc, ta, first + i, cur, A.h, iterZero, true, lam);  // variant 0 only: flatten refuses the full-Q bending layout in this mode
                if (!more) break;
                if (!is_pipelined<T>()) { nxt = load_streamed<T>(ta, first + j); if (is_xpbd<T>() && !iterZero) lamN = __ldcg(ta.lambda + first + j); }
                cur = nxt; lam = lamN; i = j;
            }
        })
}

// semi-implicit Euler of one particle whose position is `x` (TimeStepController.cpp:112-118, TimeIntegration.cpp:7-19)
__device__ __forceinline__ bool integrate_particle(const ResidentArgs &A, unsigned slot, float4 &x) {
    __stcs(A.lastp + slot, __ldcs(A.oldp + slot));
    __stcs(A.oldp + slot, x);
    float4 v = __ldcs(A.vel + slot);
    if (v.w == 0.0f) return false;  // v.w carries the mass
    v.x = fmaf(A.gx, A.h, v.x); v.y = fmaf(A.gy, A.h, v.y); v.z = fmaf(A.gz, A.h, v.z);
    x.x = fmaf(v.x, A.h, x.x); x.y = fmaf(v.y, A.h, x.y); x.z = fmaf(v.z, A.h, x.z);
    __stcs(A.vel + slot, v);
    return true;
}
// TimeIntegration::velocityUpdateFirstOrder / SecondOrder (TimeIntegration.cpp:42-51, 69-79)
__device__ __forceinline__ void velocity_particle(const ResidentArgs &A, unsigned slot, const float4 &x) {
    if (x.w == 0.0f) return;
    const float4 o = __ldcs(A.oldp + slot);
    float4 v = __ldcs(A.vel + slot);
    if (!A.secondOrder) {
        v.x = A.invH * (x.x - o.x); v.y = A.invH * (x.y - o.y); v.z = A.invH * (x.z - o.z);
    } else {
        const float4 l = __ldcs(A.lastp + slot);
        v.x = A.invH * (1.5f * x.x - 2.0f * o.x + 0.5f * l.x);
        v.y = A.invH * (1.5f * x.y - 2.0f * o.y + 0.5f * l.y);
        v.z = A.invH * (1.5f * x.z - 2.0f * o.z + 0.5f * l.z);
    }
    __stcs(A.vel + slot, v);
}

// X warps of a CTA act as one party of the counter protocol (148 arrivals and 148 pollers per colour instead of one per warp: the
// counter is ONE L2 address).  Named barrier 1 synchronises the CTA's XT threads (whole warps); its first thread talks to the counter.
//   wait  : thread 0 spins until every arrival of the earlier phases is visible (ld.acquire.gpu), then releases its CTA's X warps;
//   arrive: after the X warps' stores -- barrier (orders them before thread 0), then thread 0 publishes with red.release.gpu
//           (the fence is cumulative over what the barrier ordered: the cooperative-groups grid.sync pattern).
__device__ __forceinline__ void x_bar(unsigned xThreads) { asm volatile("bar.sync 1, %0;" :: "r"(xThreads) : "memory"); }
__device__ __forceinline__ void x_wait(const unsigned long long *counter, unsigned long long target, unsigned xtid, unsigned xThreads, int relaxedPoll) {
    if (xtid == 0) {
        if (relaxedPoll) {
            while (ld_relaxed_u64(counter) < target) __nanosleep(20);
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        } else {
            while (ld_acquire_u64(counter) < target) { }
        }
    }
    x_bar(xThreads);
}
__device__ __forceinline__ void x_arrive(unsigned long long *counter, unsigned xtid, unsigned xThreads) {
    x_bar(xThreads);
    if (xtid == 0) red_release_add_u64(counter, 1ull);
}

// colour barrier of the cluster: arrive (publishes this thread's shared-memory stores), then wait; one CTA: a block barrier
__device__ __forceinline__ void colour_arrive(bool single) { if (!single) cluster_arrive(); }
__device__ __forceinline__ void colour_wait(bool single) { if (single) __syncthreads(); else cluster_wait(); }

template <unsigned MASK, int THREADS, bool SINGLE>
__global__ void __launch_bounds__(THREADS, 1) k_step_resident(const __grid_constant__ ResidentArgs A) {
    extern __shared__ float4 tile[];
    RunEntry *runs = reinterpret_cast<RunEntry *>(tile + A.tileCap);
    unsigned *sArrive = reinterpret_cast<unsigned *>(runs + A.nBuckets);  // [1 + nColours]
    unsigned *sColour = sArrive + 1 + A.nColours;                         // [nColours + 1]
    const unsigned cta = blockIdx.x;  // tile id = cluster * C + rank (1-D grid, cluster dimension C)
    const unsigned t0 = __ldg(A.tileStart + cta), nSm = __ldg(A.tileSmem + cta), nAll = __ldg(A.tileStart + cta + 1) - t0;
    const unsigned nGl = nAll - nSm;
    // thread roles: the first RT threads run the shared-memory items; the last XT threads (whole warps) run the X items and the
    // global-homed particles and are the only ones that ever wait for another cluster
    const unsigned XT = A.xThreads, RT = THREADS - XT;
    const bool isX = threadIdx.x >= RT;
    const unsigned xtid = threadIdx.x - RT;
    constexpr bool single = SINGLE;  // one CTA per cluster: plain shared-memory accesses, __syncthreads between colours
    typedef typename std::conditional<SINGLE, LocalXAcc, ClusterAcc>::type XAcc;
    typedef typename std::conditional<SINGLE, LocalAcc, SmemAcc>::type RAcc;
    XAcc accX; RAcc accR;
    if constexpr (SINGLE) { accX.pos = A.pos; accX.tile = tile; accR.tile = tile; }
    else { accX.pos = A.pos; accX.tileBase = smem_u32(tile); accR.tileBase = smem_u32(tile); }
    unsigned long long xTarget = A.xBase;  // counter value once every arrival of the phases before the current one is in
    unsigned phase = 0;
    auto stamp = [&](unsigned k) {
        if (A.trace && phase < A.tracePhases && (threadIdx.x == 0 || (k == 1 && threadIdx.x == RT))) A.trace[((size_t)phase * gridDim.x + cta) * 4 + k] = globaltimer_ns();
    };

    // ---- once per launch: the tile and this CTA's view of the phase structure go to shared memory (after a cluster barrier L1 is
    // invalid, so anything read per phase from global memory would cost an L2 round trip on the critical path of every colour)
    for (unsigned i = threadIdx.x; i < nSm; i += THREADS) tile[tile_swizzle(i)] = __ldcg(A.pos + t0 + i);
    for (unsigned bi = threadIdx.x; bi < A.nBuckets; bi += THREADS) {
        const Bucket b = load_bucket(A.buckets, bi);
        const unsigned *off = A.tileOff + (size_t)bi * (2u * A.nTiles + 1u) + 2u * cta;
        const unsigned o0 = __ldg(off), o1 = __ldg(off + 1), o2 = __ldg(off + 2);
        RunEntry r; r.type = b.type; r.firstX = b.first + o0; r.nX = o1 - o0; r.firstR = b.first + o1; r.nR = o2 - o1; r.rotR = r.pad0 = r.pad1 = 0u;
        runs[bi] = r;
    }
    for (unsigned i = threadIdx.x; i < 1u + A.nColours; i += THREADS) { sArrive[i] = __ldg(A.xArrive + i); sColour[i] = __ldg(A.colourStart + i); }
    __syncthreads();
    for (unsigned c = threadIdx.x; c < A.nColours; c += THREADS) {  // rotations: bucket k of a colour starts where bucket k-1 ended
        unsigned accR = 0;
        for (unsigned bi = sColour[c]; bi < sColour[c + 1]; bi++) {
            runs[bi].rotR = (accR % RT) & ~7u; accR += runs[bi].nR;  // multiple of 8: the quarter-warps of a run stay aligned with its item order
        }
    }
    __syncthreads();
    Prefetched pre;
    pre.bucket = 0xffffffffu; pre.lam = 0.0f;
    if (!single) { cluster_arrive(); cluster_wait(); }  // every tile of the cluster is loaded (and every CTA has started) before the first DSMEM access

    for (unsigned sub = 0; sub < A.subSteps; sub++) {
        // ---- prologue: lastX = oldX; oldX = x; semi-implicit Euler
        stamp(0);
        if (isX) {  // global-homed particles: what other clusters wait for
            for (unsigned i = xtid; i < nGl; i += XT) {
                float4 x = __ldcg(A.pos + t0 + nSm + i);
                if (integrate_particle(A, t0 + nSm + i, x)) __stcg(A.pos + t0 + nSm + i, x);
            }
            if (nGl) x_arrive(A.xCounter, xtid, XT);
            stamp(1);
        } else {
            for (unsigned i = threadIdx.x; i < nSm; i += RT) {
                float4 x = tile[tile_swizzle(i)];
                if (integrate_particle(A, t0 + i, x)) tile[tile_swizzle(i)] = x;
            }
            if (A.rb.n && cta == 0 && threadIdx.x < A.rb.n) rb_integrate_body(A.rb, threadIdx.x, A.h, A.gx, A.gy, A.gz);
        }
        xTarget += sArrive[0];
        colour_arrive(single);
        if (A.nColours) prefetch_first<MASK>(A, runs[sColour[0]], sColour[0], true, isX ? xtid : threadIdx.x, isX, pre);
        stamp(2); colour_wait(single); stamp(3);
        phase++;

        // ---- coloured Gauss-Seidel sweeps (TimeStepController.cpp:270-286)
        for (unsigned it = 0; it < A.maxIter; it++) {
            const bool iterZero = (it == 0);
            for (unsigned c = 0; c < A.nColours; c++) {
                const unsigned b0 = sColour[c], b1 = sColour[c + 1];
                stamp(0);
                if (A.l2Prefetch && threadIdx.x < kPrefetchLanes) {  // the next colour's operand runs of this CTA start travelling DRAM -> L2 now
                    const unsigned cn = (c + 1 == A.nColours) ? 0u : c + 1u;
                    for (unsigned bi = sColour[cn]; bi < sColour[cn + 1]; bi++) {
                        const RunEntry r = runs[bi];
                        PBD_FOR_TYPE(MASK, r.type, l2_prefetch_run<T>(A.types[T], r.firstX, r.nX + r.nR, threadIdx.x);)
                    }
                }
                if (isX) {
                    // X items of every bucket of the colour: they touch global-homed particles and are ordered across clusters by the counter
                    bool mine = false;  // CTA-uniform: does this CTA own an X item of this colour?
                    for (unsigned bi = b0; bi < b1; bi++) mine = mine || (runs[bi].nX != 0u);
                    if (mine) {
                        x_wait(A.xCounter, xTarget, xtid, XT, A.relaxedPoll);
#pragma unroll 1
                        for (unsigned bi = b0; bi < b1; bi++) run_items<MASK, false>(A, accX, runs[bi].type, runs[bi].firstX, runs[bi].nX, iterZero, xtid, XT, pre.bucket == bi, pre);
                        x_arrive(A.xCounter, xtid, XT);
                    }
                    stamp(1);
                } else {
                    // everything else: shared memory of this cluster only
#pragma unroll 1
                    for (unsigned bi = b0; bi < b1; bi++) {
                        const unsigned rot = runs[bi].rotR;  // 0 for the first bucket of a colour (the one prefetch_first serves)
                        run_items<MASK, kDeepPipeline && SINGLE && MASK == kMaskClothXPBD>(A, accR, runs[bi].type, runs[bi].firstR, runs[bi].nR, iterZero,
                                                                           threadIdx.x >= rot ? threadIdx.x - rot : threadIdx.x + RT - rot, RT, pre.bucket == bi, pre);
                    }
                }
                xTarget += sArrive[1 + c];
                colour_arrive(single);
                {   // this thread's first item of the next colour streams in while the CTAs synchronise
                    const bool lastColour = (c + 1 == A.nColours);
                    if (!lastColour || it + 1 < A.maxIter) {
                        const unsigned nb = sColour[lastColour ? 0u : c + 1u];
                        prefetch_first<MASK>(A, runs[nb], nb, false, isX ? xtid : threadIdx.x, isX, pre);
                    } else pre.bucket = 0xffffffffu;
                }
                stamp(2); colour_wait(single); stamp(3);
                phase++;
            }
        }

        // ---- epilogue: velocity update (own particles only; the global-homed ones need the last colour's X items of every cluster)
        if (isX) {
            if (nGl) x_wait(A.xCounter, xTarget, xtid, XT, A.relaxedPoll);
            for (unsigned i = xtid; i < nGl; i += XT) velocity_particle(A, t0 + nSm + i, __ldcg(A.pos + t0 + nSm + i));
        } else {
            for (unsigned i = threadIdx.x; i < nSm; i += RT) velocity_particle(A, t0 + i, tile[tile_swizzle(i)]);
            if (A.rb.n && cta == 0 && threadIdx.x < A.rb.n) rb_velocity_body(A.rb, threadIdx.x, A.invH, A.twoInvH, A.secondOrder);
        }
    }
    for (unsigned i = threadIdx.x; i < nSm; i += THREADS) __stcg(A.pos + t0 + i, tile[tile_swizzle(i)]);
}

}  // namespace pbdk
