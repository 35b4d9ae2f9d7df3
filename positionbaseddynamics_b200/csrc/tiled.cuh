// positionbaseddynamics_b200/csrc/tiled.cuh
//
// k_step_tiled: one TimeStepController::step with the particles of each SM's *tile* resident in shared memory.
//
// Why.  Measured (profiles/README.md section 2): the per-bucket kernels are bound by L2 sector throughput, because every
// projection gathers and scatters its 2-4 particles through L2 and uses half of each 32-byte sector.  Here the particles
// are partitioned spatially into one tile per SM (host side, recursive coordinate bisection on the rest positions).  A
// particle is PRIVATE to its tile when every constraint touching it lies entirely inside the tile; private particles are
// loaded into the CTA's shared memory once per substep and all their gathers/scatters stay on chip.  Only SHARED
// particles (touched by a constraint spanning two tiles) keep living in global memory and travel through L2.
//
// Exactness.  The colour phases are kept (grid barrier between colours, as in k_step_persistent): within a colour no two
// constraints share a particle, across colours every constraint sees exactly the values the reference's sweep would
// produce, no matter which CTA executes it.  Same projection code as every other mode (project_streamed_acc), so the
// result is bit-identical to the graph / launch / persistent modes (tested).
//
// Particle indices of a constraint are encoded at flatten time: bit 31 set -> slot in the executing CTA's tile
// (shared memory), else device slot in the global array.  Constraints are stored bucket by bucket and, inside a bucket,
// tile by tile, inside a tile first the ones that touch at least one global particle ("spanning"), then the ones whose
// particles are all in the tile's shared memory ("private"); tileOff[bucket][2t], [2t+1], [2t+2] delimit the two runs of CTA t.
//
// Pipeline inside a CTA (THREADS - 32 worker threads + one manager warp):
//   * the manager streams the NEXT colour's constraint operands (indices + rest data of the CTA's runs, contiguous per
//     array) from HBM into a double-buffered shared-memory stage with bulk async copies (cp.async.bulk + mbarrier
//     complete_tx), so the workers never wait for DRAM;
//   * split grid barrier per colour: the workers run the spanning constraints first and announce them on a named
//     barrier without waiting; the manager collects the announcements, publishes the CTA's arrival (fence + atomic)
//     and spins for the release while the workers already run the private constraints (shared memory only).
#pragma once
#include "persistent.cuh"

namespace pbdk {

constexpr unsigned kSmemFlag = 0x80000000u;
constexpr int kTileCapacity = 8192;                 // private particles per tile kept in shared memory (128 KB)
constexpr unsigned kStageBytes = 48u * 1024u;       // one stage buffer (two of them)
constexpr size_t kTiledSmemBytes = (size_t)kTileCapacity * sizeof(float4) + 2u * kStageBytes + 64u;
// bank swizzle of the tile: slot s lives at sp[s ^ ((s >> 3) & 7)] (a permutation inside every aligned 64-slot block), so that
// the stride-2 / stride-4 slot patterns of one colour's constraints spread over all eight 16-byte bank groups
__host__ __device__ __forceinline__ unsigned tile_swizzle(unsigned s) { return s ^ ((s >> 3) & 7u); }
constexpr int kStreamArrays = 8;                    // idx a/b/c, gv0, gv1, gs0, gs1, lambda (XPBD multipliers)

struct TiledArgs {
    float4 *pos, *vel, *oldp, *lastp;
    const Bucket *buckets;
    const unsigned *tileOff;     // [nBuckets][2 nTiles + 1] offsets relative to the bucket's `first`: tile t = [2t] spanning.. [2t+1] private.. [2t+2]
    const unsigned *tileStart;   // [nTiles + 1] device slots: tile t owns [tileStart[t], tileStart[t+1])
    const unsigned *tilePrivate; // [nTiles] number of leading slots of the tile that live in shared memory
    unsigned nBuckets, subSteps, maxIter;
    float h, invH, gx, gy, gz;
    int secondOrder, trackLast;
    unsigned long long *barrier;
    unsigned long long barrierBase;
    unsigned long long *trace;   // development aid (PBD_B200_TRACE): 4 timestamps per (colour phase, CTA)
    unsigned tracePhases;
    int traceWorker;             // -1: the manager records (collect, arrive, release, colour end); w >= 0: worker warp w records (start, operands there, spanning done, private done)
    int swizzle;                 // tile slots are bank-swizzled (tile_swizzle); must match the index encoding done at flatten time
    int stageLambda;             // XPBD multipliers travel with the staged operands (needs >= 2 colours: the copy for phase p + 1 starts during phase p)
    int stage;                   // 0: workers read the constraint stream straight from global memory (A/B knob)
    TypeArrays types[PBD_NUM_TYPES];
};

struct TileAcc {
    float4 *pos;  // global
    float4 *sp;   // this CTA's tile
    __device__ __forceinline__ float4 ld(unsigned idx) const { return (idx & kSmemFlag) ? sp[idx & ~kSmemFlag] : __ldcg(pos + idx); }
    __device__ __forceinline__ void st(unsigned idx, const float4 &v) const {
        if (v.w != 0.0f) { if (idx & kSmemFlag) sp[idx & ~kSmemFlag] = v; else __stcg(pos + idx, v); }
    }
    __device__ __forceinline__ float4 *global() const { return pos; }
};

// ---- mbarrier / bulk-copy helpers --------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- which arrays a constraint type streams (mirror of load_streamed): element size of stream slot r, 0 = unused ---------------
template <int T> __host__ __device__ constexpr unsigned stream_es(int r) {
    constexpr bool two = (T == PBD_DISTANCE || T == PBD_DISTANCE_XPBD), tri = (T == PBD_FEMTRIANGLE || T == PBD_STRAINTRIANGLE);
    constexpr bool g0 = tri || T == PBD_ISOBENDING || T == PBD_ISOBENDING_XPBD || T == PBD_FEMTET || T == PBD_FEMTET_XPBD || T == PBD_STRAINTET;
    constexpr bool g1 = (T == PBD_FEMTET || T == PBD_FEMTET_XPBD || T == PBD_STRAINTET);
    constexpr bool s0 = two || T == PBD_FEMTRIANGLE || T == PBD_DIHEDRAL || T == PBD_VOLUME || T == PBD_VOLUME_XPBD || g1;
    constexpr bool s1 = (T == PBD_FEMTET || T == PBD_FEMTET_XPBD);
    constexpr bool xpbd = (T == PBD_DISTANCE_XPBD || T == PBD_VOLUME_XPBD || T == PBD_ISOBENDING_XPBD || T == PBD_FEMTET_XPBD);
    return r == 0 ? (two ? 8u : (tri ? 4u : 16u)) : (r == 1 || r == 2) ? (tri ? 4u : 0u) : r == 3 ? (g0 ? 16u : 0u) : r == 4 ? (g1 ? 16u : 0u)
         : r == 5 ? (s0 ? 4u : 0u) : r == 6 ? (s1 ? 4u : 0u) : (xpbd ? 4u : 0u);
}
template <int T, int R> __device__ __forceinline__ const unsigned char *stream_ptr(const TypeArrays &ta) {
    constexpr bool two = (T == PBD_DISTANCE || T == PBD_DISTANCE_XPBD), tri = (T == PBD_FEMTRIANGLE || T == PBD_STRAINTRIANGLE);
    const void *p = nullptr;
    if (R == 0) p = two ? (const void *)ta.idx2 : (tri ? (const void *)ta.idx3[0] : (const void *)ta.idx4);
    else if (R == 1) p = ta.idx3[1];
    else if (R == 2) p = ta.idx3[2];
    else if (R == 3) p = ta.gv[0];
    else if (R == 4) p = ta.gv[1];
    else if (R == 5) p = ta.gs[0];
    else if (R == 6) p = ta.gs[1];
    else p = ta.lambda;
    return static_cast<const unsigned char *>(p);
}

// Where the run [g0, g0 + n) of a type lands in the stage buffer.  Every array of the run is copied from its 16-byte
// aligned start (head = misalignment of element g0) with a size rounded up to 16; `staged` items fit, the rest of the run is
// read from global memory.  Deterministic in its arguments: the manager (who copies) and the workers (who read) agree.
struct RunPlan { unsigned o0, o1, o2, o3, o4, o5, o6, o7; unsigned staged; };
template <int R> __device__ __forceinline__ unsigned &plan_off(RunPlan &pl) {
    if (R == 0) return pl.o0; if (R == 1) return pl.o1; if (R == 2) return pl.o2; if (R == 3) return pl.o3;
    if (R == 4) return pl.o4; if (R == 5) return pl.o5; if (R == 6) return pl.o6; return pl.o7;
}
template <int T> __host__ __device__ constexpr unsigned stream_item_bytes() {
    return stream_es<T>(0) + stream_es<T>(1) + stream_es<T>(2) + stream_es<T>(3) + stream_es<T>(4) + stream_es<T>(5) + stream_es<T>(6) + stream_es<T>(7);
}
template <int T> __host__ __device__ constexpr unsigned stream_slack() {
    unsigned s = 0;
    for (int r = 0; r < kStreamArrays; r++) if (stream_es<T>(r)) s += 32u;
    return s;
}
template <int T, int R> __device__ __forceinline__ void plan_one(unsigned g0, unsigned m, unsigned &running, RunPlan &pl) {
    constexpr unsigned es = stream_es<T>(R);
    plan_off<R>(pl) = 0;
    if (es == 0 || m == 0) return;
    const unsigned head = (g0 * es) & 15u;
    plan_off<R>(pl) = running + head;
    running += (head + m * es + 15u) & ~15u;
}
template <int T>
__device__ __forceinline__ void plan_run(unsigned g0, unsigned n, unsigned &running, RunPlan &pl) {
    constexpr unsigned perItem = stream_item_bytes<T>(), slack = stream_slack<T>();
    const unsigned avail = kStageBytes - running;
    unsigned m = n;
    if ((unsigned long long)n * perItem + slack > avail) m = (avail > slack) ? (avail - slack) / perItem : 0u;
    pl.staged = m;
    plan_one<T, 0>(g0, m, running, pl); plan_one<T, 1>(g0, m, running, pl); plan_one<T, 2>(g0, m, running, pl); plan_one<T, 3>(g0, m, running, pl);
    plan_one<T, 4>(g0, m, running, pl); plan_one<T, 5>(g0, m, running, pl); plan_one<T, 6>(g0, m, running, pl); plan_one<T, 7>(g0, m, running, pl);
}
// manager lane R copies stream slot R of the run
template <int T, int R>
__device__ __forceinline__ void issue_one(const TypeArrays &ta, unsigned g0, RunPlan &pl, unsigned lane, unsigned dst0, unsigned bar) {
    constexpr unsigned es = stream_es<T>(R);
    if (es == 0 || pl.staged == 0 || lane != (unsigned)R) return;
    const unsigned head = (g0 * es) & 15u;
    const unsigned bytes = (head + pl.staged * es + 15u) & ~15u;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(dst0 + plan_off<R>(pl) - head, stream_ptr<T, R>(ta) + (size_t)g0 * es - head, bytes, bar);
}

template <int T>
__device__ __forceinline__ Streamed load_streamed_stage(const unsigned char *stage, const RunPlan &pl, unsigned j) {
    Streamed s;
    s.g0 = make_float4(0.f, 0.f, 0.f, 0.f); s.g1 = s.g0; s.s0 = 0.0f; s.s1 = 0.0f;
    constexpr bool two = (T == PBD_DISTANCE || T == PBD_DISTANCE_XPBD), tri = (T == PBD_FEMTRIANGLE || T == PBD_STRAINTRIANGLE);
    if (two) { const uint2 b = *reinterpret_cast<const uint2 *>(stage + pl.o0 + 8u * j); s.b = make_uint4(b.x, b.y, 0u, 0u); }
    else if (tri) s.b = make_uint4(*reinterpret_cast<const unsigned *>(stage + pl.o0 + 4u * j), *reinterpret_cast<const unsigned *>(stage + pl.o1 + 4u * j),
                                   *reinterpret_cast<const unsigned *>(stage + pl.o2 + 4u * j), 0u);
    else s.b = *reinterpret_cast<const uint4 *>(stage + pl.o0 + 16u * j);
    constexpr bool g0 = tri || T == PBD_ISOBENDING || T == PBD_ISOBENDING_XPBD || T == PBD_FEMTET || T == PBD_FEMTET_XPBD || T == PBD_STRAINTET;
    constexpr bool g1 = (T == PBD_FEMTET || T == PBD_FEMTET_XPBD || T == PBD_STRAINTET);
    constexpr bool s0 = two || T == PBD_FEMTRIANGLE || T == PBD_DIHEDRAL || T == PBD_VOLUME || T == PBD_VOLUME_XPBD || g1;
    constexpr bool s1 = (T == PBD_FEMTET || T == PBD_FEMTET_XPBD);
    if (g0) s.g0 = *reinterpret_cast<const float4 *>(stage + pl.o3 + 16u * j);
    if (g1) s.g1 = *reinterpret_cast<const float4 *>(stage + pl.o4 + 16u * j);
    if (s0) s.s0 = *reinterpret_cast<const float *>(stage + pl.o5 + 4u * j);
    if (s1) s.s1 = *reinterpret_cast<const float *>(stage + pl.o6 + 4u * j);
    return s;
}

// named barrier 1: the workers announce "my spanning constraints are done" without waiting; the manager warp collects
template <int THREADS> __device__ __forceinline__ void cta_arrive() { asm volatile("bar.arrive 1, %0;" :: "n"(THREADS) : "memory"); }
template <int THREADS> __device__ __forceinline__ void cta_collect() { asm volatile("bar.sync 1, %0;" :: "n"(THREADS) : "memory"); }

template <unsigned MASK, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_step_tiled(const __grid_constant__ TiledArgs a) {
    extern __shared__ __align__(128) unsigned char tiledSmem[];
    float4 *sp = reinterpret_cast<float4 *>(tiledSmem);
    unsigned char *stageBuf = tiledSmem + (size_t)kTileCapacity * sizeof(float4);
    const unsigned barAddr = smem_u32(stageBuf + 2u * kStageBytes);  // two mbarriers (8 bytes each)
    constexpr unsigned W = THREADS - 32;                                // worker threads; the last warp is the manager
    const unsigned t = threadIdx.x, tile = blockIdx.x, nTiles = gridDim.x;
    const bool manager = (t >= W);
    const unsigned lane = t & 31u;
    const unsigned p0 = __ldg(a.tileStart + tile), p1 = __ldg(a.tileStart + tile + 1), nPriv = __ldg(a.tilePrivate + tile);
    const unsigned offStride = 2u * nTiles + 1u;
    const TileAcc acc{a.pos, sp};
    unsigned long long target = a.barrierBase;
    unsigned phase = 0;  // colour phases executed so far in this launch

    if (t == 0) { mbar_init(barAddr, 1); mbar_init(barAddr + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();

    // manager: stream the runs of the colour that starts at bucket `bs` into stage buffer (ph & 1)
    auto issue_phase = [&](unsigned bs, unsigned ph) {
        const unsigned bar = barAddr + 8u * (ph & 1u);
        const unsigned dst0 = smem_u32(stageBuf + (size_t)(ph & 1u) * kStageBytes);
        const unsigned colour = __ldg(&a.buckets[bs].colour);
        unsigned running = 0;
#pragma unroll 1
        for (unsigned k = bs; k < a.nBuckets && __ldg(&a.buckets[k].colour) == colour; k++) {
            const Bucket b = load_bucket(a.buckets, k);
            const unsigned *off = a.tileOff + (size_t)k * offStride + 2u * tile;
            const unsigned c0 = __ldg(off), c2 = __ldg(off + 2);
            if (c2 == c0) continue;
            PBD_FOR_TYPE(MASK, b.type,
                RunPlan pl;
                const unsigned g0 = b.first + c0;
                const TypeArrays &ta = a.types[T];
                plan_run<T>(g0, c2 - c0, running, pl);
                issue_one<T, 0>(ta, g0, pl, lane, dst0, bar); issue_one<T, 1>(ta, g0, pl, lane, dst0, bar); issue_one<T, 2>(ta, g0, pl, lane, dst0, bar);
                issue_one<T, 3>(ta, g0, pl, lane, dst0, bar); issue_one<T, 4>(ta, g0, pl, lane, dst0, bar); issue_one<T, 5>(ta, g0, pl, lane, dst0, bar);
                issue_one<T, 6>(ta, g0, pl, lane, dst0, bar); issue_one<T, 7>(ta, g0, pl, lane, dst0, bar);)
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
    };

    if (manager && a.stage && a.nBuckets) issue_phase(0u, 0u);

    for (unsigned sub = 0; sub < a.subSteps; sub++) {
        // ---- prologue on the tile's own particles: lastX = oldX; oldX = x; semi-implicit Euler; private ones go to smem -------
        for (unsigned i = p0 + t; i < p1; i += THREADS) {
            float4 x = __ldcg(a.pos + i);
            if (a.trackLast) __stcs(a.lastp + i, __ldcs(a.oldp + i));
            __stcg(a.oldp + i, x);
            float4 v = __ldcs(a.vel + i);
            if (v.w != 0.0f) {
                v.x = fmaf(a.gx, a.h, v.x); v.y = fmaf(a.gy, a.h, v.y); v.z = fmaf(a.gz, a.h, v.z);
                x.x = fmaf(v.x, a.h, x.x); x.y = fmaf(v.y, a.h, x.y); x.z = fmaf(v.z, a.h, x.z);
                __stcs(a.vel + i, v);
            }
            if (i - p0 < nPriv) sp[a.swizzle ? tile_swizzle(i - p0) : i - p0] = x;  // private: lives in shared memory for the whole substep
            else if (v.w != 0.0f) __stcg(a.pos + i, x);  // shared: other CTAs read it through L2
        }
        if (nTiles > 1) grid_barrier(a.barrier, target);  // shared particles integrated everywhere (also orders the smem tile inside the CTA)
        else __syncthreads();

        // ---- coloured Gauss-Seidel sweeps: one split grid-barrier phase per colour -------------------------------------------------
        for (unsigned it = 0; it < a.maxIter; it++) {
            const bool iterZero = (it == 0);
            unsigned bi = 0;
            while (bi < a.nBuckets) {
                const unsigned colour = __ldg(&a.buckets[bi].colour);
                unsigned be = bi + 1;
                while (be < a.nBuckets && __ldg(&a.buckets[be].colour) == colour) be++;
                const bool lastPhase = (be == a.nBuckets) && (it + 1 == a.maxIter) && (sub + 1 == a.subSteps);
                target += nTiles;
                if (manager) {
                    const bool tr = a.trace && a.traceWorker < 0 && phase < a.tracePhases && lane == 0;
                    unsigned long long *rec = a.trace + ((size_t)phase * nTiles + tile) * 4;
                    cta_collect<THREADS>();                      // every worker's spanning constraints are done
                    if (tr) rec[0] = globaltimer_ns();
                    if (lane == 0 && nTiles > 1) {
                        // release: the workers' stores were ordered before this thread by the named barrier; every particle /
                        // multiplier access of this kernel is an L2 access (ld.cg / st.cg), so no L1 invalidation is needed
                        asm volatile("fence.acq_rel.gpu;" ::: "memory");
                        asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" :: "l"(a.barrier), "l"(1ull) : "memory");
                        if (tr) rec[1] = globaltimer_ns();
                    }
                    __syncwarp();
                    // while the arrival travels: start streaming the next colour (its buffer was released by the sync that ended
                    // phase - 1).  Issued after the fence on purpose: the fence would wait for the copies.
                    if (a.stage && !lastPhase) {
                        // the multipliers of the next colour were written with generic stores (by the sync that ended phase - 1 at the
                        // latest); the bulk copy reads them through the async proxy
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                        issue_phase(be < a.nBuckets ? be : 0u, phase + 1u);
                    }
                    if (lane == 0 && nTiles > 1) {
                        while (ld_acquire_u64(a.barrier) < target) { }
                        if (tr) rec[2] = globaltimer_ns();
                    }
                    __syncwarp();
                    __syncthreads();                             // end of the colour (the workers join after their private constraints)
                    if (tr) rec[3] = globaltimer_ns();
                } else {
                    const unsigned char *stage = stageBuf + (size_t)(phase & 1u) * kStageBytes;
                    const bool tr = a.trace && a.traceWorker >= 0 && phase < a.tracePhases && t == 32u * (unsigned)a.traceWorker;
                    unsigned long long *rec = a.trace + ((size_t)phase * nTiles + tile) * 4;
                    if (tr) rec[0] = globaltimer_ns();
                    if (a.stage) mbar_wait(barAddr + 8u * (phase & 1u), (phase >> 1) & 1u);
                    if (tr) rec[1] = globaltimer_ns();
                    unsigned rot = 0;  // items of this colour handed out so far (mod W): the next run starts at that thread
#pragma unroll 1
                    for (int part = 0; part < 2; part++) {
                        unsigned running = 0;
#pragma unroll 1
                        for (unsigned k = bi; k < be; k++) {
                            const Bucket b = load_bucket(a.buckets, k);
                            const unsigned *off = a.tileOff + (size_t)k * offStride + 2u * tile;
                            const unsigned c0 = __ldg(off), c1 = __ldg(off + 1), c2 = __ldg(off + 2);
                            if (c2 == c0) continue;
                            const unsigned j0 = part ? c1 - c0 : 0u, j1 = part ? c2 - c0 : c1 - c0;  // relative to the run start
                            PBD_FOR_TYPE(MASK, b.type,
                                const TypeArrays &ta = a.types[T];
                                const unsigned g0 = b.first + c0;
                                RunPlan pl;
                                if (a.stage) plan_run<T>(g0, c2 - c0, running, pl); else pl.staged = 0;
                                for (unsigned j = j0 + ((t + W - rot) % W); j < j1; j += W) {
                                    const bool st = (j < pl.staged);
                                    const Streamed s = st ? load_streamed_stage<T>(stage, pl, j) : load_streamed<T>(ta, g0 + j);
                                    const float *lam = (st && a.stageLambda && stream_es<T>(7)) ? reinterpret_cast<const float *>(stage + pl.o7 + 4u * j) : nullptr;
                                    project_streamed_acc<T>(acc, ta, g0 + j, s, a.h, iterZero, lam);
                                })
                            rot = (rot + (j1 - j0)) % W;
                        }
                        if (tr) rec[2 + part] = globaltimer_ns();
                        if (part == 0) cta_arrive<THREADS>();
                    }
                    __syncthreads();                             // end of the colour: released by the manager after the grid barrier
                }
                phase++;
                bi = be;
            }
        }
        if (a.nBuckets == 0) __syncthreads();

        // ---- epilogue on the tile's own particles: write the private ones back, velocity update ----------------------------------
        for (unsigned i = p0 + t; i < p1; i += THREADS) {
            float4 x;
            if (i - p0 < nPriv) { x = sp[a.swizzle ? tile_swizzle(i - p0) : i - p0]; if (x.w != 0.0f) __stcg(a.pos + i, x); }
            else x = __ldcg(a.pos + i);
            if (x.w == 0.0f) continue;
            const float4 o = __ldcg(a.oldp + i);
            float4 v = __ldcs(a.vel + i);
            if (!a.secondOrder) {
                v.x = a.invH * (x.x - o.x); v.y = a.invH * (x.y - o.y); v.z = a.invH * (x.z - o.z);
            } else {
                const float4 l = __ldcs(a.lastp + i);
                v.x = a.invH * (1.5f * x.x - 2.0f * o.x + 0.5f * l.x);
                v.y = a.invH * (1.5f * x.y - 2.0f * o.y + 0.5f * l.y);
                v.z = a.invH * (1.5f * x.z - 2.0f * o.z + 0.5f * l.z);
            }
            __stcs(a.vel + i, v);
        }
        __syncthreads();  // the next substep's prologue rewrites sp[]
    }
}

}  // namespace pbdk
