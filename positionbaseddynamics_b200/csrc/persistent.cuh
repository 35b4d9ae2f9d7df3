// positionbaseddynamics_b200/csrc/persistent.cuh
//
// k_step_persistent: one TimeStepController::step (Simulation/TimeStepController.cpp:75-241, particle part) as a single
// cooperative launch.  The grid is co-resident (one CTA per SM); every CTA walks the same phase list
//     per substep:  integrate | B | maxIter x ( colour 0 | B | colour 1 | B | ... ) | B | velocity update
// with a grid-wide barrier B between colour phases only (the reference's "for group in groups" loop,
// TimeStepController.cpp:272-286, is sequential over colours and parallel inside one).  Buckets of different types inside
// one colour touch disjoint particles and need no barrier between them; the velocity update of particle i and the next
// substep's integration of particle i are done by the same thread, so no barrier separates them either.
//
// Latency hiding: the indices and rest data a thread needs for its first constraint of the NEXT phase are streamed from
// HBM *before* it arrives at the barrier (they never change during a step), and inside a phase the loop over a thread's
// constraints is software-pipelined the same way.  After the barrier only the L2-resident particle gather, the arithmetic
// and the scatter remain on the critical path.
//
// Specialisation: the kernel is instantiated for a few masks of constraint types (cloth, tets, everything); types outside
// the mask are compiled out, which keeps the light instantiations at <= 64 registers so that 1024 threads fit on an SM.
//
// Barrier: monotonically increasing 64-bit arrival counter in global memory (one atomic per CTA per phase, thread 0
// spins with ld.acquire.gpu).  The gpu-scope fence before the arrival publishes the CTA's stores, the one after the
// spin drops the SM's L1 lines so that cached particle gathers (ld.ca) can never see data of an earlier phase.
#pragma once
#include "kernels.cuh"

namespace pbdk {

struct PersistentArgs {
    float4 *pos, *vel, *oldp, *lastp;
    unsigned n;
    const Bucket *buckets;
    unsigned nBuckets, subSteps, maxIter;
    float h, invH, gx, gy, gz;
    int secondOrder, trackLast;
    unsigned long long *barrier;
    unsigned long long barrierBase;  // counter value when this launch starts
    TypeArrays types[PBD_NUM_TYPES]; // by value: lives in the kernel's constant bank, indexed with compile-time T
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void grid_barrier(unsigned long long *counter, unsigned long long &target) {
    target += gridDim.x;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();  // release: publish this CTA's stores
        atomicAdd(counter, 1ull);
        while (ld_acquire_u64(counter) < target) { }
        __threadfence();  // acquire side: gpu-scope fence also invalidates this SM's L1
    }
    __syncthreads();
}

constexpr unsigned type_bit(int t) { return 1u << t; }
constexpr unsigned kMaskClothXPBD = type_bit(PBD_DISTANCE_XPBD) | type_bit(PBD_ISOBENDING_XPBD);
constexpr unsigned kMaskLight = type_bit(PBD_DISTANCE) | type_bit(PBD_DISTANCE_XPBD) | type_bit(PBD_DIHEDRAL) | type_bit(PBD_ISOBENDING) |
                                type_bit(PBD_ISOBENDING_XPBD) | type_bit(PBD_VOLUME) | type_bit(PBD_VOLUME_XPBD) | type_bit(PBD_FEMTRIANGLE);
constexpr unsigned kMaskAll = (1u << PBD_NUM_TYPES) - 1u;

// dispatch a functor on the runtime type, restricted to the compiled-in mask
#define PBD_FOR_TYPE(MASK, type, ...)                                                                                  \
    switch (type) {                                                                                                      \
    case PBD_DISTANCE:        if constexpr ((MASK) & type_bit(PBD_DISTANCE))        { constexpr int T = PBD_DISTANCE;        __VA_ARGS__ } break; \
    case PBD_DISTANCE_XPBD:   if constexpr ((MASK) & type_bit(PBD_DISTANCE_XPBD))   { constexpr int T = PBD_DISTANCE_XPBD;   __VA_ARGS__ } break; \
    case PBD_DIHEDRAL:        if constexpr ((MASK) & type_bit(PBD_DIHEDRAL))        { constexpr int T = PBD_DIHEDRAL;        __VA_ARGS__ } break; \
    case PBD_ISOBENDING:      if constexpr ((MASK) & type_bit(PBD_ISOBENDING))      { constexpr int T = PBD_ISOBENDING;      __VA_ARGS__ } break; \
    case PBD_ISOBENDING_XPBD: if constexpr ((MASK) & type_bit(PBD_ISOBENDING_XPBD)) { constexpr int T = PBD_ISOBENDING_XPBD; __VA_ARGS__ } break; \
    case PBD_FEMTRIANGLE:     if constexpr ((MASK) & type_bit(PBD_FEMTRIANGLE))     { constexpr int T = PBD_FEMTRIANGLE;     __VA_ARGS__ } break; \
    case PBD_STRAINTRIANGLE:  if constexpr ((MASK) & type_bit(PBD_STRAINTRIANGLE))  { constexpr int T = PBD_STRAINTRIANGLE;  __VA_ARGS__ } break; \
    case PBD_VOLUME:          if constexpr ((MASK) & type_bit(PBD_VOLUME))          { constexpr int T = PBD_VOLUME;          __VA_ARGS__ } break; \
    case PBD_VOLUME_XPBD:     if constexpr ((MASK) & type_bit(PBD_VOLUME_XPBD))     { constexpr int T = PBD_VOLUME_XPBD;     __VA_ARGS__ } break; \
    case PBD_FEMTET:          if constexpr ((MASK) & type_bit(PBD_FEMTET))          { constexpr int T = PBD_FEMTET;          __VA_ARGS__ } break; \
    case PBD_FEMTET_XPBD:     if constexpr ((MASK) & type_bit(PBD_FEMTET_XPBD))     { constexpr int T = PBD_FEMTET_XPBD;     __VA_ARGS__ } break; \
    case PBD_STRAINTET:       if constexpr ((MASK) & type_bit(PBD_STRAINTET))       { constexpr int T = PBD_STRAINTET;       __VA_ARGS__ } break; \
    default: break;                                                                                                      \
    }

__device__ __forceinline__ Bucket load_bucket(const Bucket *buckets, unsigned bi) {
    Bucket b;
    const int4 raw = __ldg(reinterpret_cast<const int4 *>(buckets) + bi);  // Bucket is 16 bytes
    b.type = raw.x; b.first = (unsigned)raw.y; b.count = (unsigned)raw.z; b.colour = (unsigned)raw.w;
    return b;
}

template <unsigned MASK, bool CA, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_step_persistent(const __grid_constant__ PersistentArgs a) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned stride = gridDim.x * blockDim.x;
    unsigned long long target = a.barrierBase;

    for (unsigned sub = 0; sub < a.subSteps; sub++) {
        // ---- prologue: lastX = oldX; oldX = x; semi-implicit Euler ------------------------------------------
        for (unsigned i = tid; i < a.n; i += stride) {
            float4 x = __ldcg(a.pos + i);
            if (a.trackLast) __stcs(a.lastp + i, __ldcs(a.oldp + i));
            __stcg(a.oldp + i, x);
            float4 v = __ldcs(a.vel + i);
            if (v.w != 0.0f) {
                v.x = fmaf(a.gx, a.h, v.x); v.y = fmaf(a.gy, a.h, v.y); v.z = fmaf(a.gz, a.h, v.z);
                x.x = fmaf(v.x, a.h, x.x); x.y = fmaf(v.y, a.h, x.y); x.z = fmaf(v.z, a.h, x.z);
                __stcs(a.vel + i, v);
                __stcg(a.pos + i, x);
            }
        }

        // ---- coloured Gauss-Seidel sweeps -------------------------------------------------------------------
        // `pre` holds the streamed part of this thread's first constraint of the bucket about to be processed.
        Streamed pre;
        pre.b = make_uint4(0u, 0u, 0u, 0u); pre.g0 = make_float4(0.f, 0.f, 0.f, 0.f); pre.g1 = pre.g0; pre.s0 = 0.f; pre.s1 = 0.f;
        Bucket nb;
        nb.type = -1; nb.first = 0u; nb.count = 0u; nb.colour = 0u;
        if (a.nBuckets) {
            nb = load_bucket(a.buckets, 0);
            if (tid < nb.count) { PBD_FOR_TYPE(MASK, nb.type, pre = load_streamed<T>(a.types[T], nb.first + tid);) }
        }
        grid_barrier(a.barrier, target);  // integration done everywhere before the first colour reads positions

        for (unsigned it = 0; it < a.maxIter; it++) {
            const bool iterZero = (it == 0);
            for (unsigned bi = 0; bi < a.nBuckets; bi++) {
                const Bucket b = nb;
                // bucket that follows this one in execution order (wraps into the next sweep)
                const bool lastOfSweep = (bi + 1 == a.nBuckets);
                const bool more = !lastOfSweep || (it + 1 < a.maxIter);
                if (more) nb = load_bucket(a.buckets, lastOfSweep ? 0u : bi + 1);
                // software-pipelined walk over this thread's constraints of bucket b
                PBD_FOR_TYPE(MASK, b.type,
                    const TypeArrays &ta = a.types[T];
                    Streamed cur = pre;
                    for (unsigned i = tid; i < b.count; i += stride) {
                        Streamed nxt = cur;
                        if (i + stride < b.count) nxt = load_streamed<T>(ta, b.first + i + stride);
                        project_streamed<T, CA>(a.pos, ta, b.first + i, cur, a.h, iterZero);
                        cur = nxt;
                    })
                // stream the first constraint of the next bucket, then synchronise if it starts a new colour phase
                if (more) {
                    if (tid < nb.count) { PBD_FOR_TYPE(MASK, nb.type, pre = load_streamed<T>(a.types[T], nb.first + tid);) }
                    if (nb.colour != b.colour || lastOfSweep) grid_barrier(a.barrier, target);
                }
            }
        }
        if (a.nBuckets) grid_barrier(a.barrier, target);  // all projections done before velocities are derived

        // ---- epilogue: velocity update ----------------------------------------------------------------------
        for (unsigned i = tid; i < a.n; i += stride) {
            const float4 x = __ldcg(a.pos + i);
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
    }
}

}  // namespace pbdk
