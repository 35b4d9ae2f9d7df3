// positionbaseddynamics_b200/csrc/persistent.cuh
//
// k_step_persistent: one TimeStepController::step (Simulation/TimeStepController.cpp:75-241, particle part) as a single
// cooperative launch.  The grid is sized to be co-resident (SM count x occupancy); every CTA walks the same bucket list
//     per substep:  integrate | barrier | maxIter x ( colour 0 | barrier | colour 1 | ... | barrier ) | velocity update
// and a grid-wide barrier separates the colour phases (the reference's "for group in groups" loop,
// TimeStepController.cpp:272-286, is sequential over colours and parallel inside one).  Buckets of different types inside
// one colour touch disjoint particles and need no barrier between them.  The velocity update of particle i and the next
// substep's integration of particle i are done by the same thread, so no barrier is needed between them either.
//
// Barrier: monotonically increasing 64-bit arrival counter in global memory (one atomic per CTA per phase, thread 0
// spins with ld.acquire.gpu).  Particle data is read/written with .cg (L2) accesses, so no L1 line can go stale
// across phases; the __threadfence() before the arrival publishes the CTA's stores.
#pragma once
#include "kernels.cuh"

namespace pbdk {

constexpr int kPersistentThreads = 512;

struct PersistentArgs {
    float4 *pos, *vel, *oldp, *lastp;
    unsigned n;
    const TypeArrays *types;  // [PBD_NUM_TYPES] in global memory
    const Bucket *buckets;
    unsigned nBuckets, subSteps, maxIter;
    float h, invH, gx, gy, gz;
    int secondOrder, trackLast;
    unsigned long long *barrier;
    unsigned long long barrierBase;  // counter value when this launch starts
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
        __threadfence();  // acquire side: gpu-scope fence also drops this SM's L1 lines (stale particle data of the last phase)
    }
    __syncthreads();
}

template <int T, bool CA>
__device__ __forceinline__ void sweep_bucket(float4 *pos, const TypeArrays &ta, const Bucket &b, float h, bool iterZero,
                                             unsigned tid, unsigned stride) {
    for (unsigned i = tid; i < b.count; i += stride) process_constraint<T, CA>(pos, ta, b.first + i, h, iterZero);
}

template <bool CA>
__global__ void __launch_bounds__(kPersistentThreads) k_step_persistent(PersistentArgs a) {
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
        grid_barrier(a.barrier, target);

        // ---- coloured Gauss-Seidel sweeps -------------------------------------------------------------------
        for (unsigned it = 0; it < a.maxIter; it++) {
            const bool iterZero = (it == 0);
            unsigned colour = a.nBuckets ? __ldg(&a.buckets[0].colour) : 0u;
            for (unsigned bi = 0; bi < a.nBuckets; bi++) {
                Bucket b;
                b.type = __ldg(&a.buckets[bi].type); b.first = __ldg(&a.buckets[bi].first);
                b.count = __ldg(&a.buckets[bi].count); b.colour = __ldg(&a.buckets[bi].colour);
                if (b.colour != colour) { grid_barrier(a.barrier, target); colour = b.colour; }
                const TypeArrays ta = a.types[b.type];
                switch (b.type) {
#define SB(T) case T: sweep_bucket<T, CA>(a.pos, ta, b, a.h, iterZero, tid, stride); break;
                    SB(PBD_DISTANCE) SB(PBD_DISTANCE_XPBD) SB(PBD_DIHEDRAL) SB(PBD_ISOBENDING) SB(PBD_ISOBENDING_XPBD)
                    SB(PBD_FEMTRIANGLE) SB(PBD_STRAINTRIANGLE) SB(PBD_VOLUME) SB(PBD_VOLUME_XPBD) SB(PBD_FEMTET)
                    SB(PBD_FEMTET_XPBD) SB(PBD_STRAINTET)
#undef SB
                default: break;
                }
            }
            grid_barrier(a.barrier, target);
        }

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
