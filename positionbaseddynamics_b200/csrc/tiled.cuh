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
// tile by tile; tileOff[bucket * (nTiles + 1) + t] .. [t + 1] is the range CTA t executes.
#pragma once
#include "persistent.cuh"

namespace pbdk {

constexpr unsigned kSmemFlag = 0x80000000u;
constexpr int kTileCapacity = 12288;  // private particles per tile kept in shared memory (192 KB)
constexpr size_t kTiledSmemBytes = (size_t)kTileCapacity * sizeof(float4);

struct TiledArgs {
    float4 *pos, *vel, *oldp, *lastp;
    const Bucket *buckets;
    const unsigned *tileOff;     // [nBuckets][nTiles + 1] offsets relative to the bucket's `first`
    const unsigned *tileStart;   // [nTiles + 1] device slots: tile t owns [tileStart[t], tileStart[t+1])
    const unsigned *tilePrivate; // [nTiles] number of leading slots of the tile that live in shared memory
    unsigned nBuckets, subSteps, maxIter;
    float h, invH, gx, gy, gz;
    int secondOrder, trackLast;
    unsigned long long *barrier;
    unsigned long long barrierBase;
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

template <unsigned MASK, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_step_tiled(const __grid_constant__ TiledArgs a) {
    extern __shared__ float4 sp[];
    const unsigned t = threadIdx.x, NT = blockDim.x, tile = blockIdx.x, nTiles = gridDim.x;
    const unsigned p0 = __ldg(a.tileStart + tile), p1 = __ldg(a.tileStart + tile + 1), nPriv = __ldg(a.tilePrivate + tile);
    const TileAcc acc{a.pos, sp};
    unsigned long long target = a.barrierBase;

    for (unsigned sub = 0; sub < a.subSteps; sub++) {
        // ---- prologue on the tile's own particles: lastX = oldX; oldX = x; semi-implicit Euler; private ones go to smem -------
        for (unsigned i = p0 + t; i < p1; i += NT) {
            float4 x = __ldcg(a.pos + i);
            if (a.trackLast) __stcs(a.lastp + i, __ldcs(a.oldp + i));
            __stcg(a.oldp + i, x);
            float4 v = __ldcs(a.vel + i);
            if (v.w != 0.0f) {
                v.x = fmaf(a.gx, a.h, v.x); v.y = fmaf(a.gy, a.h, v.y); v.z = fmaf(a.gz, a.h, v.z);
                x.x = fmaf(v.x, a.h, x.x); x.y = fmaf(v.y, a.h, x.y); x.z = fmaf(v.z, a.h, x.z);
                __stcs(a.vel + i, v);
            }
            if (i - p0 < nPriv) sp[i - p0] = x;       // private: lives in shared memory for the whole substep
            else if (v.w != 0.0f) __stcg(a.pos + i, x);  // shared: other CTAs read it through L2
        }
        grid_barrier(a.barrier, target);  // shared particles integrated everywhere (also orders the smem tile inside the CTA)

        // ---- coloured Gauss-Seidel sweeps -------------------------------------------------------------------------------------
        for (unsigned it = 0; it < a.maxIter; it++) {
            const bool iterZero = (it == 0);
            for (unsigned bi = 0; bi < a.nBuckets; bi++) {
                const Bucket b = load_bucket(a.buckets, bi);
                const unsigned c0 = __ldg(a.tileOff + (size_t)bi * (nTiles + 1) + tile), c1 = __ldg(a.tileOff + (size_t)bi * (nTiles + 1) + tile + 1);
                PBD_FOR_TYPE(MASK, b.type,
                    const TypeArrays &ta = a.types[T];
                    for (unsigned i = c0 + t; i < c1; i += blockDim.x) {
                        const Streamed s = load_streamed<T>(ta, b.first + i);
                        project_streamed_acc<T>(acc, ta, b.first + i, s, a.h, iterZero);
                    })
                const bool lastOfSweep = (bi + 1 == a.nBuckets);
                const unsigned nextColour = lastOfSweep ? 0xffffffffu : __ldg(&a.buckets[bi + 1].colour);
                if (nextColour != b.colour) grid_barrier(a.barrier, target);  // colour boundary (and end of sweep)
            }
        }
        if (a.nBuckets == 0) __syncthreads();

        // ---- epilogue on the tile's own particles: write the private ones back, velocity update ----------------------------------
        for (unsigned i = p0 + t; i < p1; i += NT) {
            float4 x;
            if (i - p0 < nPriv) { x = sp[i - p0]; if (x.w != 0.0f) __stcg(a.pos + i, x); }
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
