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
// Why shared-memory staging.  ncu on the per-bucket kernels shows the path is latency bound, not bandwidth bound (SM busy
// 17 %, DRAM 25 %, L2 26 %, 82 % of issue slots without an eligible warp, long-scoreboard stalls on the particle
// gathers): a colour phase is only ~2e5 short dependent chains  index -> gather -> ~200 instructions -> scatter.  With the
// tuple held in registers the number of gathers in flight is capped by the register file.  Here every thread instead
//   1. streams the indices / rest data / multipliers of ALL its constraints of the next phase into shared memory with
//      cp.async (LDGSTS) *before* arriving at the barrier (these never depend on the previous phase),
//   2. after the barrier issues the particle gathers of all those constraints as cp.async.cg 16-byte copies into shared
//      memory (one commit group per constraint, nothing held in registers, L2 -> SMEM bypassing L1, so no stale lines),
//   3. projects constraint k as soon as its group has landed (cp.async.wait_group), scattering with st.global.cg.
// The whole phase therefore exposes one L2 round trip instead of one per constraint, and the barrier latency overlaps the
// HBM stream of step 1.
//
// Specialisation: instantiated for a few masks of constraint types (cloth, light, everything); types outside the mask are
// compiled out so that the light instantiations stay within 64 registers (1024 threads per SM).
//
// Barrier: monotonically increasing 64-bit arrival counter in global memory (one atomic per CTA per phase, thread 0
// spins with ld.acquire.gpu, gpu-scope fence before the arrival publishes the CTA's stores).
#pragma once
#include <utility>
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
    unsigned long long *trace;       // development aid (PBD_B200_TRACE): per-CTA globaltimer stamps of the first phases, or nullptr
    unsigned tracePhases;
    TypeArrays types[PBD_NUM_TYPES]; // by value: lives in the kernel's constant bank, indexed with compile-time T
};

// shared-memory budget per CTA (one CTA per SM): gather area, streamed-operand area, multiplier area
constexpr int kGatherF4 = 8192;  // 128 KB of particle float4s
constexpr int kStreamF4 = 4096;  //  64 KB of indices / rest data
constexpr int kLambdaF = 4096;   //  16 KB of XPBD multipliers
constexpr size_t kPersistentSmemBytes = (size_t)kGatherF4 * 16 + (size_t)kStreamF4 * 16 + (size_t)kLambdaF * 4;

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}
// trace record per (phase, CTA): [0] barrier exit of the previous phase is [3] of phase-1; [0] warp 0 done, [1] CTA done
// (after __syncthreads), [2] arrived (after fence+atomic), [3] released
__device__ __forceinline__ void grid_barrier(unsigned long long *counter, unsigned long long &target, unsigned long long *trace = nullptr,
                                             unsigned tracePhases = 0, unsigned phase = 0) {
    target += gridDim.x;
    const bool tr = trace && phase < tracePhases && threadIdx.x == 0;
    unsigned long long *rec = trace + ((size_t)phase * gridDim.x + blockIdx.x) * 4;
    if (tr) rec[0] = globaltimer_ns();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (tr) rec[1] = globaltimer_ns();
        __threadfence();  // release: publish this CTA's stores
        atomicAdd(counter, 1ull);
        if (tr) rec[2] = globaltimer_ns();
        while (ld_acquire_u64(counter) < target) { }
        if (tr) rec[3] = globaltimer_ns();
    }
    __syncthreads();
}

// ---- cp.async helpers (LDGSTS) -----------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16_cg(void *dst, const void *src) {  // L2 -> SMEM, bypasses L1
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <class F, int... Is>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, Is...>) { (f(std::integral_constant<int, Is>{}), ...); }
template <int N, class F> __device__ __forceinline__ void static_for(F &&f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }

// per-type shape of the staged data: bodies, 16-byte streamed slots per constraint
template <int T> struct Staged {
    static constexpr int NB = (T == PBD_DISTANCE || T == PBD_DISTANCE_XPBD) ? 2 : ((T == PBD_FEMTRIANGLE || T == PBD_STRAINTRIANGLE) ? 3 : 4);
    static constexpr int NS = (T == PBD_DISTANCE || T == PBD_DISTANCE_XPBD || T == PBD_SHAPEMATCHING) ? 1 : ((T == PBD_FEMTET || T == PBD_FEMTET_XPBD || T == PBD_STRAINTET) ? 4 : 2);
    static constexpr bool XPBD = (T == PBD_DISTANCE_XPBD || T == PBD_VOLUME_XPBD || T == PBD_ISOBENDING_XPBD || T == PBD_FEMTET_XPBD);
};
template <int T, int THREADS> struct PerThread {
    static constexpr int a = kGatherF4 / THREADS / Staged<T>::NB, b = kStreamF4 / THREADS / Staged<T>::NS, c = kLambdaF / THREADS;
    static constexpr int IPT = (a < b ? (a < c ? a : c) : (b < c ? b : c));  // constraints per thread per pass
    static_assert(IPT >= 1, "shared-memory budget too small for this type");
};

constexpr unsigned type_bit(int t) { return 1u << t; }
constexpr unsigned kMaskClothXPBD = type_bit(PBD_DISTANCE_XPBD) | type_bit(PBD_ISOBENDING_XPBD);
constexpr unsigned kMaskLight = type_bit(PBD_DISTANCE) | type_bit(PBD_DISTANCE_XPBD) | type_bit(PBD_DIHEDRAL) | type_bit(PBD_ISOBENDING) |
                                type_bit(PBD_ISOBENDING_XPBD) | type_bit(PBD_VOLUME) | type_bit(PBD_VOLUME_XPBD) | type_bit(PBD_FEMTRIANGLE);
constexpr unsigned kMaskAll = (1u << PBD_BALLJOINT) - 1u;  // every particle constraint type; rigid-body joints are not staged (graph mode only)

// dispatch a statement on the runtime type, restricted to the compiled-in mask (T is a constant inside the statement)
#define PBD_FOR_TYPE(MASK, type, ...)                                                                                    \
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
    case PBD_SHAPEMATCHING:   if constexpr ((MASK) & type_bit(PBD_SHAPEMATCHING))   { constexpr int T = PBD_SHAPEMATCHING;   __VA_ARGS__ } break; \
    default: break;                                                                                                      \
    }

__device__ __forceinline__ Bucket load_bucket(const Bucket *buckets, unsigned bi) {
    Bucket b;
    const int4 raw = __ldg(reinterpret_cast<const int4 *>(buckets) + bi);  // Bucket is 16 bytes
    b.type = raw.x; b.first = (unsigned)raw.y; b.count = (unsigned)raw.z; b.colour = (unsigned)raw.w;
    return b;
}

// Step 1: asynchronously stream indices, rest data and multipliers of this thread's constraints [base + tid + k*stride]
// of bucket b into shared memory.  Slot layout per constraint (16-byte slots, thread-major => conflict-free LDS.128):
//   slot 0: particle indices (2-body: x,y + rest length in .z; 3-body: x,y,z + area in .w)
//   slot 1: float4 geometry g0 (Kp / invRestMat) or the scalar s0 in .x      slot 2: g1      slot 3: (s0, s1)
template <int T, int THREADS>
__device__ __forceinline__ void stream_constraints(const TypeArrays &a, const Bucket &b, unsigned base, unsigned tid, unsigned stride,
                                                   float4 *sS, float *sL, bool wantLambda) {
    constexpr int IPT = PerThread<T, THREADS>::IPT;
    constexpr int NS = Staged<T>::NS;
    const unsigned t = threadIdx.x;
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        const unsigned li = base + tid + (unsigned)k * stride;
        if (li < b.count) {
            const unsigned i = b.first + li;
            float4 *s0 = sS + (size_t)(k * NS) * THREADS + t;
            if (T == PBD_DISTANCE || T == PBD_DISTANCE_XPBD) {
                cp_async8(s0, a.idx2 + i);
                cp_async4(reinterpret_cast<float *>(s0) + 2, a.gs[0] + i);
            } else if (T == PBD_FEMTRIANGLE || T == PBD_STRAINTRIANGLE) {
                unsigned *w = reinterpret_cast<unsigned *>(s0);
                cp_async4(w, a.idx3[0] + i); cp_async4(w + 1, a.idx3[1] + i); cp_async4(w + 2, a.idx3[2] + i);
                if (T == PBD_FEMTRIANGLE) cp_async4(w + 3, a.gs[0] + i);
                cp_async16_cg(s0 + THREADS, a.gv[0] + i);
            } else {
                cp_async16_cg(s0, a.idx4 + i);
                if (T == PBD_DIHEDRAL || T == PBD_VOLUME || T == PBD_VOLUME_XPBD) cp_async4(s0 + THREADS, a.gs[0] + i);
                if (T == PBD_ISOBENDING || T == PBD_ISOBENDING_XPBD) cp_async16_cg(s0 + THREADS, a.gv[0] + i);
                if (T == PBD_FEMTET || T == PBD_FEMTET_XPBD || T == PBD_STRAINTET) {
                    cp_async16_cg(s0 + THREADS, a.gv[0] + i);
                    cp_async16_cg(s0 + 2 * THREADS, a.gv[1] + i);
                    cp_async4(s0 + 3 * THREADS, a.gs[0] + i);
                    if (T != PBD_STRAINTET) cp_async4(reinterpret_cast<float *>(s0 + 3 * THREADS) + 1, a.gs[1] + i);
                }
            }
            if (Staged<T>::XPBD && wantLambda) cp_async4(sL + (size_t)k * THREADS + t, a.lambda + i);
        }
    }
    cp_async_commit();
}

// Steps 2+3 for one pass: gather all particle tuples with cp.async, project each constraint as its tuple lands, scatter.
template <int T, bool CA, int THREADS>
__device__ __forceinline__ void project_pass(float4 *pos, const TypeArrays &a, const Bucket &b, unsigned base, unsigned tid, unsigned stride,
                                             const float4 *sS, const float *sL, float4 *sG, float dt, bool iterZero) {
    constexpr int IPT = PerThread<T, THREADS>::IPT;
    constexpr int NS = Staged<T>::NS;
    constexpr int NB = Staged<T>::NB;
    const unsigned t = threadIdx.x;
    cp_async_wait_all();  // this thread's streamed operands have landed (it reads only its own slots)
    // issue every gather of this pass; one commit group per constraint
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        const unsigned li = base + tid + (unsigned)k * stride;
        if (li < b.count) {
            const uint4 idx = *reinterpret_cast<const uint4 *>(sS + (size_t)(k * NS) * THREADS + t);
            float4 *g = sG + (size_t)(k * NB) * THREADS + t;
            cp_async16_cg(g, pos + idx.x);
            cp_async16_cg(g + THREADS, pos + idx.y);
            if (NB >= 3) cp_async16_cg(g + 2 * THREADS, pos + idx.z);
            if (NB >= 4) cp_async16_cg(g + 3 * THREADS, pos + idx.w);
        }
        cp_async_commit();
    }
    static_for<IPT>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        cp_async_wait_group<IPT - 1 - k>();
        const unsigned li = base + tid + (unsigned)k * stride;
        if (li < b.count) {
            const unsigned i = b.first + li;
            const float4 *s = sS + (size_t)(k * NS) * THREADS + t;
            const float4 *g = sG + (size_t)(k * NB) * THREADS + t;
            const uint4 idx = *reinterpret_cast<const uint4 *>(s);
            float4 p0 = g[0], p1 = g[THREADS], p2, p3;
            if (NB >= 3) p2 = g[2 * THREADS];
            if (NB >= 4) p3 = g[3 * THREADS];
            float lam = 0.0f;
            if (Staged<T>::XPBD && !iterZero) lam = sL[(size_t)k * THREADS + t];

            if (T == PBD_DISTANCE) {
                project_distance(p0, p1, __uint_as_float(idx.z), matv(a, 0, i));
            } else if (T == PBD_DISTANCE_XPBD) {
                project_distance_xpbd(p0, p1, __uint_as_float(idx.z), xpbd_alpha(matv(a, 0, i), dt), lam);
            } else if (T == PBD_FEMTRIANGLE) {
                const FemTriMaterial m = femtri_material(matv(a, 0, i), matv(a, 1, i), matv(a, 2, i), matv(a, 3, i), matv(a, 4, i));
                project_femtriangle(p0, p1, p2, __uint_as_float(idx.w), s[THREADS], m);
            } else if (T == PBD_STRAINTRIANGLE) {
                project_straintriangle(p0, p1, p2, s[THREADS], matv(a, 0, i), matv(a, 1, i), matv(a, 2, i), matv(a, 3, i) != 0.0f, matv(a, 4, i) != 0.0f);
            } else if (T == PBD_DIHEDRAL) {
                project_dihedral(p0, p1, p2, p3, s[THREADS].x, matv(a, 0, i));
            } else if (T == PBD_VOLUME) {
                project_volume<false>(p0, p1, p2, p3, s[THREADS].x, matv(a, 0, i), 0.0f, lam);
            } else if (T == PBD_VOLUME_XPBD) {
                const float k_ = matv(a, 0, i);
                project_volume<true>(p0, p1, p2, p3, s[THREADS].x, k_, xpbd_alpha(k_, dt), lam);
            } else if (T == PBD_ISOBENDING || T == PBD_ISOBENDING_XPBD) {
                constexpr bool X = (T == PBD_ISOBENDING_XPBD);
                const float k_ = matv(a, 0, i);
                const float alpha = X ? xpbd_alpha(k_, dt) : 0.0f;
                if (a.variant == 0) project_isobending_rank1<X>(p0, p1, p2, p3, s[THREADS], k_, alpha, lam);
                else project_isobending_fullq<X>(p0, p1, p2, p3, s[THREADS], __ldg(a.gv[1] + i), __ldg(a.gv[2] + i), __ldg(a.gv[3] + i), k_, alpha, lam);
            } else if (T == PBD_FEMTET || T == PBD_FEMTET_XPBD || T == PBD_STRAINTET) {
                const float4 g0 = s[THREADS], g1 = s[2 * THREADS], sc = s[3 * THREADS];
                M3 inv;
                inv.m[0][0] = g0.x; inv.m[0][1] = g0.y; inv.m[0][2] = g0.z; inv.m[1][0] = g0.w;
                inv.m[1][1] = g1.x; inv.m[1][2] = g1.y; inv.m[2][0] = g1.z; inv.m[2][1] = g1.w;
                inv.m[2][2] = sc.x;
                if (T == PBD_STRAINTET) project_straintet(p0, p1, p2, p3, inv, matv(a, 0, i), matv(a, 1, i), matv(a, 2, i) != 0.0f, matv(a, 3, i) != 0.0f);
                else project_femtet<(T == PBD_FEMTET_XPBD)>(p0, p1, p2, p3, sc.y, inv, matv(a, 0, i), matv(a, 1, i), dt, lam);
            }

            else if (T == PBD_SHAPEMATCHING) {  // only the indices are staged; the 96 B of frozen rest data stream straight from global
                project_shapematching(p0, p1, p2, p3, __ldg(a.gv[0] + i), __ldg(a.gv[1] + i), __ldg(a.gv[2] + i), __ldg(a.gv[3] + i), __ldg(a.gv[4] + i),
                                      __ldg(a.gv[5] + i), matv(a, 0, i));
            }

            if (Staged<T>::XPBD) __stcg(a.lambda + i, lam);
            stp(pos + idx.x, p0); stp(pos + idx.y, p1);
            if (NB >= 3) stp(pos + idx.z, p2);
            if (NB >= 4) stp(pos + idx.w, p3);
        }
    });
}

template <unsigned MASK, bool CA, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_step_persistent(const __grid_constant__ PersistentArgs a) {
    extern __shared__ float4 smem[];
    float4 *sG = smem;
    float4 *sS = smem + kGatherF4;
    float *sL = reinterpret_cast<float *>(smem + kGatherF4 + kStreamF4);

    // Warp-interleaved global thread index: warp w of CTA c owns items [(w*gridDim + c)*32, +32) of every row of `stride`
    // items.  Rows stay coalesced per warp, and a partial last row is spread round-robin over all SMs instead of
    // landing on the first few CTAs (measured with PBD_B200_TRACE: 35-70 % slower CTAs 0..33 with the blocked mapping).
    const unsigned tid = ((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32u + (threadIdx.x & 31u);
    const unsigned stride = gridDim.x * blockDim.x;
    unsigned long long target = a.barrierBase;
    unsigned phase = 0;

    for (unsigned sub = 0; sub < a.subSteps; sub++) {
        // first bucket's operands stream in while the particles are integrated
        Bucket nb;
        nb.type = -1; nb.first = 0u; nb.count = 0u; nb.colour = 0u;
        if (a.nBuckets) {
            nb = load_bucket(a.buckets, 0);
            PBD_FOR_TYPE(MASK, nb.type, stream_constraints<T, THREADS>(a.types[T], nb, 0u, tid, stride, sS, sL, false);)
        }

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
        grid_barrier(a.barrier, target, a.trace, a.tracePhases, phase++);  // integration done everywhere before the first colour reads positions

        // ---- coloured Gauss-Seidel sweeps -------------------------------------------------------------------
        for (unsigned it = 0; it < a.maxIter; it++) {
            const bool iterZero = (it == 0);
            for (unsigned bi = 0; bi < a.nBuckets; bi++) {
                const Bucket b = nb;
                const bool lastOfSweep = (bi + 1 == a.nBuckets);
                const bool more = !lastOfSweep || (it + 1 < a.maxIter);
                if (more) nb = load_bucket(a.buckets, lastOfSweep ? 0u : bi + 1);

                PBD_FOR_TYPE(MASK, b.type,
                    const TypeArrays &ta = a.types[T];
                    const unsigned perPass = stride * (unsigned)PerThread<T, THREADS>::IPT;
                    for (unsigned base = 0; base < b.count; base += perPass) {
                        if (base != 0) stream_constraints<T, THREADS>(ta, b, base, tid, stride, sS, sL, !iterZero);  // pass 0 was streamed ahead
                        project_pass<T, CA, THREADS>(a.pos, ta, b, base, tid, stride, sS, sL, sG, a.h, iterZero);
                    })

                if (more) {
                    // stream the next bucket's operands, then synchronise if it starts a new colour phase
                    const bool nextIterZero = lastOfSweep ? false : iterZero;
                    PBD_FOR_TYPE(MASK, nb.type, stream_constraints<T, THREADS>(a.types[T], nb, 0u, tid, stride, sS, sL, !nextIterZero);)
                    if (nb.colour != b.colour || lastOfSweep) grid_barrier(a.barrier, target, a.trace, a.tracePhases, phase++);
                }
            }
        }
        if (a.nBuckets) grid_barrier(a.barrier, target, a.trace, a.tracePhases, phase++);  // all projections done before velocities are derived

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
