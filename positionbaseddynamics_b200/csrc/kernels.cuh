// positionbaseddynamics_b200/csrc/kernels.cuh
//
// sm_100a kernels of the PBD/XPBD step.  One thread projects one constraint: it loads the constraint's particle
// indices (coalesced), gathers the particle float4s straight from L2 (ld.global.cg -- within one colour every
// particle is touched by at most one constraint, so there is no L1 reuse to win and, in the persistent kernel,
// L1 must not serve stale lines across colour phases), runs the projection in registers and scatters the float4s
// back (st.global.cg).  Colours guarantee disjoint particle sets, so no atomics.
//
//   k_integrate   : TimeStepController.cpp:112-118 + TimeIntegration::semiImplicitEuler (TimeIntegration.cpp:7-19);
//                   gravity is a uniform (TimeStep::clearAccelerations sets a = g for every dynamic particle, TimeStep.cpp:28-62)
//   k_project<T>  : one (colour,type) bucket of TimeStepController::positionConstraintProjection (TimeStepController.cpp:270-286)
//   k_velocity    : TimeStepController.cpp:155-162 + TimeIntegration::velocityUpdateFirstOrder/SecondOrder (TimeIntegration.cpp:42-51, 69-79)
//   k_step_persistent : the whole step in one cooperative launch, grid barrier between colour phases
#pragma once
#include "device_image.h"
#include "solvers.cuh"

namespace pbdk {

// Programmatic dependent launch (PDL): every kernel of the step is launched with the programmatic-stream-serialization
// attribute, announces "dependents may launch" immediately and waits for its predecessor only right before it touches
// particle data.  The next bucket's CTAs therefore get scheduled while the current bucket drains, and their coalesced
// streaming loads of indices / rest data (the HBM-latency part) overlap the predecessor's tail.  Without the launch
// attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// wait that is data-dependent on the streamed operands, so that ptxas cannot sink their loads below the wait
// (ptxas hoists an operand-less ACQBULK above independent loads; predicating the wait on the loaded data -- both
// predicate senses wait, so it always executes -- pins it behind them)
__device__ __forceinline__ void pdl_wait_after(unsigned a, float b) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, %1;\n\t@p griddepcontrol.wait;\n\t@!p griddepcontrol.wait;\n\t}" :: "r"(a), "r"(__float_as_uint(b)) : "memory");
}

// Particle gather flavour (template parameter CA of the kernels):
//   CA = false: ld.global.cg -- L2 only.
//   CA = true : ld.global.ca -- through L1, so the 2-4 gathers of one constraint and of neighbouring threads that fall
//               into the same 32 B sector / 128 B line are served once.  Safe because L1 is invalidated at every kernel
//               boundary and, in the persistent kernel, by the gpu-scope fence on the acquire side of the grid barrier.
template <bool CA> __device__ __forceinline__ float4 ldp(const float4 *p) { return CA ? *p : __ldcg(p); }
__device__ __forceinline__ void stp(float4 *p, const float4 &v) { if (v.w != 0.0f) __stcg(p, v); }  // static particles never move

__device__ __forceinline__ float matv(const TypeArrays &a, int slot, unsigned i) {
    return a.mat[slot] ? __ldg(a.mat[slot] + i) : a.matU[slot];
}
__device__ __forceinline__ float xpbd_alpha(float k, float dt) { return (k != 0.0f) ? 1.0f / (k * dt * dt) : 0.0f; }

// Gather -> project -> scatter for constraint i (index into the type's arrays).
template <int T, bool CA, bool WAIT = false>
__device__ __forceinline__ void process_constraint(float4 *pos, const TypeArrays &a, unsigned i, float dt,
                                                   bool iterZero) {
    if (T == PBD_DISTANCE || T == PBD_DISTANCE_XPBD) {
        const uint2 b = __ldg(a.idx2 + i);
        const float rest = __ldg(a.gs[0] + i);
        if (WAIT) pdl_wait_after(b.x, rest);
        float4 p0 = ldp<CA>(pos + b.x), p1 = ldp<CA>(pos + b.y);
        const float k = matv(a, 0, i);
        if (T == PBD_DISTANCE) {
            project_distance(p0, p1, rest, k);
        } else {
            float lam = iterZero ? 0.0f : __ldcg(a.lambda + i);
            project_distance_xpbd(p0, p1, rest, xpbd_alpha(k, dt), lam);
            __stcg(a.lambda + i, lam);
        }
        stp(pos + b.x, p0); stp(pos + b.y, p1);
    } else if (T == PBD_FEMTRIANGLE || T == PBD_STRAINTRIANGLE) {
        const unsigned b0 = __ldg(a.idx3[0] + i), b1 = __ldg(a.idx3[1] + i), b2 = __ldg(a.idx3[2] + i);
        const float4 inv = __ldg(a.gv[0] + i);
        if (WAIT) pdl_wait_after(b0 + b1 + b2, inv.x);
        float4 p0 = ldp<CA>(pos + b0), p1 = ldp<CA>(pos + b1), p2 = ldp<CA>(pos + b2);
        if (T == PBD_FEMTRIANGLE) {
            const FemTriMaterial m = femtri_material(matv(a, 0, i), matv(a, 1, i), matv(a, 2, i), matv(a, 3, i), matv(a, 4, i));
            project_femtriangle(p0, p1, p2, __ldg(a.gs[0] + i), inv, m);
        } else {
            project_straintriangle(p0, p1, p2, inv, matv(a, 0, i), matv(a, 1, i), matv(a, 2, i), matv(a, 3, i) != 0.0f, matv(a, 4, i) != 0.0f);
        }
        stp(pos + b0, p0); stp(pos + b1, p1); stp(pos + b2, p2);
    } else {
        const uint4 b = __ldg(a.idx4 + i);
        // stream the per-constraint constants before waiting on the predecessor kernel (they never change during a step)
        float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
        float s0 = 0.0f, s1 = 0.0f;
        if (T == PBD_DIHEDRAL || T == PBD_VOLUME || T == PBD_VOLUME_XPBD) s0 = __ldg(a.gs[0] + i);
        if (T == PBD_ISOBENDING || T == PBD_ISOBENDING_XPBD) g0 = __ldg(a.gv[0] + i);
        if (T == PBD_FEMTET || T == PBD_FEMTET_XPBD || T == PBD_STRAINTET) { g0 = __ldg(a.gv[0] + i); g1 = __ldg(a.gv[1] + i); s0 = __ldg(a.gs[0] + i); }
        if (T == PBD_FEMTET || T == PBD_FEMTET_XPBD) s1 = __ldg(a.gs[1] + i);
        if (WAIT) pdl_wait_after(b.x, g0.x + g1.x + s0 + s1);
        float4 p0 = ldp<CA>(pos + b.x), p1 = ldp<CA>(pos + b.y), p2 = ldp<CA>(pos + b.z), p3 = ldp<CA>(pos + b.w);
        if (T == PBD_DIHEDRAL) {
            project_dihedral(p0, p1, p2, p3, s0, matv(a, 0, i));
        } else if (T == PBD_VOLUME) {
            float dummy = 0.0f;
            project_volume<false>(p0, p1, p2, p3, s0, matv(a, 0, i), 0.0f, dummy);
        } else if (T == PBD_VOLUME_XPBD) {
            float lam = iterZero ? 0.0f : __ldcg(a.lambda + i);
            const float k = matv(a, 0, i);
            project_volume<true>(p0, p1, p2, p3, s0, k, xpbd_alpha(k, dt), lam);
            __stcg(a.lambda + i, lam);
        } else if (T == PBD_ISOBENDING || T == PBD_ISOBENDING_XPBD) {
            constexpr bool X = (T == PBD_ISOBENDING_XPBD);
            const float k = matv(a, 0, i);
            float lam = 0.0f;
            if (X && !iterZero) lam = __ldcg(a.lambda + i);
            const float alpha = X ? xpbd_alpha(k, dt) : 0.0f;
            if (a.variant == 0) {
                project_isobending_rank1<X>(p0, p1, p2, p3, g0, k, alpha, lam);
            } else {
                project_isobending_fullq<X>(p0, p1, p2, p3, g0, __ldg(a.gv[1] + i), __ldg(a.gv[2] + i), __ldg(a.gv[3] + i), k, alpha, lam);
            }
            if (X) __stcg(a.lambda + i, lam);
        } else if (T == PBD_FEMTET || T == PBD_FEMTET_XPBD) {
            constexpr bool X = (T == PBD_FEMTET_XPBD);
            const float4 m0 = g0, m1 = g1;
            M3 inv;
            inv.m[0][0] = m0.x; inv.m[0][1] = m0.y; inv.m[0][2] = m0.z; inv.m[1][0] = m0.w;
            inv.m[1][1] = m1.x; inv.m[1][2] = m1.y; inv.m[2][0] = m1.z; inv.m[2][1] = m1.w;
            inv.m[2][2] = s0;
            const float vol = s1;
            float lam = 0.0f;
            if (X && !iterZero) lam = __ldcg(a.lambda + i);
            project_femtet<X>(p0, p1, p2, p3, vol, inv, matv(a, 0, i), matv(a, 1, i), dt, lam);
            if (X) __stcg(a.lambda + i, lam);
        } else if (T == PBD_STRAINTET) {
            const float4 m0 = g0, m1 = g1;
            M3 inv;
            inv.m[0][0] = m0.x; inv.m[0][1] = m0.y; inv.m[0][2] = m0.z; inv.m[1][0] = m0.w;
            inv.m[1][1] = m1.x; inv.m[1][2] = m1.y; inv.m[2][0] = m1.z; inv.m[2][1] = m1.w;
            inv.m[2][2] = s0;
            project_straintet(p0, p1, p2, p3, inv, matv(a, 0, i), matv(a, 1, i), matv(a, 2, i) != 0.0f, matv(a, 3, i) != 0.0f);
        }
        stp(pos + b.x, p0); stp(pos + b.y, p1); stp(pos + b.z, p2); stp(pos + b.w, p3);
    }
}

constexpr int kProjectThreads = 256;

template <int T, bool CA>
__global__ void __launch_bounds__(kProjectThreads) k_project(float4 *pos, TypeArrays a, unsigned first,
                                                             unsigned count, float dt, int iterZero) {
    pdl_launch_dependents();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) process_constraint<T, CA, true>(pos, a, first + i, dt, iterZero != 0);
    // threads without a constraint simply exit: an exited thread counts as having passed the dependency
}

// lastX = oldX; oldX = x; if (mass != 0) { v += g h; x += v h }
__global__ void __launch_bounds__(256) k_integrate(float4 *__restrict__ pos, float4 *__restrict__ vel,
                                                   float4 *__restrict__ oldp, float4 *__restrict__ lastp, unsigned n,
                                                   float h, float gx, float gy, float gz, int trackLast) {
    pdl_launch_dependents();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pdl_wait();
    float4 x = __ldcg(pos + i);
    if (trackLast) __stcs(lastp + i, __ldcs(oldp + i));
    __stcg(oldp + i, x);
    float4 v = __ldcs(vel + i);
    if (v.w != 0.0f) {  // v.w carries the mass (TimeIntegration.cpp:14 tests mass != 0)
        v.x = fmaf(gx, h, v.x); v.y = fmaf(gy, h, v.y); v.z = fmaf(gz, h, v.z);
        x.x = fmaf(v.x, h, x.x); x.y = fmaf(v.y, h, x.y); x.z = fmaf(v.z, h, x.z);
        __stcs(vel + i, v);
        __stcg(pos + i, x);
    }
}

// v = (1/h)(x - oldX)   or   (1/h)(1.5 x - 2 oldX + 0.5 lastX)
__global__ void __launch_bounds__(256) k_velocity(const float4 *__restrict__ pos, float4 *__restrict__ vel,
                                                  const float4 *__restrict__ oldp, const float4 *__restrict__ lastp,
                                                  unsigned n, float invH, int secondOrder) {
    pdl_launch_dependents();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pdl_wait();
    const float4 x = __ldcg(pos + i);
    if (x.w == 0.0f) return;  // invMass == 0  <=>  mass == 0 (ParticleData::setMass keeps them consistent)
    const float4 o = __ldcg(oldp + i);
    float vx, vy, vz;
    if (!secondOrder) {
        vx = invH * (x.x - o.x); vy = invH * (x.y - o.y); vz = invH * (x.z - o.z);
    } else {
        const float4 l = __ldcs(lastp + i);
        vx = invH * (1.5f * x.x - 2.0f * o.x + 0.5f * l.x);
        vy = invH * (1.5f * x.y - 2.0f * o.y + 0.5f * l.y);
        vz = invH * (1.5f * x.z - 2.0f * o.z + 0.5f * l.z);
    }
    // keep the mass in .w: re-read only that lane's .w would cost a full sector anyway, so read-modify-write the float4
    float4 v = __ldcs(vel + i);
    v.x = vx; v.y = vy; v.z = vz;
    __stcs(vel + i, v);
}

// host AoS-3 <-> device float4 conversion (std::vector<Vector3r> layout on the host side)
__global__ void k_pack3(const float *__restrict__ src, float4 *__restrict__ dst, unsigned n, int keepW) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 d = keepW ? dst[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    d.x = src[3 * i]; d.y = src[3 * i + 1]; d.z = src[3 * i + 2];
    dst[i] = d;
}
__global__ void k_unpack3(const float4 *__restrict__ src, float *__restrict__ dst, unsigned n) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s = src[i];
    dst[3 * i] = s.x; dst[3 * i + 1] = s.y; dst[3 * i + 2] = s.z;
}
__global__ void k_set_w(float4 *__restrict__ pos, float4 *__restrict__ vel, const float *__restrict__ mass, unsigned n) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float m = mass[i];
    pos[i].w = (m != 0.0f) ? 1.0f / m : 0.0f;  // ParticleData::setMass (ParticleData.h:239-246)
    vel[i].w = m;
}

}  // namespace pbdk
