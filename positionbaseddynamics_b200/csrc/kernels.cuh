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
//   k_step_resident (resident.cuh): the whole step in one launch, positions resident in the shared memory of thread-block clusters
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
__device__ __forceinline__ float xpbd_alpha(float k, float dt) { return (k != 0.0f) ? frcp(k * dt * dt) : 0.0f; }

// What a constraint streams from HBM before it touches any particle: indices + per-constraint constants.  These never
// change during a step, so they can be fetched ahead of the dependency on the previous colour phase (before the PDL
// wait in k_project, before the grid barrier in the persistent kernel).
struct Streamed {
    uint4 b;        // particle indices (x,y[,z[,w]])
    float4 g0, g1;  // float4 geometry (Kp / invRestMat rows)
    float s0, s1;   // scalar geometry (rest length / angle / volume / area / invRestMat(2,2))
};

template <int T>
__device__ __forceinline__ Streamed load_streamed(const TypeArrays &a, unsigned i) {
    Streamed s;
    s.g0 = make_float4(0.f, 0.f, 0.f, 0.f); s.g1 = s.g0; s.s0 = 0.0f; s.s1 = 0.0f;
    if (T == PBD_BALLJOINT || T == PBD_RB_PARTICLE_BALLJOINT) {
        const uint2 b = __ldg(a.idx2 + i);
        s.b = make_uint4(b.x, b.y, 0u, 0u);
        s.g0 = __ldg(a.gv[0] + i);
        if (T == PBD_BALLJOINT) s.g1 = __ldg(a.gv[1] + i);
    } else if (T == PBD_DISTANCE || T == PBD_DISTANCE_XPBD) {
        const uint2 b = __ldg(a.idx2 + i);
        s.b = make_uint4(b.x, b.y, 0u, 0u);
        s.s0 = __ldg(a.gs[0] + i);
    } else if (T == PBD_FEMTRIANGLE || T == PBD_STRAINTRIANGLE) {
        s.b = make_uint4(__ldg(a.idx3[0] + i), __ldg(a.idx3[1] + i), __ldg(a.idx3[2] + i), 0u);
        s.g0 = __ldg(a.gv[0] + i);
        if (T == PBD_FEMTRIANGLE) s.s0 = __ldg(a.gs[0] + i);
    } else {
        s.b = __ldg(a.idx4 + i);
        if (T == PBD_DIHEDRAL || T == PBD_VOLUME || T == PBD_VOLUME_XPBD) s.s0 = __ldg(a.gs[0] + i);
        if (T == PBD_ISOBENDING || T == PBD_ISOBENDING_XPBD) s.g0 = __ldg(a.gv[0] + i);
        if (T == PBD_FEMTET || T == PBD_FEMTET_XPBD || T == PBD_STRAINTET) { s.g0 = __ldg(a.gv[0] + i); s.g1 = __ldg(a.gv[1] + i); s.s0 = __ldg(a.gs[0] + i); }
        if (T == PBD_FEMTET || T == PBD_FEMTET_XPBD) s.s1 = __ldg(a.gs[1] + i);
    }
    return s;
}

// Particle accessors: where a constraint's particle index points to.
//   GlobalAcc : the float4 array in global memory (L2-resident), index = device slot
//   ClusterAcc (resident.cuh): bit 31 set -> slot in the shared-memory tile of a CTA of the executing cluster (DSMEM), else global
// handle(idx) resolves the address once; ld / st take the handle (a particle is read and written through the same one)
template <bool CA> struct GlobalAcc {
    typedef float4 *Handle;
    float4 *pos;
    __device__ __forceinline__ Handle handle(unsigned idx) const { return pos + idx; }
    __device__ __forceinline__ float4 ld(Handle h) const { return ldp<CA>(h); }
    __device__ __forceinline__ void st(Handle h, const float4 &v) const { stp(h, v); }
};

// Gather -> project -> scatter for constraint i (index into the type's arrays) whose streamed part is already here.
// Joints couple rigid bodies (their own small state arrays, L2 only) with each other or with a particle.
template <int T, class Acc>
__device__ __forceinline__ void project_joint(const Acc &acc, const TypeArrays &a, const Streamed &s) {
    float4 X0 = __ldcg(a.rbX + s.b.x), Q0 = __ldcg(a.rbQ + s.b.x);
    const float4 I0 = __ldg(a.rbIinv + s.b.x);
    if (T == PBD_BALLJOINT) {
        float4 X1 = __ldcg(a.rbX + s.b.y), Q1 = __ldcg(a.rbQ + s.b.y);
        const float4 I1 = __ldg(a.rbIinv + s.b.y);
        project_balljoint(X0, Q0, mk(I0.x, I0.y, I0.z), X1, Q1, mk(I1.x, I1.y, I1.z), mk(s.g0.x, s.g0.y, s.g0.z), mk(s.g1.x, s.g1.y, s.g1.z));
        if (X1.w != 0.0f) { __stcg(a.rbX + s.b.y, X1); __stcg(a.rbQ + s.b.y, Q1); }
    } else {
        const typename Acc::Handle hp = acc.handle(s.b.y);
        float4 p = acc.ld(hp);
        project_rb_particle_balljoint(X0, Q0, mk(I0.x, I0.y, I0.z), p, mk(s.g0.x, s.g0.y, s.g0.z));
        acc.st(hp, p);
    }
    if (X0.w != 0.0f) { __stcg(a.rbX + s.b.x, X0); __stcg(a.rbQ + s.b.x, Q0); }
}

// VAR: compile-time layout variant of the type (IsometricBending: 0 = rank-1 Kp, 1 = full Q), -1 = read a.variant at run time
template <int T, class Acc, int VAR = -1>
__device__ __forceinline__ void project_streamed_acc(const Acc &acc, const TypeArrays &a, unsigned i, const Streamed &s, float dt,
                                                     bool iterZero, bool lambdaGiven = false, float lambdaValue = 0.0f) {
    if (T == PBD_BALLJOINT || T == PBD_RB_PARTICLE_BALLJOINT) { project_joint<T>(acc, a, s); return; }
    constexpr bool XPBD = (T == PBD_DISTANCE_XPBD || T == PBD_VOLUME_XPBD || T == PBD_ISOBENDING_XPBD || T == PBD_FEMTET_XPBD);
    constexpr int NB = (T == PBD_DISTANCE || T == PBD_DISTANCE_XPBD) ? 2 : ((T == PBD_FEMTRIANGLE || T == PBD_STRAINTRIANGLE) ? 3 : 4);
    const typename Acc::Handle h0 = acc.handle(s.b.x), h1 = acc.handle(s.b.y), h2 = acc.handle(NB >= 3 ? s.b.z : s.b.x), h3 = acc.handle(NB >= 4 ? s.b.w : s.b.x);
    float4 p0 = acc.ld(h0), p1 = acc.ld(h1), p2, p3;
    if (NB >= 3) p2 = acc.ld(h2);
    if (NB >= 4) p3 = acc.ld(h3);
    float lam = 0.0f;
    if (XPBD && !iterZero) lam = lambdaGiven ? lambdaValue : __ldcg(a.lambda + i);  // m_lambda; zero at the first sweep of a substep (Constraints.cpp:1241-1242)

    if (T == PBD_DISTANCE) {
        project_distance(p0, p1, s.s0, matv(a, 0, i));
    } else if (T == PBD_DISTANCE_XPBD) {
        project_distance_xpbd(p0, p1, s.s0, xpbd_alpha(matv(a, 0, i), dt), lam);
    } else if (T == PBD_FEMTRIANGLE) {
        const FemTriMaterial m = femtri_material(matv(a, 0, i), matv(a, 1, i), matv(a, 2, i), matv(a, 3, i), matv(a, 4, i));
        project_femtriangle(p0, p1, p2, s.s0, s.g0, m);
    } else if (T == PBD_STRAINTRIANGLE) {
        project_straintriangle(p0, p1, p2, s.g0, matv(a, 0, i), matv(a, 1, i), matv(a, 2, i), matv(a, 3, i) != 0.0f, matv(a, 4, i) != 0.0f);
    } else if (T == PBD_DIHEDRAL) {
        project_dihedral(p0, p1, p2, p3, s.s0, matv(a, 0, i));
    } else if (T == PBD_VOLUME) {
        project_volume<false>(p0, p1, p2, p3, s.s0, matv(a, 0, i), 0.0f, lam);
    } else if (T == PBD_VOLUME_XPBD) {
        const float k = matv(a, 0, i);
        project_volume<true>(p0, p1, p2, p3, s.s0, k, xpbd_alpha(k, dt), lam);
    } else if (T == PBD_ISOBENDING || T == PBD_ISOBENDING_XPBD) {
        constexpr bool X = (T == PBD_ISOBENDING_XPBD);
        const float k = matv(a, 0, i);
        const float alpha = X ? xpbd_alpha(k, dt) : 0.0f;
        if (VAR == 0 || (VAR < 0 && a.variant == 0)) project_isobending_rank1<X>(p0, p1, p2, p3, s.g0, k, alpha, lam);
        else project_isobending_fullq<X>(p0, p1, p2, p3, s.g0, __ldg(a.gv[1] + i), __ldg(a.gv[2] + i), __ldg(a.gv[3] + i), k, alpha, lam);
    } else if (T == PBD_FEMTET || T == PBD_FEMTET_XPBD || T == PBD_STRAINTET) {
        M3 inv;
        inv.m[0][0] = s.g0.x; inv.m[0][1] = s.g0.y; inv.m[0][2] = s.g0.z; inv.m[1][0] = s.g0.w;
        inv.m[1][1] = s.g1.x; inv.m[1][2] = s.g1.y; inv.m[2][0] = s.g1.z; inv.m[2][1] = s.g1.w;
        inv.m[2][2] = s.s0;
        if (T == PBD_STRAINTET) project_straintet(p0, p1, p2, p3, inv, matv(a, 0, i), matv(a, 1, i), matv(a, 2, i) != 0.0f, matv(a, 3, i) != 0.0f);
        else project_femtet<(T == PBD_FEMTET_XPBD)>(p0, p1, p2, p3, s.s1, inv, matv(a, 0, i), matv(a, 1, i), dt, lam);
    } else if (T == PBD_SHAPEMATCHING) {
        project_shapematching(p0, p1, p2, p3, __ldg(a.gv[0] + i), __ldg(a.gv[1] + i), __ldg(a.gv[2] + i), __ldg(a.gv[3] + i), __ldg(a.gv[4] + i),
                              __ldg(a.gv[5] + i), matv(a, 0, i));
    }

    if (XPBD) __stcg(a.lambda + i, lam);
    acc.st(h0, p0); acc.st(h1, p1);
    if (NB >= 3) acc.st(h2, p2);
    if (NB >= 4) acc.st(h3, p3);
}

template <int T, bool CA, int VAR = -1>
__device__ __forceinline__ void project_streamed(float4 *pos, const TypeArrays &a, unsigned i, const Streamed &s, float dt, bool iterZero) {
    project_streamed_acc<T, GlobalAcc<CA>, VAR>(GlobalAcc<CA>{pos}, a, i, s, dt, iterZero);
}

#ifndef PBD_PROJECT_THREADS
#define PBD_PROJECT_THREADS 256
#endif
constexpr int kProjectThreads = PBD_PROJECT_THREADS;

template <int T, bool CA, int VAR>
__global__ void __launch_bounds__(kProjectThreads) k_project(float4 *pos, TypeArrays a, unsigned first,
                                                             unsigned count, float dt, int iterZero) {
    pdl_launch_dependents();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;  // an exited thread counts as having passed the dependency
    const Streamed s = load_streamed<T>(a, first + i);
    pdl_wait_after(s.b.x ^ s.b.y, s.g0.x + s.g1.x + s.s0 + s.s1);
    project_streamed<T, CA, VAR>(pos, a, first + i, s, dt, iterZero != 0);
}

// ---- Jacobi comparison path (north_star: "a Jacobi path uses atomicAdd for comparison") -------------------------------------------
// One launch per constraint TYPE and sweep over ALL its constraints, colours ignored: every projection reads the positions of the
// sweep's start and adds its correction to a scratch float4 per particle with one vector atomicAdd (x, y, z, 1); k_jacobi_apply then
// moves every touched particle by the average of its corrections.  A different algorithm from the reference's Gauss-Seidel sweep
// (slower convergence, no ordering), so it is not parity-gated; it exists to price the colouring against atomics (profiles/README.md).
struct JacobiAcc {
    typedef unsigned Handle;
    const float4 *pos; float4 *delta;
    __device__ __forceinline__ Handle handle(unsigned idx) const { return idx; }
    __device__ __forceinline__ float4 ld(Handle h) const { return __ldg(pos + h); }
    __device__ __forceinline__ void st(Handle h, const float4 &v) const {
        if (v.w == 0.0f) return;
        const float4 o = __ldg(pos + h);
        atomicAdd(delta + h, make_float4(v.x - o.x, v.y - o.y, v.z - o.z, 1.0f));  // RED.E.ADD.F32x4 (sm_90+)
    }
};
template <int T, int VAR>
__global__ void __launch_bounds__(kProjectThreads) k_project_jacobi(const float4 *pos, float4 *delta, TypeArrays a, unsigned count, float dt, int iterZero) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const Streamed s = load_streamed<T>(a, i);
    project_streamed_acc<T, JacobiAcc, VAR>(JacobiAcc{pos, delta}, a, i, s, dt, iterZero != 0);
}
__global__ void __launch_bounds__(256) k_jacobi_apply(float4 *__restrict__ pos, float4 *__restrict__ delta, unsigned n) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 d = delta[i];
    if (d.w == 0.0f) return;
    const float inv = 1.0f / d.w;
    float4 x = pos[i];
    x.x = fmaf(d.x, inv, x.x); x.y = fmaf(d.y, inv, x.y); x.z = fmaf(d.z, inv, x.z);
    pos[i] = x;
    delta[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Several (colour,type) buckets of ONE colour in a single launch: buckets of a colour touch disjoint particles, so they
// need no ordering among themselves (the reference runs a whole colour group under one `omp parallel for`,
// TimeStepController.cpp:275-285).  A CTA finds its segment from blockIdx and dispatches on the segment's type.
// Used when a colour holds more than one constraint type (e.g. FEMTet + Volume in cfg3: 112 buckets but 74 colours).
constexpr int kMultiSegments = 4;
struct MultiArgs {
    int nSeg;
    int type[kMultiSegments];
    unsigned first[kMultiSegments], count[kMultiSegments], blockStart[kMultiSegments + 1];
    TypeArrays arrays[kMultiSegments];
};

template <bool CA>
__global__ void __launch_bounds__(kProjectThreads) k_project_multi(float4 *pos, const __grid_constant__ MultiArgs m, float dt, int iterZero) {
    pdl_launch_dependents();
    int s = 0;
#pragma unroll
    for (int k = 1; k < kMultiSegments; k++)
        if (k < m.nSeg && blockIdx.x >= m.blockStart[k]) s = k;
    const unsigned i = (blockIdx.x - m.blockStart[s]) * blockDim.x + threadIdx.x;
    if (i >= m.count[s]) return;
    const TypeArrays &a = m.arrays[s];
    const unsigned ci = m.first[s] + i;
#define PM(T) case T: { const Streamed st = load_streamed<T>(a, ci); pdl_wait_after(st.b.x ^ st.b.y, st.g0.x + st.g1.x + st.s0 + st.s1); \
                        project_streamed<T, CA>(pos, a, ci, st, dt, iterZero != 0); } break;
    switch (m.type[s]) {
        PM(PBD_DISTANCE) PM(PBD_DISTANCE_XPBD) PM(PBD_DIHEDRAL) PM(PBD_ISOBENDING) PM(PBD_ISOBENDING_XPBD) PM(PBD_FEMTRIANGLE)
        PM(PBD_STRAINTRIANGLE) PM(PBD_VOLUME) PM(PBD_VOLUME_XPBD) PM(PBD_FEMTET) PM(PBD_FEMTET_XPBD) PM(PBD_STRAINTET) PM(PBD_SHAPEMATCHING) PM(PBD_BALLJOINT) PM(PBD_RB_PARTICLE_BALLJOINT)
    default: break;
    }
#undef PM
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

// Rigid-body prologue / epilogue (a handful of bodies: one thread each).
//   integrate: TimeStepController.cpp:97-107 + TimeIntegration::semiImplicitEuler / semiImplicitEulerRotation (TimeIntegration.cpp:7-39)
//   velocity : TimeStepController.cpp:139-152 + velocityUpdate*/angularVelocityUpdate* (TimeIntegration.cpp:42-95; the angular update is
//              first order in both modes, as in the reference)
struct RbState { float4 *X, *Q, *V, *W, *oldX, *lastX, *oldQ, *lastQ; const float4 *I, *Iinv; unsigned n; };

__device__ __forceinline__ void rb_integrate_body(const RbState &r, unsigned i, float h, float gx, float gy, float gz) {
    float4 X = __ldcg(r.X + i), Q = __ldcg(r.Q + i), V = r.V[i], W = r.W[i];
    r.lastX[i] = r.oldX[i]; r.oldX[i] = X;
    r.lastQ[i] = r.oldQ[i]; r.oldQ[i] = Q;
    if (V.w != 0.0f) {  // V.w = mass
        V.x = fmaf(gx, h, V.x); V.y = fmaf(gy, h, V.y); V.z = fmaf(gz, h, V.z);
        X.x = fmaf(V.x, h, X.x); X.y = fmaf(V.y, h, X.y); X.z = fmaf(V.z, h, X.z);
        const M3 R = qmatrix(Q);
        const float4 I = r.I[i], Ii = r.Iinv[i];
        const M3 Iw = world_tensor(R, mk(I.x, I.y, I.z)), Jw = world_tensor(R, mk(Ii.x, Ii.y, Ii.z));
        V3 om = mk(W.x, W.y, W.z);
        const V3 t = -cross(om, mvec(Iw, om));  // torque = 0
        om = om + mvec(Jw, t) * h;
        const float4 dq = qmul(make_float4(om.x, om.y, om.z, 0.0f), Q);
        const float hh = h * 0.5f;
        Q = qnormalize(make_float4(Q.x + hh * dq.x, Q.y + hh * dq.y, Q.z + hh * dq.z, Q.w + hh * dq.w));
        W.x = om.x; W.y = om.y; W.z = om.z;
        __stcg(r.X + i, X); __stcg(r.Q + i, Q); r.V[i] = V; r.W[i] = W;
    }
}
__device__ __forceinline__ void rb_velocity_body(const RbState &r, unsigned i, float invH, float twoInvH, int secondOrder) {
    float4 V = r.V[i];
    if (V.w == 0.0f) return;
    const float4 X = __ldcg(r.X + i), Q = __ldcg(r.Q + i), o = r.oldX[i], oq = r.oldQ[i];
    if (!secondOrder) { V.x = invH * (X.x - o.x); V.y = invH * (X.y - o.y); V.z = invH * (X.z - o.z); }
    else {
        const float4 l = r.lastX[i];
        V.x = invH * (1.5f * X.x - 2.0f * o.x + 0.5f * l.x); V.y = invH * (1.5f * X.y - 2.0f * o.y + 0.5f * l.y); V.z = invH * (1.5f * X.z - 2.0f * o.z + 0.5f * l.z);
    }
    const float4 rel = qmul(Q, make_float4(-oq.x, -oq.y, -oq.z, oq.w));
    float4 W = r.W[i];
    W.x = rel.x * twoInvH; W.y = rel.y * twoInvH; W.z = rel.z * twoInvH;
    r.V[i] = V; r.W[i] = W;
}

__global__ void k_rb_integrate(RbState r, float h, float gx, float gy, float gz) {
    pdl_launch_dependents();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r.n) return;
    pdl_wait();
    rb_integrate_body(r, i, h, gx, gy, gz);
}

__global__ void k_rb_velocity(RbState r, float invH, float twoInvH, int secondOrder) {
    pdl_launch_dependents();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r.n) return;
    pdl_wait();
    rb_velocity_body(r, i, invH, twoInvH, secondOrder);
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

// host AoS-3 <-> device float4 conversion (std::vector<Vector3r> layout on the host side); particle i of the host lives
// in device slot slot[i] (de-interleaved formula of device_image.h, or the tile-major permutation of the tiled mode)
__global__ void k_relayout(const float4 *__restrict__ src, float4 *__restrict__ dst, unsigned n, const unsigned *__restrict__ oldSlot,
                           const unsigned *__restrict__ newSlot) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[newSlot[i]] = src[oldSlot[i]];
}
__global__ void k_pack3(const float *__restrict__ src, float4 *__restrict__ dst, unsigned n, int keepW, const unsigned *__restrict__ slot) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned s = slot[i];
    float4 d = keepW ? dst[s] : make_float4(0.f, 0.f, 0.f, 0.f);
    d.x = src[3 * i]; d.y = src[3 * i + 1]; d.z = src[3 * i + 2];
    dst[s] = d;
}
__global__ void k_unpack3(const float4 *__restrict__ src, float *__restrict__ dst, unsigned n, const unsigned *__restrict__ slot) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s = src[slot[i]];
    dst[3 * i] = s.x; dst[3 * i + 1] = s.y; dst[3 * i + 2] = s.z;
}
__global__ void k_set_w(float4 *__restrict__ pos, float4 *__restrict__ vel, const float *__restrict__ mass, unsigned n, const unsigned *__restrict__ slot) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float m = mass[i];
    const unsigned s = slot[i];
    pos[s].w = (m != 0.0f) ? 1.0f / m : 0.0f;  // ParticleData::setMass (ParticleData.h:239-246)
    vel[s].w = m;
}

}  // namespace pbdk
