// positionbaseddynamics_b200/csrc/solvers.cuh
//
// Device-side constraint projections for the B200 engine: one `__device__` function per constraint type,
// each operating on a particle tuple already gathered into registers (float4 = x,y,z,invMass) and returning
// the position corrections.  These are the sm_100a counterparts of the reference's stateless solver
// functions; each one cites the function whose result it reproduces:
//   PositionBasedDynamics/PositionBasedDynamics.cpp  (solve_Distance/Volume/IsometricBending/FEMTriangle/
//                                                     FEMTetra/StrainTetra/Dihedral/StrainTriangle)
//   PositionBasedDynamics/XPBD.cpp                   (XPBD variants)
//   PositionBasedDynamics/MathFunctions.cpp          (3x3 Jacobi eigen-decomposition, SVD with inversion handling)
// No tensor cores: the largest object is a 3x3 matrix per thread.  All absolute eps thresholds of the
// reference (1e-6) are kept literally, because they decide which projections are skipped.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pbdk {

#define PBD_EPS 1.0e-6f

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 xyz(const float4 &p) { return mk(p.x, p.y, p.z); }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return mk(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
__device__ __forceinline__ float sq(V3 a) { return dot(a, a); }
__device__ __forceinline__ float comp(const V3 &a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
// The library is compiled with -fmad=false so that a projection is the same arithmetic in every kernel it is inlined into
// (the execution modes are bit-identical, tested); every multiply-add that should fuse is therefore written as fmaf().
__device__ __forceinline__ V3 madd(V3 a, float s, V3 b) { return mk(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z)); }  // a s + b
__device__ __forceinline__ float fma3(float a0, float b0, float a1, float b1, float a2, float b2) { return fmaf(a0, b0, fmaf(a1, b1, a2 * b2)); }
// 1/x and a/b by MUFU.RCP (<= 1 ulp / 2 ulp): none of the reference's branches tests a quotient, only its operands
__device__ __forceinline__ float frcp(float x) { return __fdividef(1.0f, x); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdividef(a, b); }
// x += c only for dynamic particles (every <X>Constraint::solvePositionConstraint: "if (invMass != 0) x += corr")
__device__ __forceinline__ void apply(float4 &p, V3 c) {
    if (p.w != 0.0f) { p.x += c.x; p.y += c.y; p.z += c.z; }
}
__device__ __forceinline__ void apply(float4 &p, V3 g, float s) {  // x += g s
    if (p.w != 0.0f) { p.x = fmaf(g.x, s, p.x); p.y = fmaf(g.y, s, p.y); p.z = fmaf(g.z, s, p.z); }
}

struct M3 { float m[3][3]; };  // always indexed with compile-time constants after unrolling -> registers

__device__ __forceinline__ float det3(const M3 &a) {
    const float c0 = fmaf(a.m[1][1], a.m[2][2], -(a.m[1][2] * a.m[2][1]));
    const float c1 = fmaf(a.m[1][0], a.m[2][2], -(a.m[1][2] * a.m[2][0]));
    const float c2 = fmaf(a.m[1][0], a.m[2][1], -(a.m[1][1] * a.m[2][0]));
    return fmaf(a.m[0][0], c0, fmaf(a.m[0][2], c2, -(a.m[0][1] * c1)));
}
__device__ __forceinline__ M3 mul(const M3 &a, const M3 &b) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = fma3(a.m[i][0], b.m[0][j], a.m[i][1], b.m[1][j], a.m[i][2], b.m[2][j]);
    return r;
}
__device__ __forceinline__ M3 mulT(const M3 &a, const M3 &b) {  // a * b^T
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = fma3(a.m[i][0], b.m[j][0], a.m[i][1], b.m[j][1], a.m[i][2], b.m[j][2]);
    return r;
}
__device__ __forceinline__ M3 Tmul(const M3 &a, const M3 &b) {  // a^T * b
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = fma3(a.m[0][i], b.m[0][j], a.m[1][i], b.m[1][j], a.m[2][i], b.m[2][j]);
    return r;
}

// ---------------------------------------------------------------------------------------------------------
// Distance.  PositionBasedDynamics::solve_DistanceConstraint (PositionBasedDynamics.cpp:13-34)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void project_distance(float4 &p0, float4 &p1, float rest, float k) {
    const float wSum = p0.w + p1.w;
    if (wSum == 0.0f) return;
    V3 n = xyz(p1) - xyz(p0);
    const float d2 = sq(n);
    const float d = sqrtf(d2);
    if (d2 > 0.0f) n = n * frcp(d);  // Eigen normalize() leaves the zero vector untouched
    const float c = k * fdiv(d - rest, wSum);
    apply(p0, n, c * p0.w);
    apply(p1, n, -(c * p1.w));
}

// XPBD::solve_DistanceConstraint (XPBD.cpp:14-60).  alpha = 1/(k dt^2) (0 when k == 0) is a bucket uniform.
__device__ __forceinline__ void project_distance_xpbd(float4 &p0, float4 &p1, float rest, float alpha, float &lambda) {
    float K = p0.w + p1.w;
    V3 n = xyz(p0) - xyz(p1);
    const float d = sqrtf(sq(n));
    const float C = d - rest;
    if (!(d > 1.0e-6f)) return;
    K += alpha;
    if (!(fabsf(K) > 1.0e-6f)) return;
    const float dl = -fdiv(fmaf(alpha, lambda, C), K);
    lambda += dl;
    const float c = dl * frcp(d);  // pt = (n / d) dl
    apply(p0, n, c * p0.w);
    apply(p1, n, -(c * p1.w));
}

// ---------------------------------------------------------------------------------------------------------
// Volume.  PositionBasedDynamics::solve_VolumeConstraint (PositionBasedDynamics.cpp:104-142),
//          XPBD::solve_VolumeConstraint (XPBD.cpp:63-109)
// ---------------------------------------------------------------------------------------------------------
template <bool XPBD>
__device__ __forceinline__ void project_volume(float4 &q0, float4 &q1, float4 &q2, float4 &q3, float restVolume,
                                               float k, float alpha, float &lambda) {
    const V3 p0 = xyz(q0), p1 = xyz(q1), p2 = xyz(q2), p3 = xyz(q3);
    const float volume = (1.0f / 6.0f) * dot(cross(p1 - p0, p2 - p0), p3 - p0);
    if (!XPBD && k == 0.0f) return;
    const V3 g0 = cross(p1 - p2, p3 - p2);
    const V3 g1 = cross(p2 - p0, p3 - p0);
    const V3 g2 = cross(p0 - p1, p3 - p1);
    const V3 g3 = cross(p1 - p0, p2 - p0);
    float K = fmaf(q0.w, sq(g0), fmaf(q1.w, sq(g1), fmaf(q2.w, sq(g2), q3.w * sq(g3))));
    float s;
    if (XPBD) {
        K += alpha;
        if (fabsf(K) < PBD_EPS) return;
        const float dl = -fdiv(fmaf(alpha, lambda, volume - restVolume), K);
        lambda += dl;
        s = dl;
    } else {
        if (fabsf(K) < PBD_EPS) return;
        s = -fdiv(k * (volume - restVolume), K);
    }
    apply(q0, g0, s * q0.w);
    apply(q1, g1, s * q1.w);
    apply(q2, g2, s * q2.w);
    apply(q3, g3, s * q3.w);
}

// ---------------------------------------------------------------------------------------------------------
// Isometric bending.  PositionBasedDynamics::solve_IsometricBendingConstraint (PositionBasedDynamics.cpp:186-236),
// XPBD::solve_IsometricBendingConstraint (XPBD.cpp:153-213).
//
// The reference evaluates E = 1/2 sum Q_jk x_j.x_k and grad_j = sum_k Q_jk x_k on ABSOLUTE positions and relies on
// the rows of Q summing to zero; in fp32 that cancellation is catastrophic.  The matrix built by
// init_IsometricBendingConstraint (PositionBasedDynamics.cpp:145-183) is rank one, Q = coef K K^T with coef < 0 and
// sum(K) = 0.  With Kp = sqrt(-coef) K:   y = sum_k Kp_k (x_k - x_0),  E = -1/2 |y|^2,  grad_j = -Kp_j y.
// Same mathematics, translation invariant, 4 floats instead of 16.  Particle order of the constraint is
// (opp0, opp1, edge0, edge1) and the solver permutes to x = {p2, p3, p0, p1} (PositionBasedDynamics.cpp:195-196);
// Kp is stored in the solver's (permuted) order.
// ---------------------------------------------------------------------------------------------------------
template <bool XPBD>
__device__ __forceinline__ void project_isobending_rank1(float4 &q0, float4 &q1, float4 &q2, float4 &q3, float4 Kp,
                                                         float k, float alpha, float &lambda) {
    // solver order: x[0]=p2, x[1]=p3, x[2]=p0, x[3]=p1
    const V3 x0 = xyz(q2);
    const V3 y = madd(xyz(q3) - x0, Kp.y, madd(xyz(q0) - x0, Kp.z, (xyz(q1) - x0) * Kp.w));
    const float yy = sq(y);
    const float energy = -0.5f * yy;
    // sum_j w_j |grad_j|^2 = |y|^2 sum_j w_j Kp_j^2   (w_j == 0 contributes nothing, as in the reference's skip)
    const float wk0 = q0.w * Kp.z, wk1 = q1.w * Kp.w, wk2 = q2.w * Kp.x, wk3 = q3.w * Kp.y;  // w_j Kp_j
    float sum = yy * fmaf(wk2, Kp.x, fmaf(wk3, Kp.y, fmaf(wk0, Kp.z, wk1 * Kp.w)));
    float s;  // corr_j = s * w_j * grad_j = -s * w_j * Kp_j * y
    if (XPBD) {
        sum += alpha;
        if (!(fabsf(sum) > PBD_EPS)) return;
        const float dl = -fdiv(fmaf(alpha, lambda, energy), sum);
        lambda += dl;
        s = dl;
    } else {
        if (!(fabsf(sum) > PBD_EPS)) return;
        s = -k * fdiv(energy, sum);
    }
    apply(q0, y, -s * wk0);
    apply(q1, y, -s * wk1);
    apply(q2, y, -s * wk2);
    apply(q3, y, -s * wk3);
}

// General (user-modified) Q: literal evaluation as the reference does it.  Q rows in the solver's order.
template <bool XPBD>
__device__ __forceinline__ void project_isobending_fullq(float4 &q0, float4 &q1, float4 &q2, float4 &q3, float4 Q0,
                                                         float4 Q1, float4 Q2, float4 Q3, float k, float alpha,
                                                         float &lambda) {
    const V3 x[4] = {xyz(q2), xyz(q3), xyz(q0), xyz(q1)};
    const float w[4] = {q2.w, q3.w, q0.w, q1.w};
    const float Q[4][4] = {{Q0.x, Q0.y, Q0.z, Q0.w}, {Q1.x, Q1.y, Q1.z, Q1.w}, {Q2.x, Q2.y, Q2.z, Q2.w}, {Q3.x, Q3.y, Q3.z, Q3.w}};
    float energy = 0.0f;
    V3 g[4];
#pragma unroll
    for (int j = 0; j < 4; j++) g[j] = mk(0.f, 0.f, 0.f);
#pragma unroll
    for (int kk = 0; kk < 4; kk++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            energy += Q[j][kk] * dot(x[kk], x[j]);
            g[j] = g[j] + x[kk] * Q[j][kk];
        }
    energy *= 0.5f;
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (w[j] != 0.0f) sum += w[j] * sq(g[j]);
    float s;
    if (XPBD) {
        sum += alpha;
        if (!(fabsf(sum) > PBD_EPS)) return;
        const float dl = -(energy + alpha * lambda) / sum;
        lambda += dl;
        s = dl;
    } else {
        if (!(fabsf(sum) > PBD_EPS)) return;
        s = -k * (energy / sum);
    }
    apply(q0, g[2] * (s * w[2]));
    apply(q1, g[3] * (s * w[3]));
    apply(q2, g[0] * (s * w[0]));
    apply(q3, g[1] * (s * w[1]));
}

// ---------------------------------------------------------------------------------------------------------
// FEM triangle (orthotropic membrane StVK).  PositionBasedDynamics::solve_FEMTriangleConstraint
// (PositionBasedDynamics.cpp:844-930).  inv = m_invRestMat (2x2, row-major in a float4: 00,01,10,11).
// The elasticity tensor entries C00,C01,C10,C11,C22 depend only on the material -> precomputed per bucket.
// ---------------------------------------------------------------------------------------------------------
struct FemTriMaterial { float C00, C01, C10, C11, C22; };
__device__ __forceinline__ FemTriMaterial femtri_material(float Ex, float Ey, float Exy, float nuxy, float nuyx) {
    FemTriMaterial m;
    const float den = frcp(1.0f - nuxy * nuyx);
    m.C00 = Ex * den; m.C01 = Ex * nuyx * den; m.C11 = Ey * den; m.C10 = Ey * nuxy * den; m.C22 = Exy;
    return m;
}
__device__ __forceinline__ void project_femtriangle(float4 &q0, float4 &q1, float4 &q2, float area, float4 inv,
                                                    const FemTriMaterial &mat) {
    const V3 p13 = xyz(q0) - xyz(q2), p23 = xyz(q1) - xyz(q2);
    // F (3x2) = [p13 p23] * inv
    const V3 F0 = madd(p13, inv.x, p23 * inv.z);
    const V3 F1 = madd(p13, inv.y, p23 * inv.w);
    const float e00 = fmaf(0.5f, sq(F0), -0.5f);
    const float e11 = fmaf(0.5f, sq(F1), -0.5f);
    const float e01 = 0.5f * dot(F0, F1);
    const float s00 = fmaf(mat.C00, e00, mat.C01 * e11);
    const float s11 = fmaf(mat.C10, e00, mat.C11 * e11);
    const float s01 = mat.C22 * e01;
    // first Piola-Kirchhoff (3x2) = F * S
    const V3 P0 = madd(F0, s00, F1 * s01);
    const V3 P1 = madd(F0, s01, F1 * s11);
    const float psi = 0.5f * fmaf(e00, s00, fmaf(2.0f * e01, s01, e11 * s11));
    const float energy = area * psi;
    // H = area * P * inv^T ; gradC[0] = H col 0, gradC[1] = H col 1
    const V3 g0 = madd(P0, inv.x, P1 * inv.y) * area;
    const V3 g1 = madd(P0, inv.z, P1 * inv.w) * area;
    const V3 g2 = -g0 - g1;
    const float sum = fmaf(q0.w, sq(g0), fmaf(q1.w, sq(g1), q2.w * sq(g2)));
    if (!(fabsf(sum) > PBD_EPS)) return;
    const float s = -fdiv(energy, sum);
    apply(q0, g0, s * q0.w);
    apply(q1, g1, s * q1.w);
    apply(q2, g2, s * q2.w);
}

// ---------------------------------------------------------------------------------------------------------
// 3x3 symmetric Jacobi eigen-decomposition + SVD with inversion handling (only reached for inverted tets).
// MathFunctions::jacobiRotate / eigenDecomposition / svdWithInversionHandling (MathFunctions.cpp:11-43, 46-75, 261-388)
// ---------------------------------------------------------------------------------------------------------
template <int P, int Q>
__device__ __forceinline__ void jacobi_rotate(M3 &A, M3 &R) {
    if (A.m[P][Q] == 0.0f) return;
    const float d = (A.m[P][P] - A.m[Q][Q]) / (2.0f * A.m[P][Q]);
    float t = 1.0f / (fabsf(d) + sqrtf(d * d + 1.0f));
    if (d < 0.0f) t = -t;
    const float c = 1.0f / sqrtf(t * t + 1.0f);
    const float s = t * c;
    A.m[P][P] += t * A.m[P][Q];
    A.m[Q][Q] -= t * A.m[P][Q];
    A.m[P][Q] = A.m[Q][P] = 0.0f;
    constexpr int K = 3 - P - Q;  // the remaining index
    const float Akp = c * A.m[K][P] + s * A.m[K][Q];
    const float Akq = -s * A.m[K][P] + c * A.m[K][Q];
    A.m[K][P] = A.m[P][K] = Akp;
    A.m[K][Q] = A.m[Q][K] = Akq;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float Rkp = c * R.m[k][P] + s * R.m[k][Q];
        const float Rkq = -s * R.m[k][P] + c * R.m[k][Q];
        R.m[k][P] = Rkp;
        R.m[k][Q] = Rkq;
    }
}

__device__ __noinline__ void eigen_decomposition(const M3 &A, M3 &vecs, V3 &vals) {
    M3 D = A;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) vecs.m[i][j] = (i == j) ? 1.0f : 0.0f;
    for (int iter = 0; iter < 10; iter++) {
        int sel = 0;
        float mx = fabsf(D.m[0][1]);
        float a = fabsf(D.m[0][2]);
        if (a > mx) { sel = 1; mx = a; }
        a = fabsf(D.m[1][2]);
        if (a > mx) { sel = 2; mx = a; }
        if (mx < 1.0e-15f) break;
        if (sel == 0) jacobi_rotate<0, 1>(D, vecs);
        else if (sel == 1) jacobi_rotate<0, 2>(D, vecs);
        else jacobi_rotate<1, 2>(D, vecs);
    }
    vals = mk(D.m[0][0], D.m[1][1], D.m[2][2]);
}

__device__ __forceinline__ void negate_col(M3 &A, int c) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
        if (c == 0) A.m[r][0] = -A.m[r][0];
        else if (c == 1) A.m[r][1] = -A.m[r][1];
        else A.m[r][2] = -A.m[r][2];
    }
}
__device__ __forceinline__ int argmin3(float a, float b, float c) {  // first strict minimum below FLT_MAX, as the reference's scan
    int pos = 0; float mn = 3.402823466e+38f;
    if (a < mn) { pos = 0; mn = a; }
    if (b < mn) { pos = 1; mn = b; }
    if (c < mn) { pos = 2; mn = c; }
    return pos;
}

__device__ __noinline__ void svd_inversion(const M3 &A, float sig[3], M3 &U, M3 &VT) {
    M3 AtA = Tmul(A, A);
    M3 V; V3 S;
    eigen_decomposition(AtA, V, S);
    if (det3(V) < 0.0f) negate_col(V, argmin3(S.x, S.y, S.z));
    sig[0] = sqrtf(fmaxf(S.x, 0.0f)); sig[1] = sqrtf(fmaxf(S.y, 0.0f)); sig[2] = sqrtf(fmaxf(S.z, 0.0f));
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) VT.m[i][j] = V.m[j][i];
    int chk = 0, pos = 0;
#pragma unroll
    for (int l = 0; l < 3; l++)
        if (fabsf(sig[l]) < 1.0e-4f) { pos = l; chk++; }
    U = mul(A, V);
    if (chk > 1) {
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) U.m[i][j] = (i == j) ? 1.0f : 0.0f;
    } else {
#pragma unroll
        for (int l = 0; l < 3; l++) {
            if (chk == 1 && l == pos) continue;
            const float inv = 1.0f / sig[l];
#pragma unroll
            for (int m = 0; m < 3; m++) U.m[m][l] *= inv;
        }
        if (chk == 1) {
            // rebuild the degenerate column as the normalised cross product of the other two (in index order)
            V3 c0 = mk(U.m[0][0], U.m[1][0], U.m[2][0]), c1 = mk(U.m[0][1], U.m[1][1], U.m[2][1]), c2 = mk(U.m[0][2], U.m[1][2], U.m[2][2]);
            V3 v = (pos == 0) ? cross(c1, c2) : ((pos == 1) ? cross(c0, c2) : cross(c0, c1));
            const float z = sq(v);
            if (z > 0.0f) v = v * (1.0f / sqrtf(z));
#pragma unroll
            for (int m = 0; m < 3; m++) {
                const float val = comp(v, m);
                if (pos == 0) U.m[m][0] = val; else if (pos == 1) U.m[m][1] = val; else U.m[m][2] = val;
            }
        }
    }
    if (det3(U) < 0.0f) {
        const int p2 = argmin3(sig[0], sig[1], sig[2]);
        if (p2 == 0) sig[0] = -sig[0]; else if (p2 == 1) sig[1] = -sig[1]; else sig[2] = -sig[2];
        negate_col(U, p2);
    }
}

// ---------------------------------------------------------------------------------------------------------
// FEM tetrahedron (StVK energy constraint).  FEMTetConstraint::solvePositionConstraint (Constraints.cpp:1777-1825),
// PositionBasedDynamics::solve_FEMTetraConstraint (PositionBasedDynamics.cpp:1109-1169), computeGreenStrainAndPiolaStress
// (:958-1008), computeGradCGreen (:1011-1031), computeGreenStrainAndPiolaStressInversion (:1034-1104);
// XPBD_FEMTetConstraint (Constraints.cpp:1854-1906) + XPBD::solve_FEMTetraConstraint (XPBD.cpp:217-294).
// inv = m_invRestMat (row-major), restVolume = m_volume.  mu/lambda: Lame parameters (divided by E for XPBD).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fem_energy_gradients(const V3 &x1, const V3 &x2, const V3 &x3, const V3 &x4,
                                                     const M3 &inv, float restVolume, float mu, float lambda,
                                                     bool inversionBranch, float &energy, V3 J[4]) {
    const V3 p14 = x1 - x4, p24 = x2 - x4, p34 = x3 - x4;
    M3 F;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        F.m[0][c] = fma3(p14.x, inv.m[0][c], p24.x, inv.m[1][c], p34.x, inv.m[2][c]);
        F.m[1][c] = fma3(p14.y, inv.m[0][c], p24.y, inv.m[1][c], p34.y, inv.m[2][c]);
        F.m[2][c] = fma3(p14.z, inv.m[0][c], p24.z, inv.m[1][c], p34.z, inv.m[2][c]);
    }
    M3 sigma;
    float psi = 0.0f, trace;
    if (!inversionBranch) {
        M3 e;  // Green strain 1/2 (F^T F - I)
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = a; b < 3; b++) {
                const float s = fma3(F.m[0][a], F.m[0][b], F.m[1][a], F.m[1][b], F.m[2][a], F.m[2][b]);
                e.m[a][b] = e.m[b][a] = (a == b) ? fmaf(0.5f, s, -0.5f) : 0.5f * s;
            }
        trace = e.m[0][0] + e.m[1][1] + e.m[2][2];
        const float ltrace = lambda * trace, mu2 = 2.0f * mu;
        M3 S;
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) {
                S.m[a][b] = (a == b) ? fmaf(mu2, e.m[a][b], ltrace) : mu2 * e.m[a][b];
                psi = fmaf(e.m[a][b], e.m[a][b], psi);
            }
        sigma = mul(F, S);
    } else {
        float hatF[3]; M3 U, VT;
        svd_inversion(F, hatF, U, VT);
        float eh[3], sv[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (hatF[j] < 0.577f) hatF[j] = 0.577f;  // clamp small singular values
            eh[j] = 0.5f * (hatF[j] * hatF[j] - 1.0f);
        }
        trace = eh[0] + eh[1] + eh[2];
        const float ltrace = lambda * trace;
#pragma unroll
        for (int j = 0; j < 3; j++) sv[j] = hatF[j] * (2.0f * mu * eh[j] + ltrace);
        // epsilon = U diag(eh) VT ; sigma = U diag(sv) VT
        M3 e;
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) {
                e.m[a][b] = U.m[a][0] * eh[0] * VT.m[0][b] + U.m[a][1] * eh[1] * VT.m[1][b] + U.m[a][2] * eh[2] * VT.m[2][b];
                sigma.m[a][b] = U.m[a][0] * sv[0] * VT.m[0][b] + U.m[a][1] * sv[1] * VT.m[1][b] + U.m[a][2] * sv[2] * VT.m[2][b];
                psi += e.m[a][b] * e.m[a][b];
            }
    }
    psi = fmaf(mu, psi, 0.5f * lambda * trace * trace);
    energy = restVolume * psi;
    // H = sigma * inv^T * restVolume; J[c] = column c of H
    M3 H = mulT(sigma, inv);
    J[0] = mk(H.m[0][0], H.m[1][0], H.m[2][0]) * restVolume;
    J[1] = mk(H.m[0][1], H.m[1][1], H.m[2][1]) * restVolume;
    J[2] = mk(H.m[0][2], H.m[1][2], H.m[2][2]) * restVolume;
    J[3] = -J[0] - J[1] - J[2];
}

template <bool XPBD>
__device__ __forceinline__ void project_femtet(float4 &q0, float4 &q1, float4 &q2, float4 &q3, float restVolume,
                                               const M3 &inv, float E, float nu, float dt, float &multiplier) {
    if (E <= 0.0f) return;
    if (nu < 0.0f || nu > 0.49f) return;
    const V3 p0 = xyz(q0), p1 = xyz(q1), p2 = xyz(q2), p3 = xyz(q3);
    // currentVolume (Constraints.cpp:1795) and volume (PositionBasedDynamics.cpp:1133) are the same triple product
    const float volume = dot(cross(p1 - p0, p2 - p0), p3 - p0) * (1.0f / 6.0f);
    // handleInversion = (volume / restVolume) < 0.2 and the inversion branch = handleInversion && !(volume > 0).  With a positive
    // rest volume the second condition implies the first (quotient <= 0), so the IEEE division of the tested quotient is evaluated
    // only for non-positive rest volumes -- same decisions, one division less on the common path's dependency chain.
    const bool inversionBranch = !(volume > 0.0f) && (restVolume > 0.0f || (volume / restVolume) < 0.2f);
    float mu, lambda;
    if (XPBD) {  // Lame parameters divided by E (XPBD.cpp:247-248)
        mu = 0.5f * frcp(1.0f + nu);
        lambda = nu * frcp((1.0f + nu) * (1.0f - 2.0f * nu));
    } else {
        mu = 0.5f * E * frcp(1.0f + nu);
        lambda = E * nu * frcp((1.0f + nu) * (1.0f - 2.0f * nu));
    }
    float energy; V3 J[4];
    fem_energy_gradients(p0, p1, p2, p3, inv, restVolume, mu, lambda, inversionBranch, energy, J);
    float sum = fmaf(q0.w, sq(J[0]), fmaf(q1.w, sq(J[1]), fmaf(q2.w, sq(J[2]), q3.w * sq(J[3]))));
    float s;
    if (XPBD) {
        const float C = sqrtf(2.0f * energy);
        const float alpha = frcp(E * dt * dt);
        sum = fmaf(C * C, alpha, sum);
        if (sum < PBD_EPS) return;
        const float l = -fdiv(C * fmaf(alpha, multiplier, C), sum);
        multiplier += l;
        s = l;
    } else {
        if (sum < PBD_EPS) return;
        s = -fdiv(energy, sum);
    }
    apply(q0, J[0], s * q0.w);
    apply(q1, J[1], s * q1.w);
    apply(q2, J[2], s * q2.w);
    apply(q3, J[3], s * q3.w);
}

// ---------------------------------------------------------------------------------------------------------
// Strain based dynamics, tetrahedron.  PositionBasedDynamics::solve_StrainTetraConstraint
// (PositionBasedDynamics.cpp:713-805): six sequential sub-projections S_ij (i >= j) with internal Gauss-Seidel on the
// accumulated corrections; StrainTetConstraint passes scalar stiffness * Ones (Constraints.cpp:1962-1963).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void project_straintet(float4 &q0, float4 &q1, float4 &q2, float4 &q3, const M3 &inv,
                                                  float stretchK, float shearK, bool normStretch, bool normShear) {
    const V3 p0 = xyz(q0), p1 = xyz(q1), p2 = xyz(q2), p3 = xyz(q3);
    V3 c0 = mk(0, 0, 0), c1 = c0, c2 = c0, c3 = c0;
    const float w0 = q0.w, w1 = q1.w, w2 = q2.w, w3 = q3.w;
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j <= i; j++) {
            const V3 P0 = (p1 + c1) - (p0 + c0);
            const V3 P1 = (p2 + c2) - (p0 + c0);
            const V3 P2 = (p3 + c3) - (p0 + c0);
            const V3 fi = P0 * inv.m[0][i] + P1 * inv.m[1][i] + P2 * inv.m[2][i];
            const V3 fj = P0 * inv.m[0][j] + P1 * inv.m[1][j] + P2 * inv.m[2][j];
            float Sij = dot(fi, fj);
            float wi = 0.f, wj = 0.f, s1 = 0.f, s3 = 0.f;
            const bool ns = normShear && (i != j);
            if (ns) {
                wi = sqrtf(sq(fi)); wj = sqrtf(sq(fj));
                s1 = 1.0f / (wi * wj);
                s3 = s1 * s1 * s1;
            }
            V3 d[4];
            d[0] = mk(0, 0, 0);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                V3 dk = fj * inv.m[k][i] + fi * inv.m[k][j];
                if (ns) dk = dk * s1 - ((fi * (wj * wj)) * inv.m[k][i] + (fj * (wi * wi)) * inv.m[k][j]) * (Sij * s3);
                d[k + 1] = dk;
                d[0] = d[0] - dk;
            }
            if (ns) Sij *= s1;
            float lambda = w0 * sq(d[0]) + w1 * sq(d[1]) + w2 * sq(d[2]) + w3 * sq(d[3]);
            if (fabsf(lambda) < PBD_EPS) continue;
            if (i == j) {
                if (normStretch) { const float s = sqrtf(Sij); lambda = 2.0f * s * (s - 1.0f) / lambda * stretchK; }
                else lambda = (Sij - 1.0f) / lambda * stretchK;
            } else {
                lambda = Sij / lambda * shearK;
            }
            c0 = c0 - d[0] * (lambda * w0);
            c1 = c1 - d[1] * (lambda * w1);
            c2 = c2 - d[2] * (lambda * w2);
            c3 = c3 - d[3] * (lambda * w3);
        }
    }
    apply(q0, c0); apply(q1, c1); apply(q2, c2); apply(q3, c3);
}

// ---------------------------------------------------------------------------------------------------------
// Dihedral bending.  PositionBasedDynamics::solve_DihedralConstraint (PositionBasedDynamics.cpp:37-102)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void project_dihedral(float4 &q0, float4 &q1, float4 &q2, float4 &q3, float restAngle, float k) {
    if (q0.w == 0.0f && q1.w == 0.0f) return;
    const V3 p0 = xyz(q0), p1 = xyz(q1), p2 = xyz(q2), p3 = xyz(q3);
    const V3 e = p3 - p2;
    const float elen = sqrtf(sq(e));
    if (elen < PBD_EPS) return;
    const float invElen = 1.0f / elen;
    V3 n1 = cross(p2 - p0, p3 - p0); n1 = n1 * (1.0f / sq(n1));
    V3 n2 = cross(p3 - p1, p2 - p1); n2 = n2 * (1.0f / sq(n2));
    const V3 d0 = n1 * elen;
    const V3 d1 = n2 * elen;
    const V3 d2 = n1 * (dot(p0 - p3, e) * invElen) + n2 * (dot(p1 - p3, e) * invElen);
    const V3 d3 = n1 * (dot(p2 - p0, e) * invElen) + n2 * (dot(p2 - p1, e) * invElen);
    { const float z = sq(n1); if (z > 0.0f) n1 = n1 * (1.0f / sqrtf(z)); }
    { const float z = sq(n2); if (z > 0.0f) n2 = n2 * (1.0f / sqrtf(z)); }
    float dt_ = dot(n1, n2);
    dt_ = fminf(fmaxf(dt_, -1.0f), 1.0f);
    const float phi = acosf(dt_);
    float lambda = q0.w * sq(d0) + q1.w * sq(d1) + q2.w * sq(d2) + q3.w * sq(d3);
    if (lambda == 0.0f) return;
    lambda = (phi - restAngle) / lambda * k;
    if (dot(cross(n1, n2), e) > 0.0f) lambda = -lambda;
    apply(q0, d0 * (-q0.w * lambda));
    apply(q1, d1 * (-q1.w * lambda));
    apply(q2, d2 * (-q2.w * lambda));
    apply(q3, d3 * (-q3.w * lambda));
}

// ---------------------------------------------------------------------------------------------------------
// Strain based dynamics, triangle.  PositionBasedDynamics::solve_StrainTriangleConstraint
// (PositionBasedDynamics.cpp:584-688).  inv = m_invRestMat row-major in a float4 (00,01,10,11).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void project_straintriangle(float4 &q0, float4 &q1, float4 &q2, float4 inv, float kxx,
                                                       float kyy, float kxy, bool normStretch, bool normShear) {
    const V3 p0 = xyz(q0), p1 = xyz(q1), p2 = xyz(q2);
    const float w0 = q0.w, w1 = q1.w, w2 = q2.w;
    // c[i] = column i of inv (z = 0); inv(k,i): k row
    const float im[2][2] = {{inv.x, inv.y}, {inv.z, inv.w}};
    V3 c0 = mk(0, 0, 0), c1 = c0, c2 = c0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
#pragma unroll
        for (int j = 0; j <= i; j++) {
            // r[a] = (edge1[a], edge2[a]) with accumulated corrections; rc_i[a] = r[a] . c[i]
            const V3 e1 = (p1 + c1) - (p0 + c0), e2 = (p2 + c2) - (p0 + c0);
            const V3 rci = e1 * im[0][i] + e2 * im[1][i];
            const V3 rcj = e1 * im[0][j] + e2 * im[1][j];
            float Sij = dot(rci, rcj);
            V3 d[3];
            d[0] = mk(0, 0, 0);
#pragma unroll
            for (int k = 0; k < 2; k++) {
                d[k + 1] = rcj * im[k][i] + rci * im[k][j];
                d[0] = d[0] - d[k + 1];
            }
            if (i != j && normShear) {
                const float fi2 = sq(rci), fj2 = sq(rcj);
                const float fi = sqrtf(fi2), fj = sqrtf(fj2);
                d[0] = mk(0, 0, 0);
                const float s = Sij / (fi2 * fi * fj2 * fj);
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    d[k + 1] = d[k + 1] * (1.0f / (fi * fj));
                    d[k + 1] = d[k + 1] - (rci * (fj * fj)) * (im[k][i] * s);
                    d[k + 1] = d[k + 1] - (rcj * (fi * fi)) * (im[k][j] * s);
                    d[0] = d[0] - d[k + 1];
                }
                Sij = Sij / (fi * fj);
            }
            float lambda = w0 * sq(d[0]) + w1 * sq(d[1]) + w2 * sq(d[2]);
            if (lambda == 0.0f) continue;
            if (i == j) {
                const float kk = (i == 0) ? kxx : kyy;
                if (normStretch) { const float s = sqrtf(Sij); lambda = 2.0f * s * (s - 1.0f) / lambda * kk; }
                else lambda = (Sij - 1.0f) / lambda * kk;
            } else {
                lambda = Sij / lambda * kxy;
            }
            c0 = c0 - d[0] * (lambda * w0);
            c1 = c1 - d[1] * (lambda * w1);
            c2 = c2 - d[2] * (lambda * w2);
        }
    }
    apply(q0, c0); apply(q1, c1); apply(q2, c2);
}

// ---------------------------------------------------------------------------------------------------------
// Shape matching on 4-particle clusters.  ShapeMatchingConstraint::solvePositionConstraint (Constraints.cpp:2003-2028),
// PositionBasedDynamics::solve_ShapeMatchingConstraint (PositionBasedDynamics.cpp:500-558, allowStretch = false),
// MathFunctions::polarDecompositionStable (MathFunctions.cpp:180-255).  x0[] and w[] are the copies frozen into the
// constraint at creation (Constraints.cpp:1985-2001); corrections are divided by the number of clusters at the vertex.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float one_norm_cols(const M3 &A) {  // max column sum
    const float s0 = fabsf(A.m[0][0]) + fabsf(A.m[1][0]) + fabsf(A.m[2][0]);
    const float s1 = fabsf(A.m[0][1]) + fabsf(A.m[1][1]) + fabsf(A.m[2][1]);
    const float s2 = fabsf(A.m[0][2]) + fabsf(A.m[1][2]) + fabsf(A.m[2][2]);
    return fmaxf(s0, fmaxf(s1, s2));
}
__device__ __forceinline__ float inf_norm_rows(const M3 &A) {  // max row sum
    const float s0 = fabsf(A.m[0][0]) + fabsf(A.m[0][1]) + fabsf(A.m[0][2]);
    const float s1 = fabsf(A.m[1][0]) + fabsf(A.m[1][1]) + fabsf(A.m[1][2]);
    const float s2 = fabsf(A.m[2][0]) + fabsf(A.m[2][1]) + fabsf(A.m[2][2]);
    return fmaxf(s0, fmaxf(s1, s2));
}
__device__ __forceinline__ V3 row3(const M3 &A, int r) { return mk(A.m[r][0], A.m[r][1], A.m[r][2]); }
__device__ __forceinline__ void set_row3(M3 &A, int r, V3 v) { A.m[r][0] = v.x; A.m[r][1] = v.y; A.m[r][2] = v.z; }

// Mt holds M^T on entry; returns the rotation R = (converged Mt)^T
__device__ __noinline__ void polar_decomposition_stable(M3 Mt, float tolerance, M3 &R) {
    // ||M||_1 = ||M^T||_inf and vice versa
    float Mone = inf_norm_rows(Mt), Minf = one_norm_cols(Mt), Eone;
    M3 Adj;
    int guard = 0;
    do {
        set_row3(Adj, 0, cross(row3(Mt, 1), row3(Mt, 2)));
        set_row3(Adj, 1, cross(row3(Mt, 2), row3(Mt, 0)));
        set_row3(Adj, 2, cross(row3(Mt, 0), row3(Mt, 1)));
        float det = Mt.m[0][0] * Adj.m[0][0] + Mt.m[0][1] * Adj.m[0][1] + Mt.m[0][2] * Adj.m[0][2];
        if (fabsf(det) < 1.0e-12f) {
            int index = -1;
            if (sq(row3(Adj, 0)) > 1.0e-12f) index = 0;
            else if (sq(row3(Adj, 1)) > 1.0e-12f) index = 1;
            else if (sq(row3(Adj, 2)) > 1.0e-12f) index = 2;
            if (index < 0) {
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int j = 0; j < 3; j++) R.m[i][j] = (i == j) ? 1.0f : 0.0f;
                return;
            }
            // replace the degenerate row by the cross product of the other two, refresh the affected adjugate rows
            if (index == 0) { set_row3(Mt, 0, cross(row3(Mt, 1), row3(Mt, 2))); set_row3(Adj, 1, cross(row3(Mt, 2), row3(Mt, 0))); set_row3(Adj, 2, cross(row3(Mt, 0), row3(Mt, 1))); }
            else if (index == 1) { set_row3(Mt, 1, cross(row3(Mt, 2), row3(Mt, 0))); set_row3(Adj, 2, cross(row3(Mt, 0), row3(Mt, 1))); set_row3(Adj, 0, cross(row3(Mt, 1), row3(Mt, 2))); }
            else { set_row3(Mt, 2, cross(row3(Mt, 0), row3(Mt, 1))); set_row3(Adj, 0, cross(row3(Mt, 1), row3(Mt, 2))); set_row3(Adj, 1, cross(row3(Mt, 2), row3(Mt, 0))); }
            Mone = inf_norm_rows(Mt); Minf = one_norm_cols(Mt);
            det = Mt.m[0][0] * Adj.m[0][0] + Mt.m[0][1] * Adj.m[0][1] + Mt.m[0][2] * Adj.m[0][2];
        }
        const float AdjOne = one_norm_cols(Adj), AdjInf = inf_norm_rows(Adj);
        const float gamma = sqrtf(sqrtf((AdjOne * AdjInf) / (Mone * Minf)) / fabsf(det));
        const float g1 = gamma * 0.5f;
        const float g2 = 0.5f / (gamma * det);
        M3 E;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const float old = Mt.m[i][j];
                Mt.m[i][j] = g1 * old + g2 * Adj.m[i][j];
                E.m[i][j] = old - Mt.m[i][j];
            }
        Eone = one_norm_cols(E);
        Mone = one_norm_cols(Mt); Minf = inf_norm_rows(Mt);
    } while (Eone > Mone * tolerance && ++guard < 64);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) R.m[i][j] = Mt.m[j][i];
}

__device__ __forceinline__ void project_shapematching(float4 &q0, float4 &q1, float4 &q2, float4 &q3, float4 rc, float4 a0, float4 a1,
                                                      float4 a2, float4 w, float4 nc, float k) {
    const V3 x[4] = {xyz(q0), xyz(q1), xyz(q2), xyz(q3)};
    const V3 x0[4] = {mk(a0.x, a0.y, a0.z), mk(a0.w, a1.x, a1.y), mk(a1.z, a1.w, a2.x), mk(a2.y, a2.z, a2.w)};
    const float ww[4] = {w.x, w.y, w.z, w.w};
    const V3 restCm = mk(rc.x, rc.y, rc.z);
    V3 cm = mk(0.f, 0.f, 0.f);
    float wsum = 0.0f;
    float wi[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { wi[i] = 1.0f / (ww[i] + PBD_EPS); cm = cm + x[i] * wi[i]; wsum += wi[i]; }
    if (wsum == 0.0f) return;
    cm = cm * (1.0f / wsum);
    M3 At;  // A^T, A = sum w_i (x_i - cm)(x0_i - restCm)^T
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) At.m[r][c] = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const V3 q = x0[i] - restCm;
        const V3 p = (x[i] - cm) * wi[i];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) At.m[c][r] += comp(p, r) * comp(q, c);
    }
    M3 R;
    polar_decomposition_stable(At, PBD_EPS, R);
    const float inc[4] = {1.0f / nc.x, 1.0f / nc.y, 1.0f / nc.z, 1.0f / nc.w};
    float4 *qs[4] = {&q0, &q1, &q2, &q3};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const V3 q = x0[i] - restCm;
        const V3 goal = cm + mk(R.m[0][0] * q.x + R.m[0][1] * q.y + R.m[0][2] * q.z, R.m[1][0] * q.x + R.m[1][1] * q.y + R.m[1][2] * q.z,
                                R.m[2][0] * q.x + R.m[2][1] * q.y + R.m[2][2] * q.z);
        const V3 corr = (goal - x[i]) * k;
        if (ww[i] != 0.0f) { qs[i]->x += inc[i] * corr.x; qs[i]->y += inc[i] * corr.y; qs[i]->z += inc[i] * corr.z; }  // tests the frozen m_w, not the live invMass
    }
}

// ---------------------------------------------------------------------------------------------------------
// Rigid bodies in the coloured sweep.  Quaternions are float4 (x, y, z, w) with Eigen's semantics.
//   PositionBasedRigidBodyDynamics::computeMatrixK (PositionBasedRigidBodyDynamics.cpp:11-45)
//   solve_BallJoint (:212-262), solve_RigidBodyParticleBallJoint (:2168-2217), update_* (:188-210, 2149-2166)
//   RigidBody::rotationUpdated / updateInertiaW (Simulation/RigidBody.h:190-207): the world-space inverse inertia is a pure
//   function of the (normalised) rotation, so it is rebuilt from q where needed instead of being stored.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 qmul(float4 a, float4 b) {  // Hamilton product, (x,y,z,w)
    float4 r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
__device__ __forceinline__ float4 qnormalize(float4 q) {
    const float n = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    return make_float4(q.x / n, q.y / n, q.z / n, q.w / n);
}
__device__ __forceinline__ M3 qmatrix(float4 q) {
    M3 m;
    const float tx = 2.0f * q.x, ty = 2.0f * q.y, tz = 2.0f * q.z;
    const float twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    m.m[0][0] = 1.0f - (tyy + tzz); m.m[0][1] = txy - twz; m.m[0][2] = txz + twy;
    m.m[1][0] = txy + twz; m.m[1][1] = 1.0f - (txx + tzz); m.m[1][2] = tyz - twx;
    m.m[2][0] = txz - twy; m.m[2][1] = tyz + twx; m.m[2][2] = 1.0f - (txx + tyy);
    return m;
}
__device__ __forceinline__ V3 mvec(const M3 &a, V3 v) {
    return mk(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
              a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z);
}
__device__ __forceinline__ M3 world_tensor(const M3 &R, V3 d) {  // R diag(d) R^T
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = R.m[i][0] * d.x * R.m[j][0] + R.m[i][1] * d.y * R.m[j][1] + R.m[i][2] * d.z * R.m[j][2];
    return r;
}
__device__ __forceinline__ void matrix_k(V3 connector, float invMass, V3 x, const M3 &J, M3 &K) {
    if (invMass != 0.0f) {
        const V3 v = connector - x;
        const float a = v.x, b = v.y, c = v.z;
        const float j11 = J.m[0][0], j12 = J.m[0][1], j13 = J.m[0][2], j22 = J.m[1][1], j23 = J.m[1][2], j33 = J.m[2][2];
        K.m[0][0] = c * c * j22 - b * c * (j23 + j23) + b * b * j33 + invMass;
        K.m[0][1] = -(c * c * j12) + a * c * j23 + b * c * j13 - a * b * j33;
        K.m[0][2] = b * c * j12 - a * c * j22 - b * b * j13 + a * b * j23;
        K.m[1][0] = K.m[0][1];
        K.m[1][1] = c * c * j11 - a * c * (j13 + j13) + a * a * j33 + invMass;
        K.m[1][2] = -(b * c * j11) + a * c * j12 + a * b * j13 - a * a * j23;
        K.m[2][0] = K.m[0][2];
        K.m[2][1] = K.m[1][2];
        K.m[2][2] = b * b * j11 - a * b * (j12 + j12) + a * a * j22 + invMass;
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) K.m[i][j] = 0.0f;
    }
}
__device__ __forceinline__ V3 llt_solve3(const M3 &A, V3 rhs) {  // (K1 + K2).llt().solve(rhs)
    const float l00 = sqrtf(A.m[0][0]);
    const float l10 = A.m[1][0] / l00, l20 = A.m[2][0] / l00;
    const float l11 = sqrtf(A.m[1][1] - l10 * l10);
    const float l21 = (A.m[2][1] - l20 * l10) / l11;
    const float l22 = sqrtf(A.m[2][2] - l20 * l20 - l21 * l21);
    const float y0 = rhs.x / l00;
    const float y1 = (rhs.y - l10 * y0) / l11;
    const float y2 = (rhs.z - l20 * y0 - l21 * y1) / l22;
    const float z2 = y2 / l22;
    const float z1 = (y1 - l21 * z2) / l11;
    const float z0 = (y0 - l10 * z1 - l20 * z2) / l00;
    return mk(z0, z1, z2);
}
// x += invMass pt ; q += 1/2 (0, J (r x pt)) q ; normalise   (BallJoint::solvePositionConstraint, Constraints.cpp:106-121)
__device__ __forceinline__ void rb_correct(float4 &X, float4 &Q, const M3 &J, V3 r, V3 pt) {
    const V3 ot = mvec(J, cross(r, pt));
    const float4 dq = qmul(make_float4(ot.x, ot.y, ot.z, 0.0f), Q);
    X.x += X.w * pt.x; X.y += X.w * pt.y; X.z += X.w * pt.z;
    Q = qnormalize(make_float4(Q.x + 0.5f * dq.x, Q.y + 0.5f * dq.y, Q.z + 0.5f * dq.z, Q.w + 0.5f * dq.w));
}

// BallJoint between rigid bodies (X0,Q0) and (X1,Q1); l0/l1 = connectors in body space
// (noinline: a handful of joints per scene; as a real call their registers do not count against the kernels they are compiled into)
__device__ __noinline__ void project_balljoint(float4 &X0, float4 &Q0, V3 Iinv0, float4 &X1, float4 &Q1, V3 Iinv1, V3 l0, V3 l1) {
    const M3 R0 = qmatrix(Q0), R1 = qmatrix(Q1);
    const V3 x0 = xyz(X0), x1 = xyz(X1);
    const V3 c0 = mvec(R0, l0) + x0, c1 = mvec(R1, l1) + x1;  // update_BallJoint
    const M3 J0 = world_tensor(R0, Iinv0), J1 = world_tensor(R1, Iinv1);
    M3 K0, K1, K;
    matrix_k(c0, X0.w, x0, J0, K0);
    matrix_k(c1, X1.w, x1, J1, K1);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) K.m[i][j] = K0.m[i][j] + K1.m[i][j];
    const V3 pt = llt_solve3(K, c1 - c0);
    if (X0.w != 0.0f) rb_correct(X0, Q0, J0, c0 - x0, pt);
    if (X1.w != 0.0f) rb_correct(X1, Q1, J1, c1 - x1, -pt);
}

// RigidBodyParticleBallJoint: rigid body (X0,Q0) and particle p (xyz, invMass)
__device__ __noinline__ void project_rb_particle_balljoint(float4 &X0, float4 &Q0, V3 Iinv0, float4 &p, V3 l0) {
    const M3 R0 = qmatrix(Q0);
    const V3 x0 = xyz(X0);
    const V3 c0 = mvec(R0, l0) + x0;
    const M3 J0 = world_tensor(R0, Iinv0);
    M3 K;
    matrix_k(c0, X0.w, x0, J0, K);
    if (p.w != 0.0f) { K.m[0][0] += p.w; K.m[1][1] += p.w; K.m[2][2] += p.w; }
    const V3 pt = llt_solve3(K, xyz(p) - c0);
    if (X0.w != 0.0f) rb_correct(X0, Q0, J0, c0 - x0, pt);
    if (p.w != 0.0f) { p.x -= p.w * pt.x; p.y -= p.w * pt.y; p.z -= p.w * pt.z; }
}

}  // namespace pbdk
