// positionbaseddynamics_b200/csrc/device_image.h
//
// The flattened device image of a simulation model that the sm_100a kernels consume (DESIGN.md "Data layout").
//  * particles : float4 SoA  pos = (x,y,z,invMass), vel = (vx,vy,vz,mass), oldPos, lastPos, pos0
//                (ParticleData::m_x/m_invMasses/m_v/m_masses/m_oldX/m_lastX/m_x0, Simulation/ParticleData.h:91-100)
//  * constraints: per type, SoA, ordered bucket after bucket in the reference's colour order
//                (SimulationModel::m_constraintGroups, Simulation/SimulationModel.h:99), each (colour,type) bucket
//                a contiguous range [first, first+count)
#pragma once
#include <cuda_runtime.h>
#include "../../include/pbd_b200.h"

namespace pbdk {

constexpr int kMaxMat = 5;   // material parameters per type (stiffness, Young's moduli, Poisson ratios, flags)
constexpr int kMaxGeoV = 6;  // float4 geometry arrays per type
constexpr int kMaxGeoS = 2;  // scalar geometry arrays per type

// Particle placement in the device arrays.  Layout 1 ("de-interleaved") puts the even-numbered particles first and the
// odd-numbered ones behind them.  Reason (measured, profiles/README.md): the step is bound by L2 *sector* throughput, and
// constraints of one colour never share a particle, so consecutive constraints of a bucket touch particles 2 apart --
// with the identity layout a warp's gather then uses only one 16-byte half of every 32-byte sector.  After de-interleaving
// a stride-2 pattern becomes contiguous (full sectors) and a stride-1 pattern becomes two contiguous streams (also full
// sectors).  Simulated sectors per gather on cfg2: 0.94 -> 0.72 (ideal 0.5).
#if defined(__CUDACC__)
__host__ __device__
#endif
inline unsigned particle_slot(unsigned i, unsigned n, int layout) {
    return layout == 1 ? (i >> 1) + (i & 1u) * ((n + 1u) >> 1) : i;
}

struct TypeArrays {
    const uint2 *idx2;         // 2-body types
    const uint4 *idx4;         // 4-body types
    const unsigned *idx3[3];   // 3-body types (SoA, 12 B per constraint)
    const float4 *gv[kMaxGeoV];
    const float *gs[kMaxGeoS];
    const float *mat[kMaxMat]; // per-constraint material arrays, or nullptr -> matU (uniform over the type)
    float matU[kMaxMat];
    float *lambda;             // XPBD multipliers (m_lambda), zeroed at the first sweep of each substep
    float4 *rbX, *rbQ;         // joint types only: rigid-body positions (xyz, invMass) and rotations (x,y,z,w)
    const float4 *rbIinv;      // joint types only: inverse principal moments
    int variant;               // PBD_ISOBENDING*: 0 = rank-1 Kp form, 1 = full 4x4 Q
};

struct Bucket {
    int type;
    unsigned first;   // offset into the type's arrays
    unsigned count;
    unsigned colour;  // index of the reference colour group
};

// static per-type shape of the image
struct TypeShape { int nBodies, nParams, nGeoV, nGeoS, nMat; bool xpbd; };

#if defined(__CUDACC__)
__host__ __device__
#endif
inline TypeShape type_shape(int t) {
    switch (t) {
    case PBD_DISTANCE:        return {2, 2, 0, 1, 1, false};
    case PBD_DISTANCE_XPBD:   return {2, 2, 0, 1, 1, true};
    case PBD_DIHEDRAL:        return {4, 2, 0, 1, 1, false};
    case PBD_ISOBENDING:      return {4, 17, 4, 0, 1, false};  // nGeoV: 1 (rank-1) or 4 (full Q), decided at flatten
    case PBD_ISOBENDING_XPBD: return {4, 17, 4, 0, 1, true};
    case PBD_FEMTRIANGLE:     return {3, 10, 1, 1, 5, false};
    case PBD_STRAINTRIANGLE:  return {3, 9, 1, 0, 5, false};
    case PBD_VOLUME:          return {4, 2, 0, 1, 1, false};
    case PBD_VOLUME_XPBD:     return {4, 2, 0, 1, 1, true};
    case PBD_FEMTET:          return {4, 12, 2, 2, 2, false};
    case PBD_FEMTET_XPBD:     return {4, 12, 2, 2, 2, true};
    case PBD_STRAINTET:       return {4, 13, 2, 1, 4, false};
    case PBD_SHAPEMATCHING:   return {4, 24, 6, 0, 1, false};
    case PBD_BALLJOINT:       return {2, 12, 2, 0, 0, false};
    case PBD_RB_PARTICLE_BALLJOINT: return {2, 6, 1, 0, 0, false};
    default:                  return {0, 0, 0, 0, 0, false};
    }
}

#define type_shape_dev type_shape

// Algorithmic bytes per projection: indices + particle float4 read + per-constraint constants
// + particle float4 write (+ lambda read/write for XPBD).  Isometric bending uses the rank-1 figure (4 floats Kp).
inline double algorithmic_bytes(int t, int isoVariant) {
    switch (t) {
    case PBD_DISTANCE:        return 8 + 32 + 4 + 32;
    case PBD_DISTANCE_XPBD:   return 8 + 32 + 4 + 32 + 8;
    case PBD_DIHEDRAL:        return 16 + 64 + 4 + 64;
    case PBD_ISOBENDING:      return 16 + 64 + (isoVariant ? 64 : 16) + 64;
    case PBD_ISOBENDING_XPBD: return 16 + 64 + (isoVariant ? 64 : 16) + 64 + 8;
    case PBD_FEMTRIANGLE:     return 12 + 48 + 20 + 48;
    case PBD_STRAINTRIANGLE:  return 12 + 48 + 16 + 48;
    case PBD_VOLUME:          return 16 + 64 + 4 + 64;
    case PBD_VOLUME_XPBD:     return 16 + 64 + 4 + 64 + 8;
    case PBD_FEMTET:          return 16 + 64 + 40 + 64;
    case PBD_FEMTET_XPBD:     return 16 + 64 + 40 + 64 + 8;
    case PBD_STRAINTET:       return 16 + 64 + 36 + 64;
    case PBD_SHAPEMATCHING:   return 16 + 64 + 96 + 64;
    case PBD_BALLJOINT:       return 8 + 2 * 48 + 32 + 2 * 32;   // two rigid bodies (x, q, Iinv) read, (x, q) written
    case PBD_RB_PARTICLE_BALLJOINT: return 8 + 48 + 16 + 16 + 32 + 16;  // restCm + frozen x0[4] + w[4] + numClusters[4]
    default:                  return 0;
    }
}

}  // namespace pbdk
