// TEST INFRASTRUCTURE (see cuda_runtime.h in this directory): runs pbdk::k_contacts for every particle, sequentially, on the host.
#include "contacts.cuh"
#include <cstdio>
#include <vector>
using namespace pbdk;
extern "C" void run_contacts(unsigned n, float4 *pos, float4 *vel, const float4 *rbX, const float4 *rbV, const float4 *rbW,
                             unsigned nRigid, const RigidCollider *rigid, unsigned nRanges, const ParticleCollider *ranges,
                             float tol, float stiff, unsigned iters, ContactRecord *rec, unsigned *recCount, unsigned recCap) {
    std::vector<unsigned> slot(n), start(nRanges + 1, 0);
    for (unsigned i = 0; i < n; i++) slot[i] = i;
    for (unsigned r = 0; r < nRanges; r++) start[r + 1] = start[r] + ranges[r].count;
    ContactArgs A{pos, vel, slot.data(), rbX, rbV, rbW, rigid, nRigid, ranges, nRanges, start.data(), start[nRanges], tol, stiff, iters, rec, recCount, recCap};
    blockDim.x = 1;
    for (unsigned g = 0; g < A.total; g++) { blockIdx.x = g; threadIdx.x = 0; k_contacts(A); }
}
