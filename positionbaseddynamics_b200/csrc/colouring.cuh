// positionbaseddynamics_b200/csrc/colouring.cuh
//
// Exact greedy first-fit colouring of the constraint graph on the GPU.
//
// The reference colours sequentially in insertion order: constraint c takes the lowest colour that no earlier constraint
// sharing one of its bodies has taken (SimulationModel::initConstraintGroups, Simulation/SimulationModel.cpp:1033-1094).
// The colour of c therefore depends only on the colours of the EARLIER constraints incident to its bodies.  That is a
// dependency DAG whose edges are "c -> the next constraint on the same body"; processing it in topological wavefronts
// (Kahn) assigns exactly the sequential colours:
//   * incidence lists: (body, incidence) pairs, stable radix sort by body -> per body its incident constraints in insertion
//     order; from them, per constraint and body slot, the NEXT constraint on that body and whether a predecessor exists
//     (in-degree = number of body slots with a predecessor);
//   * constraints with in-degree 0 seed the first wavefront.  A wavefront is body-disjoint (two constraints sharing a body
//     are ordered by an edge), so it is coloured without races: colour = lowest zero bit of the OR of the bodies' used-colour
//     sets; the sets are updated and the in-degrees of the successors decremented with fire-and-forget atomics; a successor
//     whose in-degree reaches 0 joins the next wavefront.
// The depth of the DAG, not the constraint count, bounds the time (cfg2: 12,490 wavefronts of ~480 constraints), so the
// wavefront loop runs in ONE thread block with block-level barriers and the wavefront lists in shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pbdk {

constexpr unsigned kNoConstraint = 0xffffffffu;

struct ColourArgs {
    unsigned N, V, words;            // constraints, bodies (shared particle / rigid-body index space), 64-bit words per used-colour set
    const uint4 *body4;              // [N] bodies of the constraint (unused slots = kNoConstraint)
    const uint4 *next4;              // [N] per body slot: the next constraint on that body, kNoConstraint if none
    unsigned *indeg;                 // [N] body slots that still wait for a predecessor
    unsigned long long *used;        // [V * words] colours taken on the body
    unsigned *colour;                // [N] out
    unsigned *listA, *listB;         // [N] wavefront buffers (spill space of the shared-memory lists)
    unsigned *counters;              // [0] size of the seed wavefront in listA, [1] overflow flag (more words needed), [2] coloured count, [3] wavefronts
};

// body4 from the insertion-ordered CSR, and the sort input: key = body, value = incidence index m = 4 c + slot
__global__ void k_colour_expand(const unsigned *__restrict__ off, const unsigned *__restrict__ body, uint4 *__restrict__ body4,
                                unsigned *__restrict__ keys, unsigned *__restrict__ vals, unsigned N) {
    const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    unsigned b[4] = {kNoConstraint, kNoConstraint, kNoConstraint, kNoConstraint};
    const unsigned m0 = off[c], nb = off[c + 1] - m0;
    for (unsigned k = 0; k < nb && k < 4; k++) { b[k] = body[m0 + k]; keys[m0 + k] = b[k]; vals[m0 + k] = 4u * c + k; }
    body4[c] = make_uint4(b[0], b[1], b[2], b[3]);
}
// successor / predecessor of every incidence from the body-sorted pairs
__global__ void k_colour_links(const unsigned *__restrict__ keys, const unsigned *__restrict__ vals, unsigned M, unsigned *__restrict__ next,
                               unsigned *__restrict__ indeg) {
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M) return;
    const unsigned m = vals[p];
    next[m] = (p + 1 < M && keys[p + 1] == keys[p]) ? (vals[p + 1] >> 2) : kNoConstraint;  // next4 viewed as unsigned[4 N]
    if (p > 0 && keys[p - 1] == keys[p]) atomicAdd(indeg + (m >> 2), 1u);
}
__global__ void k_colour_seed(ColourArgs a) {
    const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.N) return;
    if (a.indeg[c] == 0u) a.listA[atomicAdd(a.counters, 1u)] = c;
}

constexpr int kColourThreads = 1024;
constexpr unsigned kColourListCap = 4096;  // wavefront entries kept in shared memory (two lists); longer wavefronts spill to global

__global__ void __launch_bounds__(kColourThreads, 1) k_colour_wavefronts(ColourArgs a) {
    __shared__ unsigned sList[2][kColourListCap];
    __shared__ unsigned nIn, nOut;
    unsigned cur = 0;
    unsigned *gin = a.listA, *gout = a.listB;
    if (threadIdx.x == 0) { nIn = a.counters[0]; nOut = 0; }
    __syncthreads();
    for (unsigned i = threadIdx.x; i < nIn && i < kColourListCap; i += blockDim.x) sList[0][i] = gin[i];
    __syncthreads();
    unsigned coloured = 0, fronts = 0;
    while (true) {
        const unsigned n = nIn;
        if (n == 0) break;
        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned c = (i < kColourListCap) ? sList[cur][i] : __ldcg(gin + i);
            const uint4 B = __ldg(a.body4 + c), X = __ldg(a.next4 + c);
            const unsigned b[4] = {B.x, B.y, B.z, B.w}, x[4] = {X.x, X.y, X.z, X.w};
            unsigned col = 0xffffffffu;
            for (unsigned w = 0; w < a.words && col == 0xffffffffu; w++) {
                unsigned long long m = 0ull;
#pragma unroll
                for (int k = 0; k < 4; k++) if (b[k] != kNoConstraint) m |= __ldcg(a.used + (size_t)b[k] * a.words + w);
                if (~m) col = w * 64u + (unsigned)(__ffsll((long long)~m) - 1);
            }
            if (col == 0xffffffffu) { a.counters[1] = 1u; col = 0u; }  // ran out of representable colours: the host retries with more words
            a.colour[c] = col;
#pragma unroll
            for (int k = 0; k < 4; k++) if (b[k] != kNoConstraint) atomicOr(a.used + (size_t)b[k] * a.words + (col >> 6), 1ull << (col & 63u));
            __threadfence_block();  // the used-colour sets before the release of the successors (same block: CTA scope is enough)
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (x[k] != kNoConstraint && atomicSub(a.indeg + x[k], 1u) == 1u) {
                    const unsigned o = atomicAdd(&nOut, 1u);
                    if (o < kColourListCap) sList[cur ^ 1u][o] = x[k]; else __stcg(gout + o, x[k]);
                }
        }
        __syncthreads();
        coloured += n; fronts++;
        if (threadIdx.x == 0) { nIn = nOut; nOut = 0; }
        cur ^= 1u;
        unsigned *t = gin; gin = gout; gout = t;
        __syncthreads();
    }
    if (threadIdx.x == 0) { a.counters[2] = coloured; a.counters[3] = fronts; }
}

}  // namespace pbdk
