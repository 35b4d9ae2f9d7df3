// TEST INFRASTRUCTURE: a stand-in for <cuda_runtime.h> that lets the body of csrc/contacts.cuh (plain C++ apart from the CUDA qualifiers)
// be compiled for the host by tests/test_contacts_host_shim.py, one "thread" at a time, so that the CPU test suite can hold the kernel's
// arithmetic against the reference without a GPU.  Never part of the product: libpbd_b200.so has no CPU path.
#pragma once
#include <cmath>
#include <algorithm>
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(x)
struct float4 { float x, y, z, w; };
struct uint3_ { unsigned x, y, z; };
static thread_local uint3_ blockIdx, blockDim, threadIdx;
static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p += v; return o; }
