// tools/barrier_bench.cu -- development micro-benchmark (not part of the product): cost of one grid-wide barrier on B200
// for the variants considered for k_step_persistent.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=true
//   -o tools/barrier_bench tools/barrier_bench.cu ; run on the GPU box.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned long long ld_acq(const unsigned long long *p) {
    unsigned long long v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long *p) {
    unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}

template <int VARIANT>
__global__ void k_bar(unsigned long long *counter, unsigned long long base, int n, float4 *buf, int storesPerPhase) {
    unsigned long long target = base;
    cg::grid_group grid = cg::this_grid();
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < n; it++) {
        for (int s = 0; s < storesPerPhase; s++) __stcg(buf + tid + (size_t)s * gridDim.x * blockDim.x, make_float4(it, s, 0, 1));
        if (VARIANT == 3) { grid.sync(); continue; }
        target += gridDim.x;
        __syncthreads();
        if (threadIdx.x == 0) {
            if (VARIANT == 0 || VARIANT == 1 || VARIANT == 2) __threadfence();
            if (VARIANT == 4) asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(counter) : "memory");
            else atomicAdd(counter, 1ull);
            if (VARIANT == 2) { while (ld_acq(counter) < target) __nanosleep(32); }
            else if (VARIANT == 4) { while (ld_relaxed(counter) < target) { } asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
            else { while (ld_acq(counter) < target) { } }
            if (VARIANT == 0) __threadfence();
        }
        __syncthreads();
    }
}

template <int V> float run(int threads, int n, int stores, unsigned long long *ctr, float4 *buf, int sms) {
    cudaMemset(ctr, 0, 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    unsigned long long base = 0; int nn = n;
    void *args[] = {&ctr, &base, &nn, &buf, &stores};
    cudaLaunchCooperativeKernel((void *)k_bar<V>, dim3(sms), dim3(threads), args, 0, 0);  // warm-up
    cudaDeviceSynchronize();
    cudaMemset(ctr, 0, 8);
    cudaEventRecord(a);
    cudaLaunchCooperativeKernel((void *)k_bar<V>, dim3(sms), dim3(threads), args, 0, 0);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
    return ms * 1e3f / n;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    unsigned long long *ctr; cudaMalloc(&ctr, 256);
    float4 *buf; cudaMalloc(&buf, (size_t)sms * 1024 * 8 * sizeof(float4));
    const int n = 2000;
    printf("SMs %d; microseconds per barrier (n=%d)\n", sms, n);
    printf("%8s %7s | %12s %12s %12s %12s %12s\n", "threads", "stores", "2fence", "1fence", "1f+nanosleep", "cg.sync", "red.rel+acq");
    for (int threads : {256, 512, 1024})
        for (int stores : {0, 2, 8})
            printf("%8d %7d | %12.3f %12.3f %12.3f %12.3f %12.3f\n", threads, stores, run<0>(threads, n, stores, ctr, buf, sms),
                   run<1>(threads, n, stores, ctr, buf, sms), run<2>(threads, n, stores, ctr, buf, sms), run<3>(threads, n, stores, ctr, buf, sms),
                   run<4>(threads, n, stores, ctr, buf, sms));
    return 0;
}
