// Development aid: ground-truth latencies on the box (grid barrier, L2 / DRAM pointer chase, cp.async tile fetch).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/latency tools/microbench/latency.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void bar_lean(unsigned * bar, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned v;
        do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int) (v - target) < 0);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}
__device__ __forceinline__ void bar_fat(unsigned * bar, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while ((int) (*(volatile unsigned *) bar - target) < 0) {}
        __threadfence();
    }
    __syncthreads();
}
// mode 0 lean, 1 fat, 2 lean + each CTA writes 16 KB before arriving
__global__ void barrier_kernel(unsigned * bar, int n, int mode, float4 * scratch) {
    unsigned target = 0;
    for (int i = 0; i < n; ++i) {
        target += gridDim.x;
        if (mode == 2) {
            float4 * dst = scratch + (size_t) blockIdx.x * 1024;
            for (int e = threadIdx.x; e < 1024; e += blockDim.x) __stcg(dst + e, make_float4(i, e, 0, 0));
        }
        if (mode == 1) bar_fat(bar, target); else bar_lean(bar, target);
    }
}
__global__ void chase_kernel(const unsigned * __restrict__ next, int n, unsigned * out, long long * cycles) {
    unsigned p = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) p = __ldcg(next + p);
    long long t1 = clock64();
    *out = p;
    *cycles = t1 - t0;
}
// every CTA fetches `kb` KB with cp.async (16 B per thread per op), waits, repeats `n` times over fresh addresses
__global__ void tile_kernel(const uint4 * __restrict__ src, size_t stride16, int kb, int n, long long * cycles) {
    extern __shared__ uint4 sm[];
    const uint4 * base = src + (size_t) blockIdx.x * stride16;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        for (int e = threadIdx.x; e < kb * 64; e += blockDim.x) {
            unsigned s = (unsigned) __cvta_generic_to_shared(sm + e);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(base + (size_t) i * kb * 64 + e));
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    int n_sm = 0, clk = 0;
    CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("SMs %d, clock %d kHz\n", n_sm, clk);
    unsigned * bar; float4 * scratch;
    CK(cudaMalloc(&bar, 256));
    CK(cudaMalloc(&scratch, (size_t) 2 * n_sm * 16384));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int per_sm = 1; per_sm <= 2; ++per_sm)
        for (int mode = 0; mode < 3; ++mode) {
            int grid = per_sm * n_sm, n = 2000, threads = 128;
            void * args[] = {&bar, &n, &mode, &scratch};
            float best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaMemset(bar, 0, 4));
                CK(cudaEventRecord(e0));
                CK(cudaLaunchCooperativeKernel((void *) barrier_kernel, dim3(grid), dim3(threads), args, 0, 0));
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                best = ms < best ? ms : best;
            }
            printf("barrier grid=%d mode=%d (%s): %.3f us each\n", grid, mode, mode == 0 ? "lean" : mode == 1 ? "threadfence+atomicAdd" : "lean + 16 KB store", best * 1e3 / n);
        }
    // pointer chase over 8 MB (L2 resident) and 2 GB (DRAM)
    for (size_t bytes : {(size_t) 8 << 20, (size_t) 2048 << 20}) {
        size_t n = bytes / 4;
        std::vector<unsigned> h(n);
        // stride permutation with a large odd step: visits lines far apart
        size_t step = (n / 2 + 12345) | 1;
        // build cycle: next[i] = (i + step) % n   (step odd & n power of two -> full cycle)
        for (size_t i = 0; i < n; ++i) h[i] = (unsigned) ((i + step) % n);
        unsigned * d; unsigned * out; long long * cyc;
        CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, 8));
        CK(cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice));
        int iters = 20000;
        chase_kernel<<<1, 1>>>(d, iters, out, cyc);      // warm (L2 case)
        chase_kernel<<<1, 1>>>(d, iters, out, cyc);
        CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        printf("chase over %zu MB: %.0f cycles per dependent load\n", bytes >> 20, (double) c / iters);
        cudaFree(d); cudaFree(out); cudaFree(cyc);
    }
    // cp.async tile fetch latency: all CTAs at once, fresh DRAM addresses
    {
        size_t per_cta = (size_t) 64 << 20 >> 4;     // 64 MB apart in uint4 units? keep total bounded: 296 * 1 MB
        per_cta = ((size_t) 1 << 20) / 16;
        uint4 * src; long long * cyc;
        CK(cudaMalloc(&src, (size_t) 2 * n_sm * per_cta * 16)); CK(cudaMalloc(&cyc, 8));
        CK(cudaMemset(src, 1, (size_t) 2 * n_sm * per_cta * 16));
        CK(cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        for (int grid : {1, n_sm, 2 * n_sm})
            for (int kb : {8, 16, 48}) {
                int n = (1 << 20) / (kb * 1024);
                n = n > 16 ? 16 : n;
                // flush L2 by touching another buffer
                CK(cudaMemset(scratch, 0, (size_t) 2 * n_sm * 16384));
                tile_kernel<<<grid, 128, 64 * 1024>>>(src, per_cta, kb, n, cyc);
                CK(cudaDeviceSynchronize());
                long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
                printf("cp.async fetch grid=%d %d KB per CTA per round (cold): %.0f cycles per round (%.2f us @1.9GHz), %.0f GB/s aggregate\n", grid, kb,
                       (double) c / n, (double) c / n / 1900.0, (double) grid * kb * 1024 / ((double) c / n / 1.9e9) / 1e9);
            }
    }
    return 0;
}
