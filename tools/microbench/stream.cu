// Development aid: memory behaviour of the chain kernel's GEMM phases without any math.
// Every CTA streams `units` k-blocks: an activation tile (rows_x x 128 B, row stride ldx bytes, shared by all CTAs) and a
// weight tile (rows_w x 128 B, row stride ldw bytes, private rows), through an S-stage cp.async ring.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/stream tools/microbench/stream.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int S>
__global__ void __launch_bounds__(128, 2)
stream_kernel(const char * x, int ldx, int rows_x, const char * w, int ldw, int rows_w, int units, int do_x, int do_w, int depth, int fence) {
    extern __shared__ __align__(128) char sm[];
    const int tid = threadIdx.x, r0 = tid >> 3, ch = tid & 7;
    const char * wb = w + (size_t) blockIdx.x * rows_w * ldw;
    auto load = [&](int u) {
        char * st = sm + (u % S) * 24576;
        if (do_x) for (int r = r0; r < rows_x; r += 16) {
            unsigned s = (unsigned) __cvta_generic_to_shared(st + r * 128 + ch * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(x + (size_t) r * ldx + (size_t) u * 128 + ch * 16));
        }
        if (do_w) for (int r = r0; r < rows_w; r += 16) {
            unsigned s = (unsigned) __cvta_generic_to_shared(st + 16384 + r * 128 + ch * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(wb + (size_t) r * ldw + (size_t) u * 128 + ch * 16));
        }
    };
    for (int s = 0; s < depth; ++s) { if (s < units) load(s); asm volatile("cp.async.commit_group;" ::: "memory"); }
    for (int it = 0; it < units; ++it) {
        if (depth == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else if (depth == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 4;" ::: "memory");
        if (fence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (it + depth < units) load(it + depth);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

int main() {
    const int K = 12800, units = K / 64;      // long rows: 200 units per CTA amortise the launch
    char * x, * w;
    const size_t wbytes = (size_t) 9600 * K * 2 * 2;      // 2 different weight matrices to rotate through (cold each time)
    CK(cudaMalloc(&x, 128 * K * 2));
    CK(cudaMalloc(&w, wbytes));
    CK(cudaMemset(x, 1, 128 * K * 2));
    CK(cudaMemset(w, 1, wbytes));
    CK(cudaFuncSetAttribute(stream_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 24576));
    CK(cudaFuncSetAttribute(stream_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 24576));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    struct Cfg { int grid, rows_w, do_x, do_w, depth; const char * name; int fence = 0; };
    const Cfg cfgs[] = {
        {160, 32, 1, 1, 3, "direct nt=32, 160 CTAs, X+W, depth 3"},
        {160, 32, 0, 1, 3, "direct nt=32, 160 CTAs, W only"},
        {160, 32, 1, 0, 3, "direct nt=32, 160 CTAs, X only"},
        {120, 32, 1, 1, 3, "direct nt=32, 120 CTAs, X+W"},
        {80, 64, 1, 1, 3, "direct nt=64, 80 CTAs, X+W"},
        {160, 32, 1, 1, 5, "direct nt=32, 160 CTAs, X+W, depth 5"},
        {160, 32, 1, 1, 2, "direct nt=32, 160 CTAs, X+W, depth 2"},
        {160, 32, 1, 1, 3, "direct nt=32, 160 CTAs, X+W, depth 3 + fence.proxy.async", 1},
        {296, 32, 1, 1, 3, "nt=32, 296 CTAs, X+W, depth 3 + fence.proxy.async", 1},
        {148, 32, 1, 1, 3, "nt=32, 148 CTAs, X+W"},
        {296, 16, 1, 1, 3, "nt=16, 296 CTAs, X+W"},
    };
    for (const Cfg & c : cfgs) {
        float best = 1e9;
        for (int rep = 0; rep < 10; ++rep) {
            const char * wr = w + (size_t) (rep % 2) * 9600 * K * 2;
            CK(cudaEventRecord(e0));
            if (c.depth == 5) stream_kernel<6><<<c.grid, 128, 6 * 24576>>>(x, K * 2, 64, wr, K * 2, c.rows_w, units, c.do_x, c.do_w, c.depth, c.fence);
            else stream_kernel<4><<<c.grid, 128, 4 * 24576>>>(x, K * 2, 64, wr, K * 2, c.rows_w, units, c.do_x, c.do_w, c.depth, c.fence);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep >= 2 && ms < best) best = ms;
        }
        const double wb = (double) c.grid * c.rows_w * K * 2 * c.do_w, xb = (double) c.grid * 64 * K * 2 * c.do_x;
        printf("%-48s %7.2f us = %.3f us/unit  (W %.1f MB at %.0f GB/s, X %.1f MB)\n", c.name, best * 1e3, best * 1e3 / units, wb / 1e6, wb / (best * 1e-3) / 1e9, xb / 1e6);
    }
    return 0;
}
