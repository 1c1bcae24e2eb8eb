// Read-bandwidth probe for the cross-attention access pattern (development aid).
//   mode 0: the decoder's layout  -- CTA (window, head) reads 128-byte chunks at a 5120-byte stride (K rows, then V rows)
//   mode 1: head-major layout      -- CTA (window, head) reads one contiguous 384 KB block
// Same load structure as cross_attn_kernel: 128 threads, 8 lanes per 128-byte row, 8 independent 16-byte loads per lane.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o readbw readbw.cu ; run: ./readbw
#include <cstdio>
#include <cuda_runtime.h>

constexpr int W = 64, H = 20, T = 1500, D = 1280, L = 4;      // L layers resident: 4 x 491.5 MB, rotated to defeat L2

template <int MODE>
__global__ void __launch_bounds__(128, 9) read_kernel(const uint4 * __restrict__ base, int layer, unsigned * sink) {
    const int w = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, sub = lane & 7, grp = lane >> 3;
    unsigned acc = 0;
    for (int pass = 0; pass < 2; ++pass) {          // K sweep, V sweep
        for (int b = 0; b < T; b += 128) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = b + warp * 4 + grp + 16 * u;
                size_t off;                          // in 16-byte units
                if (MODE == 0) off = (((size_t) w * L + layer) * T + t) * (2 * D / 8) + pass * (D / 8) + h * 8 + sub;
                else off = ((((size_t) layer * W + w) * H + h) * 2 + pass) * (size_t) T * 8 + (size_t) t * 8 + sub;
                v[u] = t < T ? __ldg(base + off) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
    const size_t bytes = (size_t) L * W * T * 2 * D * 2;
    uint4 * buf;
    unsigned * sink;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            for (int i = 0; i < 32; ++i) {
                if (mode == 0) read_kernel<0><<<dim3(W, H), 128>>>(buf, i % L, sink);
                else read_kernel<1><<<dim3(W, H), 128>>>(buf, i % L, sink);
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double per = ms / 32 * 1e3, gb = (double) W * T * 2 * D * 2 / 1e9;
            printf("mode %d (%s): %.2f us per launch, %.0f GB/s\n", mode, mode ? "head-major contiguous" : "strided 128 B @ 5120 B", per,
                   gb / (per * 1e-6));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
