// Host interface of the tcgen05/TMEM/TMA GEMM (tc_gemm.cu).
#pragma once

#include "common.cuh"

namespace wb {

// out = epilogue(A[M,K] * W[N,K]^T), A and W row-major 16-bit (f16 or bf16), f32 accumulation.
// epilogue order (mirrors the op order of the reference graphs, src/whisper.cpp:2006-2014, 2112-2237, 2300-2339):
//   v = acc + bias[n];  if (n < scale_cols) v *= scale;  if (gelu) v = gelu(v);
//   v += pos[(m % pos_rows)*N + n];  v += resid[m*ldr + n];  out32[m*ldo32+n] = v;  out16[m*ldo16+n] = (16-bit) v
struct GemmArgs {
    DType dtype = DType::F16;
    int M = 0, N = 0, K = 0;
    const void * a = nullptr;   int lda = 0;    // [M][lda]
    const void * w = nullptr;   int ldw = 0;    // [N][ldw]
    const float * bias = nullptr;               // [N] or null
    float scale = 1.0f;         int scale_cols = 0;
    bool gelu = false;
    const float * pos = nullptr; int pos_rows = 0;
    const float * resid = nullptr; int ldr = 0; // may alias out32
    void * out16 = nullptr;     int ldo16 = 0;
    // > 0 (tc_gemm only): out16 is the cross-K/V pool.  Rows are (window, t < head_major_T), columns (kv, head, 64); element
    // (w, t, kv, h, c) goes to out16[w * T * N + ((h * 2 + kv) * T + t) * 64 + c] -- each (window, head) stream contiguous.
    int head_major_T = 0;
    float * out32 = nullptr;    int ldo32 = 0;
};

// Returns false when the arguments violate the kernel's alignment contract or the launch failed.
bool tc_gemm(const GemmArgs & g, cudaStream_t stream);

// A CUtensorMap by another name (keeps <cuda.h> out of the headers): 2-D tiled map over a row-major 16-bit matrix
// [rows][ld_elems] with a box of box_rows x 64 elements and the 128-byte swizzle (the layout tcgen05 descriptors expect).
struct alignas(64) TMap {
    unsigned char bytes[128];
};
bool tc_make_tmap(TMap * tm, const void * base, int rows, int cols, int ld_elems, int box_rows, DType dt);
// 3-D variant: dims {d0 (contiguous), d1, d2}, byte strides of d1 and d2, box {box0, box1, 1}, 128-byte swizzle.
bool tc_make_tmap3d(TMap * tm, const void * base, size_t d0, size_t d1, size_t d2, size_t stride1_bytes, size_t stride2_bytes,
                    int box0, int box1, DType dt);

}  // namespace wb
