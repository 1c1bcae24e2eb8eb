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
    // > 0 (tc_gemm only, encoder QKV GEMM): columns >= vt_col0 (the V third) are NOT written to out16 but, transposed, to the
    // attention kernel's V^T scratch vt[(w * vt_H + head) * 80 + c][vt_TP] at key position t (row = w * vt_T + t): the
    // per-layer transpose kernel disappears.  The 32 rows of a warp are 32 consecutive keys = one 64-byte segment per column.
    void * vt = nullptr; int vt_col0 = 0, vt_T = 0, vt_TP = 0, vt_H = 0;
    // ---- LayerNorm folded algebraically into the decoder-step GEMMs (tc_skinny only; see tc_skinny.cu) ----
    //   LN(x) W^T + b = rstd * ((x * gamma) W^T - mean * c) + b',  c[n] = sum_k gamma[k] W[n][k],  b'[n] = b[n] + sum_k beta[k] W[n][k]
    // producer side (the GEMM that writes the residual stream out32): per 64-column tile and row, the tile's mean and centred sum
    // of squares of out32 -> ln_part_out[tile * M + row]; out16 = 16-bit(out32 * out16_gamma[n]) (the next LayerNorm's gamma)
    float2 * ln_part_out = nullptr;
    const float * out16_gamma = nullptr;
    // consumer side: a = the producer's out16 rows; ln_part_in = its ln_part_out with ln_parts = K / 64 tiles per row;
    // ln_colsum = c; bias must be b'
    const float2 * ln_part_in = nullptr; int ln_parts = 0;
    const float * ln_colsum = nullptr;
    float ln_eps = 1e-5f;
    // ---- L2 prefetch of the NEXT cross-attention's K stream, carried by the decoder-step GEMMs that run before it (tc_skinny only) ----
    // Between two cross-attention launches HBM is ~85 % idle (the GEMM chain is latency-bound), so an idle lane of every GEMM CTA
    // asks L2 (cp.async.bulk.prefetch.L2, evict-last) for the first pf_chunks x 16 KB of each (row, head) K block of the coming
    // launch; the cross-attention kernel streams with evict-first so that its own 490 MB do not push them out before use.
    // Chunk ids [0, pf_R * pf_H * pf_chunks) are dealt to pf_slots launches; this one is pf_slot.  pf_rows: the step's device rows.
    const void * pf_rows = nullptr;          // const DecRow *
    int pf_R = 0, pf_H = 0, pf_chunks = 0, pf_slot = 0, pf_slots = 1;
    size_t pf_layer_off_bytes = 0; int pf_head_bytes = 0;
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
