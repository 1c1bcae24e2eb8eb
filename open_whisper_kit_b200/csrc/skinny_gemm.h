// Weight-streaming GEMM for M <= 128 rows (skinny_gemm.cu); same argument struct and epilogue as tc_gemm (no `pos`).
#pragma once

#include "tc_gemm.h"

namespace wb {

struct SkinnyWorkspace {
    void * partial = nullptr;   size_t partial_cap = 0;    // split-K partial tiles (f32)
    void * counters = nullptr;  size_t counters_cap = 0;   // per-tile arrival tickets, zero between launches
    SkinnyWorkspace() = default;
    SkinnyWorkspace(const SkinnyWorkspace &) = delete;
    SkinnyWorkspace & operator=(const SkinnyWorkspace &) = delete;
    ~SkinnyWorkspace();
};

bool skinny_gemm(const GemmArgs & g, SkinnyWorkspace & ws, cudaStream_t stream);

}  // namespace wb
