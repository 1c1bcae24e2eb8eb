// Decoder-step GEMM (M <= 128 rows) on tcgen05 / TMEM fed by TMA (tc_skinny.cu); same argument struct, grid, cluster
// reduction and epilogue as skinny_gemm (no `pos`).
#pragma once

#include "tc_gemm.h"

namespace wb {

bool tc_skinny_usable(const GemmArgs & g);
bool tc_skinny_gemm(const GemmArgs & g, cudaStream_t stream);

}  // namespace wb
