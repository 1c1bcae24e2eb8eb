// Persistent "chain" kernel of the single-token decoder step (dec_chain.cu).
//
// Between two cross-attention launches a decoder layer is a chain of small dependent operators over R <= 128 token rows
// (reference graph src/whisper.cpp:2525-2799): out-projection + residual, LayerNorm, MLP (two GEMMs + GELU), the next
// layer's LayerNorm + QKV projection, masked self-attention with the K/V append, its out-projection, LayerNorm and the
// cross-attention query projection.  As separate launches each of them costs a dependent-launch latency (~8 us measured)
// for ~1 us of HBM traffic.  Here the whole chain is ONE cooperative launch of a persistent grid (2 CTAs per SM); the
// operators are "phases" separated by a grid-wide barrier (one atomic + an acquire spin, ~0.5 us).
#pragma once

#include "common.cuh"
#include "dec_kernels.h"
#include "tc_gemm.h"

namespace wb {

// Stream-K geometry of one GEMM phase.  The (n-tile, k-block) space is flattened into U units of one 128 x 64
// weight block each and cut into G equal contiguous ranges, one per CTA, so every SM streams the same number of weight
// bytes whatever N and K are.  A CTA emits one f32 partial 64x64 tile per output tile its range touches; whoever consumes
// the GEMM output (the residual/LayerNorm phase, the GELU phase, self-attention, the cross-attention kernel) adds the
// partial tiles of a tile in contributor order -- fixed by this geometry alone, hence bit-reproducible and atomic-free.
constexpr int SG_TILE_COLS = 128;                          // weight rows (= output columns) per tile
constexpr int SG_TILE_FLOATS = 128 * SG_TILE_COLS;         // one partial tile: up to 128 token rows, f32 (rows >= R unused)

struct SplitGeom {
    int tiles = 0;      // N / 128
    int kpt = 0;        // K / 64: k-blocks per output tile
    int U = 0;          // tiles * kpt
    int G = 0;          // CTAs taking part
    int maxc = 0;       // partial-tile slots reserved per output tile
};

#ifdef __CUDACC__
__device__ __forceinline__ int sg_cta_of(const SplitGeom & g, int u) {      // CTA whose range holds unit u
    return (int) ((((unsigned) u + 1u) * (unsigned) g.G - 1u) / (unsigned) g.U);      // host guarantees (U+1)*G < 2^32
}
// sum of the partial tiles of element quad (row r, columns col..col+3) of the GEMM output
__device__ __forceinline__ float4 sg_load4(const SplitGeom & g, const float * __restrict__ part, int r, int col) {
    const int ot = col / SG_TILE_COLS;
    const int first = sg_cta_of(g, ot * g.kpt), last = sg_cta_of(g, ot * g.kpt + g.kpt - 1);
    const float * p = part + ((size_t) ot * g.maxc) * SG_TILE_FLOATS + r * SG_TILE_COLS + (col % SG_TILE_COLS);
    float4 s = __ldcg(reinterpret_cast<const float4 *>(p));
    for (int j = 1; j <= last - first; ++j) {
        const float4 v = __ldcg(reinterpret_cast<const float4 *>(p + (size_t) j * SG_TILE_FLOATS));
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    return s;
}
#endif

// Input of a consumer that reads a GEMM's output straight from its partial tiles: value = (sum + bias) * scale.
struct SplitIn {
    const float * part = nullptr;
    const float * bias = nullptr;
    SplitGeom g;
    unsigned long long * trace = nullptr;   // development aid: [0] start of CTA (0,0), [1] latest CTA end (%globaltimer)
};

enum ChainPhaseType : int {
    CP_ROW = 0,      // x[r] (+)= bias + partial sums (or token + position embedding), optional LayerNorm -> out16
    CP_GEMM = 1,     // a[R][K] * w[N][K]^T in 128-column tiles: stream-K -> partial tiles, or direct -> bias/scale/GELU -> out16
    CP_SELF = 3,     // masked self-attention over the row's own cache (+ K/V append); a = q | k | v rows [R][3d] 16-bit
};

struct ChainPhase {
    int type = 0;
    int embed = 0;                 // CP_ROW: start from te[token] + pe[pos] instead of x
    int gelu = 0;                  // CP_GEMM direct
    int direct = 0;                // CP_GEMM: 0 stream-K -> partial tiles, 1 whole tiles -> bias / scale / GELU -> out16
    int N = 0, K = 0;              // CP_GEMM
    int tm = 0;                    // CP_GEMM: ChainParams::tm[2*tm] = activations map, [2*tm+1] = weights map
    const void * a = nullptr;  int lda = 0;     // CP_SELF: q | k | v rows
    SplitGeom g;                   // CP_GEMM: its own geometry; consumers: the producer's
    float * part = nullptr;        // partial tiles
    const float * bias = nullptr;
    float scale = 1.0f;  int scale_cols = 0;
    const float * ln_w = nullptr, * ln_b = nullptr;     // CP_ROW: LayerNorm weights (null: no LayerNorm output)
    void * out16 = nullptr;  int ldo16 = 0;
    size_t layer_off = 0;          // CP_SELF: element offset of the layer inside a sequence's self cache
};

constexpr int CHAIN_MAX_PHASES = 12;
constexpr int CHAIN_MAX_GEMMS = 6;

struct ChainCommon {
    int R = 0, d = 0, H = 0, n_ctx = 0;
    float eps = 1e-5f;
    int ref_f16_gelu = 0;
    float * x = nullptr;                 // residual stream [R][d] f32
    const DecRow * rows = nullptr;
    const void * te = nullptr;           // token embedding [n_vocab][d] 16-bit
    const float * pe = nullptr;          // positional embedding [n_ctx][d]
};

struct ChainParams {
    int n_phase = 0;
    ChainCommon c;
    unsigned * bar = nullptr;            // grid barrier counter (monotonic across launches)
    unsigned bar_base = 0;
    unsigned long long * trace = nullptr;   // development aid: [32] %globaltimer stamps of CTA 0 (phase starts, end, entry)
    ChainPhase ph[CHAIN_MAX_PHASES];
    TMap tm[2 * CHAIN_MAX_GEMMS];        // TMA descriptors of the GEMM phases (activations, weights)
};

struct ChainLauncher {
    int grid = 0;                // persistent CTAs (co-resident by construction: cooperative launch)
    bool pdl_ok = true;          // cooperative + programmatic serialization accepted by the driver
    unsigned * bar = nullptr;
    unsigned bar_count = 0;
    ~ChainLauncher();
};

// grid size (0 when the device cannot hold the kernel) -- also allocates the barrier counter
int chain_init(ChainLauncher & cl, DType dt);
SplitGeom chain_geom(int grid, int R, int N, int K, int min_units);
SplitGeom chain_geom_direct(int R, int N, int K);
size_t chain_part_floats(const SplitGeom & g, int R);
bool chain_launch(ChainLauncher & cl, DType dt, ChainParams & p, cudaStream_t stream);

}  // namespace wb
