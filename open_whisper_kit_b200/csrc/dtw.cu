// [EXPERIMENTAL in the reference] DTW token timestamps (whisper_context_params::dtw_token_timestamps, whisper_token_data::t_dtw).
//
// Reference: whisper_exp_compute_token_level_timestamps_dtw + dtw_and_backtrace + median_filter, src/whisper.cpp:8683-8998, and
// the alignment-head branch of the decoder graph, 2721-2737 / 2822-2828.  After a window's segments are known the reference
// decodes [sot, (lang), notimestamps, text tokens ..., eot] once more, keeps the cross-attention probabilities of the model's
// "alignment heads", normalises them per audio position over the tokens, median-filters along time, averages the heads and
// runs dynamic time warping over (token, audio position); a token's t_dtw is the first audio position of its run on the path.
//
// Here: the extra decode is one prompt-style batch through the engine with a capture hook -- cross_align_kernel recomputes
// softmax(q k^T) for the alignment heads only (a few (layer, head) pairs; everything else goes through the streaming
// cross-attention kernel, which never materialises probabilities) -- and the alignment itself runs on the host.
#include "dtw.h"

#include <math.h>

#include <algorithm>

namespace wb {

// ---- alignment-head presets (reference src/whisper.cpp:384-410: one list of {text layer, head} per released checkpoint) ----
namespace {
struct Head { int layer, head; };
const Head kTinyEn[] = {{1, 0}, {2, 0}, {2, 5}, {3, 0}, {3, 1}, {3, 2}, {3, 3}, {3, 4}};
const Head kTiny[] = {{2, 2}, {3, 0}, {3, 2}, {3, 3}, {3, 4}, {3, 5}};
const Head kBaseEn[] = {{3, 3}, {4, 7}, {5, 1}, {5, 5}, {5, 7}};
const Head kBase[] = {{3, 1}, {4, 2}, {4, 3}, {4, 7}, {5, 1}, {5, 2}, {5, 4}, {5, 6}};
const Head kSmallEn[] = {{6, 6}, {7, 0}, {7, 3}, {7, 8}, {8, 2}, {8, 5}, {8, 7}, {9, 0}, {9, 4}, {9, 8}, {9, 10}, {10, 0}, {10, 1}, {10, 2},
                         {10, 3}, {10, 6}, {10, 11}, {11, 2}, {11, 4}};
const Head kSmall[] = {{5, 3}, {5, 9}, {8, 0}, {8, 4}, {8, 7}, {8, 8}, {9, 0}, {9, 7}, {9, 9}, {10, 5}};
const Head kMediumEn[] = {{11, 4}, {14, 1}, {14, 12}, {14, 14}, {15, 4}, {16, 0}, {16, 4}, {16, 9}, {17, 12}, {17, 14}, {18, 7}, {18, 10},
                          {18, 15}, {20, 0}, {20, 3}, {20, 9}, {20, 14}, {21, 12}};
const Head kMedium[] = {{13, 15}, {15, 4}, {15, 15}, {16, 1}, {20, 0}, {23, 4}};
const Head kLargeV1[] = {{9, 19}, {11, 2}, {11, 4}, {11, 17}, {22, 7}, {22, 11}, {22, 17}, {23, 2}, {23, 15}};
const Head kLargeV2[] = {{10, 12}, {13, 17}, {16, 11}, {16, 12}, {16, 13}, {17, 15}, {17, 16}, {18, 4}, {18, 11}, {18, 19}, {19, 11},
                         {21, 2}, {21, 3}, {22, 3}, {22, 9}, {22, 12}, {23, 5}, {23, 7}, {23, 13}, {25, 5}, {26, 1}, {26, 12}, {27, 15}};
const Head kLargeV3[] = {{7, 0}, {10, 17}, {12, 18}, {13, 12}, {16, 1}, {17, 14}, {19, 11}, {21, 4}, {24, 1}, {25, 6}};
const Head kLargeV3Turbo[] = {{2, 4}, {2, 11}, {3, 3}, {3, 6}, {3, 11}, {3, 14}};

template <size_t N> void append(std::vector<Head> & out, const Head (&a)[N]) { out.insert(out.end(), a, a + N); }
}  // namespace

bool dtw_alignment_heads(const whisper_context_params & cp, int n_text_layer, int n_head, std::vector<std::vector<int>> & by_layer) {
    by_layer.assign(n_text_layer, {});
    std::vector<Head> heads;
    switch (cp.dtw_aheads_preset) {
        case WHISPER_AHEADS_NONE:
            wlog(GGML_LOG_LEVEL_ERROR, "%s: dtw_aheads_preset should be != DTW_AHEADS_NONE\n", __func__);
            return false;
        case WHISPER_AHEADS_N_TOP_MOST:
            if (cp.dtw_n_top > n_text_layer || cp.dtw_n_top <= 0) {
                wlog(GGML_LOG_LEVEL_ERROR, "%s: dtw_n_top must be between %d and %d for this model.", __func__, 1, n_text_layer);
                return false;
            }
            for (int il = n_text_layer - cp.dtw_n_top; il < n_text_layer; ++il)
                for (int h = 0; h < n_head; ++h) heads.push_back({il, h});
            break;
        case WHISPER_AHEADS_CUSTOM:
            if (cp.dtw_aheads.n_heads == 0 || cp.dtw_aheads.heads == nullptr) {
                wlog(GGML_LOG_LEVEL_ERROR, "%s: dtw_aheads is empty\n", __func__);
                return false;
            }
            for (size_t i = 0; i < cp.dtw_aheads.n_heads; ++i) heads.push_back({cp.dtw_aheads.heads[i].n_text_layer, cp.dtw_aheads.heads[i].n_head});
            break;
        case WHISPER_AHEADS_TINY_EN: append(heads, kTinyEn); break;
        case WHISPER_AHEADS_TINY: append(heads, kTiny); break;
        case WHISPER_AHEADS_BASE_EN: append(heads, kBaseEn); break;
        case WHISPER_AHEADS_BASE: append(heads, kBase); break;
        case WHISPER_AHEADS_SMALL_EN: append(heads, kSmallEn); break;
        case WHISPER_AHEADS_SMALL: append(heads, kSmall); break;
        case WHISPER_AHEADS_MEDIUM_EN: append(heads, kMediumEn); break;
        case WHISPER_AHEADS_MEDIUM: append(heads, kMedium); break;
        case WHISPER_AHEADS_LARGE_V1: append(heads, kLargeV1); break;
        case WHISPER_AHEADS_LARGE_V2: append(heads, kLargeV2); break;
        case WHISPER_AHEADS_LARGE_V3: append(heads, kLargeV3); break;
        case WHISPER_AHEADS_LARGE_V3_TURBO: append(heads, kLargeV3Turbo); break;
        default: return false;
    }
    // the reference concatenates the captured heads layer by layer, inside a layer in list order
    for (const Head & h : heads) {
        if (h.layer < 0 || h.layer >= n_text_layer) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: tried to set alignment head on text layer %d, but model only has %d text layers", __func__,
                 h.layer + 1, n_text_layer);
            return false;
        }
        if (h.head < 0 || h.head >= n_head) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: tried to set alignment head on head %d, but model only has %d heads", __func__, h.head + 1, n_head);
            return false;
        }
        by_layer[h.layer].push_back(h.head);
    }
    return true;
}

// ---- device: probabilities of the alignment heads -----------------------------------------------------------------------------
namespace {

// One CTA (256 threads) per (decoder row, alignment head of this layer): scores over the T audio positions of the row's window,
// softmax in f32, probabilities to out[(a0 + blockIdx.y) * R + row][t].  K sits in the head-major cross pool [head][K | V][T][64]
// (pre-scaled by dh^-0.25, as the streaming kernel reads it).
template <typename T16>
__global__ void __launch_bounds__(256)
cross_align_kernel(const T16 * __restrict__ q, int ldq, const DecRow * __restrict__ rows, const int * __restrict__ heads, size_t layer_off,
                   int T, float kq_scale, int R, int a0, float * __restrict__ out) {
    extern __shared__ float s_sc[];                   // [T]
    __shared__ float s_q[64];
    __shared__ float s_red[8];
    const int r = blockIdx.x, h = heads[blockIdx.y], tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DecRow row = rows[r];
    if (tid < 64) s_q[tid] = Half16<T16>::to_f(q[(size_t) r * ldq + h * 64 + tid]);
    __syncthreads();
    const T16 * kb = reinterpret_cast<const T16 *>(row.cross_kv) + layer_off + (size_t) h * 2 * T * 64;
    float mx = -INFINITY;
    for (int t = tid; t < T; t += 256) {
        const uint4 * kr = reinterpret_cast<const uint4 *>(kb + (size_t) t * 64);
        float acc = 0.0f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 u = __ldg(kr + c);
            const T16 * e = reinterpret_cast<const T16 *>(&u);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = fmaf(s_q[c * 8 + j], Half16<T16>::to_f(e[j]), acc);
        }
        acc *= kq_scale;
        s_sc[t] = acc;
        mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = s_red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, s_red[i]);
    __syncthreads();
    float sum = 0.0f;
    for (int t = tid; t < T; t += 256) {
        const float e = expf(s_sc[t] - mx);
        s_sc[t] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += s_red[i];
    const float inv = 1.0f / sum;
    float * o = out + ((size_t) (a0 + blockIdx.y) * R + r) * T;
    for (int t = tid; t < T; t += 256) o[t] = s_sc[t] * inv;
}

}  // namespace

void dtw_capture_layer(DType dt, const void * q, const DecRow * d_rows, int R, int d, const int * d_heads, int n_heads, size_t layer_off,
                       int T, int a0, float * out, cudaStream_t st) {
    if (R <= 0 || n_heads <= 0) return;
    const float kq_scale = powf(64.0f, -0.25f);
    const dim3 grid(R, n_heads);
    const size_t smem = (size_t) T * sizeof(float);
    if (dt == DType::F16)
        cross_align_kernel<__half><<<grid, 256, smem, st>>>(reinterpret_cast<const __half *>(q), d, d_rows, d_heads, layer_off, T, kq_scale, R, a0, out);
    else
        cross_align_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(reinterpret_cast<const __nv_bfloat16 *>(q), d, d_rows, d_heads, layer_off, T,
                                                                    kq_scale, R, a0, out);
    WB_CUDA(cudaGetLastError());
}

// ---- host: from head probabilities to a monotone token -> time path ---------------------------------------------------------------
// probs: [n_heads][n_tokens][T] f32 (T = audio context of the window), of which the first n_audio positions are used.
// skip_front tokens (the sot sequence) and the last token (eot) are left out of the alignment.  Returns, per remaining token
// (the first is <|notimestamps|>), the first audio position of its run on the DTW path, or -1 if the path never enters it.
std::vector<int> dtw_align(const float * probs, int n_heads, int n_tokens, int T, int n_audio, int skip_front, int medfilt_width) {
    const int N = n_tokens - skip_front - 1, M = n_audio;
    std::vector<int> first(std::max(N, 0), -1);
    if (N <= 0 || M <= 0 || n_heads <= 0) return first;
    // (1) per head and audio position: standardise over ALL tokens (ggml_norm, eps 1e-9: sums in double, result in float)
    //     -> z[a][i][j]
    std::vector<float> z((size_t) n_heads * n_tokens * M);
    for (int a = 0; a < n_heads; ++a)
        for (int j = 0; j < M; ++j) {
            const float * col = probs + ((size_t) a * n_tokens) * T + j;
            double sum = 0.0;
            for (int i = 0; i < n_tokens; ++i) sum += (double) col[(size_t) i * T];
            const float mean = (float) (sum / n_tokens);
            double sum2 = 0.0;
            for (int i = 0; i < n_tokens; ++i) {
                const float v = col[(size_t) i * T] - mean;
                z[((size_t) a * n_tokens + i) * M + j] = v;
                sum2 += (double) (v * v);
            }
            const float scale = 1.0f / sqrtf((float) (sum2 / n_tokens) + 1e-9f);
            for (int i = 0; i < n_tokens; ++i) z[((size_t) a * n_tokens + i) * M + j] *= scale;
        }
    // (2) median over a window of medfilt_width audio positions (reflected at both ends), (3) mean over the heads, negated:
    //     cost[i][j] of aligning token i with position j
    std::vector<float> cost((size_t) N * M);
    {
        const int hw = medfilt_width / 2;
        std::vector<float> win(medfilt_width), med((size_t) n_heads);
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < M; ++j) {
                for (int a = 0; a < n_heads; ++a) {
                    const float * zr = &z[((size_t) a * n_tokens + (i + skip_front)) * M];
                    for (int o = -hw; o <= hw; ++o) {
                        int idx = j + o;
                        if (idx < 0) idx = -idx;
                        else if (idx >= M) idx = 2 * (M - 1) - idx;
                        win[o + hw] = zr[idx];
                    }
                    std::sort(win.begin(), win.end());
                    med[a] = win[win.size() / 2];
                }
                double s = 0.0;
                for (int a = 0; a < n_heads; ++a) s += (double) med[a];
                cost[(size_t) i * M + j] = -((float) s / (float) n_heads);
            }
    }
    // (4) dynamic time warping: D[i][j] = cost + min(D[i-1][j-1], D[i-1][j], D[i][j-1]) with the reference's tie-breaking
    //     (diagonal only if strictly best, then "up" only if strictly best, else "left"), then the path from (N, M) back
    std::vector<float> D((size_t) (N + 1) * (M + 1), INFINITY);
    std::vector<signed char> step((size_t) (N + 1) * (M + 1), -1);
    auto at = [&](int i, int j) { return (size_t) i * (M + 1) + j; };
    D[at(0, 0)] = 0.0f;
    for (int j = 1; j <= M; ++j)
        for (int i = 1; i <= N; ++i) {
            const float c0 = D[at(i - 1, j - 1)], c1 = D[at(i - 1, j)], c2 = D[at(i, j - 1)];
            float c;
            signed char t;
            if (c0 < c1 && c0 < c2) { c = c0; t = 0; }
            else if (c1 < c0 && c1 < c2) { c = c1; t = 1; }
            else { c = c2; t = 2; }
            D[at(i, j)] = cost[(size_t) (i - 1) * M + (j - 1)] + c;
            step[at(i, j)] = t;
        }
    for (int j = 0; j <= M; ++j) step[at(0, j)] = 2;
    for (int i = 0; i <= N; ++i) step[at(i, 0)] = 1;
    // walking back from the end visits a token's run from its last position to its first: the last write per token wins
    for (int i = N, j = M; i > 0 || j > 0;) {
        if (i >= 1 && j >= 1) first[i - 1] = j - 1;
        else if (i >= 1) first[i - 1] = std::max(j - 1, -1);
        const signed char t = step[at(i, j)];
        if (t == 0) { --i; --j; }
        else if (t == 1) --i;
        else --j;
    }
    return first;
}

}  // namespace wb
