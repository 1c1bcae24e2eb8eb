// Fused log-mel spectrogram for sm_100a: reflect-pad, Hann window, 400-point real FFT at hop 160,
// power, sparse mel filterbank, log10, running global max -- one pass over the PCM, one write of the mel.
//
// Replaces the reference's CPU front-end: whisper_global_cache tables (src/whisper.cpp:2998-3033),
// dft/fft (3038-3102), log_mel_spectrogram_worker_thread (3104-3167) and the frame bookkeeping of
// log_mel_spectrogram (3170-3226).  The global-max clamp and (x+4)/4 normalisation (3228-3244) need the
// maximum over the whole call, so the kernel publishes the max (ordered-uint atomicMax) and the clamp is
// applied by the consumer (conv-stem im2col load, or mel_finalize below) -- no second pass over HBM.
//
// Mapping: 8 lanes per frame, 4 frames per warp, 32 frames per CTA (256 threads).
//   * PCM tile (5360 samples incl. 240 halo) -> shared memory with 16-byte coalesced loads.
//   * real FFT-400 as a complex FFT-200 of z[n] = x[2n] + i x[2n+1]:  200 = 8 x 25.
//       lane t holds z[t + 8m], m = 0..24 in registers, runs a 25-point FFT (5 x 5, radix-5 butterflies),
//       applies W200^(t*k2), then the 8-point DFT across the 8 lanes is three __shfl_xor butterfly stages.
//   * untangle to the 201 real-FFT bins (partner bin 200-k sits in lane^7), power -> shared memory.
//   * sparse filterbank (<= 2 non-zeros per FFT bin) with double accumulation as the reference,
//     log10, staged through shared memory so every mel row is written as 128-byte segments.
#include "mel.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace wb {

namespace {

constexpr int kFrame = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;
constexpr int kFramesPerCta = 32;
constexpr int kThreads = 256;
constexpr int kTileSamples = (kFramesPerCta - 1) * kHop + kFrame;          // 5360
constexpr int kTileHops = (kTileSamples + kHop - 1) / kHop;                // 34
constexpr int kPcmSmem = kTileSamples + 8 * kTileHops;                     // 8-float skew per hop
constexpr int kPowStride = 203;                                            // odd stride: conflict-light
constexpr int kOutStride = 36;                                             // == 4 (mod 32)

// Tables indexed per lane live in global memory and are staged to shared memory by every CTA; tables indexed
// with compile-time constants live in __constant__ (uniform access, folded into the FFMA operand).
struct MelTables {
    float hann[kFrame];     // periodic Hann, built on the host exactly as the reference does
    float2 tw200[8 * 25];   // W200^(t*k2), t-major
};
constexpr int kTabFloats = kFrame + 2 * 8 * 25;

__constant__ float2 c_w25[25];    // W25^k
__constant__ float2 c_w400[25];   // W400^k, k < 25
__constant__ float2 c_w16[8];     // W400^(25*k1) = W16^k1

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// forward 5-point DFT (e^{-i...}), in place on 5 registers
__device__ __forceinline__ void dft5(float2 & x0, float2 & x1, float2 & x2, float2 & x3, float2 & x4) {
    constexpr float c1 = 0.30901699437494742f;    // cos(2pi/5)
    constexpr float c2 = -0.80901699437494742f;   // cos(4pi/5)
    constexpr float s1 = 0.95105651629515357f;    // sin(2pi/5)
    constexpr float s2 = 0.58778525229247313f;    // sin(4pi/5)
    const float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
    const float2 m1 = make_float2(x0.x + c1 * t1.x + c2 * t2.x, x0.y + c1 * t1.y + c2 * t2.y);
    const float2 m2 = make_float2(x0.x + c2 * t1.x + c1 * t2.x, x0.y + c2 * t1.y + c1 * t2.y);
    const float2 u1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
    const float2 u2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
    x0 = make_float2(x0.x + t1.x + t2.x, x0.y + t1.y + t2.y);
    // y_k = m - i*u for k=1,2 ;  m + i*u for k=4,3      (-i*(a+ib) = b - ia)
    x1 = make_float2(m1.x + u1.y, m1.y - u1.x);
    x4 = make_float2(m1.x - u1.y, m1.y + u1.x);
    x2 = make_float2(m2.x + u2.y, m2.y - u2.x);
    x3 = make_float2(m2.x - u2.y, m2.y + u2.x);
}

__device__ __forceinline__ unsigned enc_ordered(float v) {
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <int I> struct IntC { static constexpr int value = I; };
template <int I, int N, typename F> __device__ __forceinline__ void static_for(F && f) {
    if constexpr (I < N) {
        f(IntC<I>{});
        static_for<I + 1, N>(f);
    }
}

// position (register index) that holds G[k2] after the in-register 25-point FFT:  k2 = ka + 5*kb  ->  5*ka + kb
__host__ __device__ constexpr int pos25(int k2) { return 5 * (k2 % 5) + (k2 / 5); }

__global__ void __launch_bounds__(kThreads, 3)
mel_kernel(const MelStream * __restrict__ streams, const float * __restrict__ tables,
           const float * __restrict__ filt, int filt_floats, int n_groups, int n_mel) {
    extern __shared__ __align__(16) float smem[];
    float * s_pcm = smem;                                   // kPcmSmem
    float * s_pow = s_pcm + kPcmSmem;                       // 32 * kPowStride
    float * s_tab = s_pow + kFramesPerCta * kPowStride;     // kTabFloats (hann | tw200)
    float * s_flt = s_tab + kTabFloats;                     // packed filterbank (see mel_plan_init)
    float * s_out = s_flt + filt_floats;                    // n_mel * kOutStride
    __shared__ float s_wmax[kThreads / 32];

    const MelStream st = streams[blockIdx.y];
    const int f0 = blockIdx.x * kFramesPerCta;
    if (f0 >= st.n_frames_fft) return;

    const int tid = threadIdx.x;
    const long long p0 = (long long) f0 * kHop;             // first padded-sample index of the tile
    const int n_pad = st.n_samples + kFrame / 2;            // samples beyond this are zero

    for (int i = tid; i < kTabFloats; i += kThreads) s_tab[i] = __ldg(&tables[i]);
    for (int i = tid; i < filt_floats; i += kThreads) s_flt[i] = __ldg(&filt[i]);
    const float * s_hann = s_tab;
    const float2 * s_tw200 = reinterpret_cast<const float2 *>(s_tab + kFrame);

    // ---- stage 1: PCM tile -> shared (padded[p] = p<200 ? x[200-p] : x[p-200], zero past the audio) ----
    {
        const float * __restrict__ x = st.pcm;
        const short * __restrict__ x16 = reinterpret_cast<const short *>(st.pcm);
        const bool i16 = st.pcm_i16 != 0;
        constexpr float kI16 = 1.0f / 32768.0f;                  // exact: a power of two
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(x) & (i16 ? 7 : 15)) == 0);
        for (int q4 = tid; q4 < kTileSamples / 4; q4 += kThreads) {
            const int q = q4 * 4;
            const long long p = p0 + q;
            float4 v;
            const long long s = p - kFrame / 2;
            if (vec_ok && s >= 0 && s + 3 < st.n_samples) {
                if (i16) {
                    const short4 h = __ldg(reinterpret_cast<const short4 *>(x16 + s));        // 8 bytes = four samples
                    v = make_float4((float) h.x * kI16, (float) h.y * kI16, (float) h.z * kI16, (float) h.w * kI16);
                } else {
                    v = __ldg(reinterpret_cast<const float4 *>(x + s));
                }
            } else {
                float e[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const long long pp = p + i;
                    long long idx = -1;
                    if (pp < kFrame / 2) {
                        const long long r = kFrame / 2 - pp;          // reflect (start only, as the reference)
                        if (r < st.n_samples) idx = r;
                    } else if (pp < n_pad) {
                        idx = pp - kFrame / 2;
                    }
                    e[i] = idx < 0 ? 0.0f : (i16 ? (float) x16[idx] * kI16 : x[idx]);
                }
                v = make_float4(e[0], e[1], e[2], e[3]);
            }
            *reinterpret_cast<float4 *>(s_pcm + q + 8 * (q / kHop)) = v;
        }
    }
    __syncthreads();

    const int lane = tid & 31;
    const int t = tid & 7;                  // lane within the frame group
    const int fl = tid >> 3;                // local frame 0..31
    const int frame = f0 + fl;
    const bool live = frame < st.n_frames_fft;

    // ---- stage 2: windowed load, z[t + 8m] ----
    float2 v[25];
    {
        const float * base = s_pcm + kHop * fl + 8 * fl + 2 * t;
#pragma unroll
        for (int m = 0; m < 25; ++m) {
            const float2 xs = *reinterpret_cast<const float2 *>(base + 16 * m + 8 * (m / 10));
            const float2 h = *reinterpret_cast<const float2 *>(&s_hann[2 * t + 16 * m]);
            v[m] = make_float2(xs.x * h.x, xs.y * h.y);
        }
    }

    // ---- stage 3: 25-point FFT in registers (m = 5*m1 + m2, k2 = ka + 5*kb) ----
#pragma unroll
    for (int m2 = 0; m2 < 5; ++m2) dft5(v[m2], v[5 + m2], v[10 + m2], v[15 + m2], v[20 + m2]);
    // now position 5*ka + m2 holds A[m2][ka]; twiddle by W25^(m2*ka)
    static_for<1, 5>([&](auto ka) {
        static_for<1, 5>([&](auto m2) {
            constexpr int KA = decltype(ka)::value, M2 = decltype(m2)::value;
            v[5 * KA + M2] = cmul(v[5 * KA + M2], c_w25[(KA * M2) % 25]);
        });
    });
#pragma unroll
    for (int ka = 0; ka < 5; ++ka) dft5(v[5 * ka], v[5 * ka + 1], v[5 * ka + 2], v[5 * ka + 3], v[5 * ka + 4]);
    // position 5*ka + kb holds G_t[ka + 5*kb]

    // ---- stage 4: W200^(t*k2), then 8-point DIF across the lanes of the frame group ----
    {
        const float2 * tw = &s_tw200[t * 25];
        static_for<1, 25>([&](auto k2) {
            constexpr int K2 = decltype(k2)::value;
            v[pos25(K2)] = cmul(v[pos25(K2)], tw[K2]);
        });
    }
    {
        // stage A (partner t^4): lower lanes a+b, upper lanes (b-a)*W8^(t&3).  "a+b or b-a" is one FMA with a per-lane
        // sign (b + sgn*a) instead of computing both and selecting.
        const bool up4 = (t & 4) != 0;
        const bool up2 = (t & 2) != 0;
        const bool up1 = (t & 1) != 0;
        const float sg4 = up4 ? -1.0f : 1.0f, sg2 = up2 ? -1.0f : 1.0f, sg1 = up1 ? -1.0f : 1.0f;
        float2 wA = make_float2(1.0f, 0.0f);
        if (up4) {
            constexpr float r = 0.70710678118654752440f;
            const int j = t & 3;
            wA = j == 0 ? make_float2(1.0f, 0.0f) : j == 1 ? make_float2(r, -r) : j == 2 ? make_float2(0.0f, -1.0f)
                                                                                         : make_float2(-r, -r);
        }
        const bool rotB = up2 && up1;      // stage B twiddle W4^(t&1) = -i on upper lanes with t&1
#pragma unroll
        for (int i = 0; i < 25; ++i) {
            float2 a = v[i];
            float2 b = make_float2(__shfl_xor_sync(0xffffffffu, a.x, 4), __shfl_xor_sync(0xffffffffu, a.y, 4));
            float2 d = make_float2(fmaf(sg4, a.x, b.x), fmaf(sg4, a.y, b.y));
            a = cmul(d, wA);
            b = make_float2(__shfl_xor_sync(0xffffffffu, a.x, 2), __shfl_xor_sync(0xffffffffu, a.y, 2));
            d = make_float2(fmaf(sg2, a.x, b.x), fmaf(sg2, a.y, b.y));
            a = rotB ? make_float2(d.y, -d.x) : d;            // * (-i)
            b = make_float2(__shfl_xor_sync(0xffffffffu, a.x, 1), __shfl_xor_sync(0xffffffffu, a.y, 1));
            v[i] = make_float2(fmaf(sg1, a.x, b.x), fmaf(sg1, a.y, b.y));
        }
    }
    // lane t now holds Z[k2 + 25*k1] at position pos25(k2), k1 = bitrev3(t)
    const int k1 = ((t & 1) << 2) | (t & 2) | ((t >> 2) & 1);

    // ---- stage 5: untangle to real-FFT bins, power -> shared ----
    {
        float * prow = s_pow + fl * kPowStride;
        const float2 w16 = c_w16[k1];
        // k2 = 0: partner is Z[25*((8-k1)&7)], register 0 of lane bitrev3((8-k1)&7)
        {
            const int kp = (8 - k1) & 7;
            const int src = (lane & ~7) | (((kp & 1) << 2) | (kp & 2) | ((kp >> 2) & 1));
            const float2 z = v[0];
            const float2 zp = make_float2(__shfl_sync(0xffffffffu, z.x, src), __shfl_sync(0xffffffffu, z.y, src));
            const float2 e = make_float2(0.5f * (z.x + zp.x), 0.5f * (z.y - zp.y));
            const float2 d = make_float2(z.x - zp.x, z.y + zp.y);          // Z - conj(Zp)
            const float2 o = make_float2(0.5f * d.y, -0.5f * d.x);        // -0.5i * d
            const float2 xk = cadd(e, cmul(w16, o));
            prow[25 * k1] = xk.x * xk.x + xk.y * xk.y;
            if (k1 == 0) {                                                 // bin 200: Re(Z0) - Im(Z0)
                const float n = z.x - z.y;
                prow[200] = n * n;
            }
        }
        static_for<1, 25>([&](auto k2) {
            constexpr int K2 = decltype(k2)::value;
            const float2 z = v[pos25(K2)];
            const float2 q = v[pos25(25 - K2)];
            const float2 zp = make_float2(__shfl_xor_sync(0xffffffffu, q.x, 7), __shfl_xor_sync(0xffffffffu, q.y, 7));
            const float2 e = make_float2(0.5f * (z.x + zp.x), 0.5f * (z.y - zp.y));
            const float2 d = make_float2(z.x - zp.x, z.y + zp.y);
            const float2 o = make_float2(0.5f * d.y, -0.5f * d.x);
            const float2 w = cmul(c_w400[K2], w16);               // W400^(k2 + 25*k1)
            const float2 xk = cadd(e, cmul(w, o));
            prow[K2 + 25 * k1] = xk.x * xk.x + xk.y * xk.y;
        });
    }
    __syncwarp();

    // ---- stage 6: sparse mel filterbank + log10; lane t takes bins t, t+8, ... ----
    // Packed table (shared memory): per group g of 8 bins {len_g, off_g}, per bin k_start, then the taps padded to the
    // group's longest filter and interleaved [tap][lane], so the trip count is warp-uniform and the weight reads are
    // conflict-free.  f32 FMA accumulation (the reference adds f32 4-term partials into a double, src/whisper.cpp:3140-3156;
    // the difference is ~1e-7 relative, i.e. ~1e-8 after log10 and /4).
    float vmax = -INFINITY;
    {
        const float * prow = s_pow + fl * kPowStride;
        const int * s_meta = reinterpret_cast<const int *>(s_flt);               // [2*n_groups] | [8*n_groups] k_start
        const int * s_k0 = s_meta + 2 * n_groups;
        const float * s_w = s_flt + 2 * n_groups + 8 * n_groups;
        for (int g = 0; g < n_groups; ++g) {
            const int len = s_meta[2 * g], off = s_meta[2 * g + 1];
            const int j = g * 8 + t;
            const float * pk = prow + s_k0[j];
            const float * wk = s_w + off + t;
            float acc = 0.0f;
            for (int i = 0; i < len; ++i) acc = fmaf(pk[i], wk[8 * i], acc);
            // log10(x) = log2(x) * log10(2); MUFU.LG2 has <= 2^-22 absolute error
            const float lg = __log2f(fmaxf(acc, 1e-10f)) * 0.30102999566398119521f;
            if (j < n_mel) {
                if (live) vmax = fmaxf(vmax, lg);
                s_out[j * kOutStride + fl] = lg;
            }
        }
    }
    vmax = warp_max(vmax);
    if (lane == 0) s_wmax[tid >> 5] = vmax;
    __syncthreads();

    // ---- stage 7: coalesced store of [n_mel][32 frames]; publish the max ----
    {
        const int warp = tid >> 5;
        const int fcol = f0 + lane;
        if (fcol < st.n_frames_fft) {
            for (int j = warp; j < n_mel; j += kThreads / 32) {
                st.out[(size_t) j * st.out_stride + fcol] = s_out[j * kOutStride + lane];
            }
        }
        if (tid == 0) {
            float m = s_wmax[0];
#pragma unroll
            for (int i = 1; i < kThreads / 32; ++i) m = fmaxf(m, s_wmax[i]);
            if (m > -INFINITY) atomicMax(st.max_enc, enc_ordered(m));
        }
    }
}

// Materialise the reference's final mel [n_mel][n_len] (clamp to max-8, (x+4)/4, -10 pad frames) from the raw
// log10 buffer: used by the parity hook and when a caller needs the reference's exact container.
__global__ void mel_finalize_kernel(const float * __restrict__ raw, int raw_stride, int n_frames_fft,
                                    const unsigned * __restrict__ max_enc, float * __restrict__ out, int n_len,
                                    int n_mel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= n_len || j >= n_mel) return;
    const float mmax = mel_decode_max(*max_enc);
    const float x = i < n_frames_fft ? raw[(size_t) j * raw_stride + i] : -10.0f;
    out[(size_t) j * n_len + i] = (fmaxf(x, mmax - 8.0f) + 4.0f) * 0.25f;
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------

MelPlan::~MelPlan() {
    if (d_tables) cudaFree(d_tables);
    if (d_w) cudaFree(d_w);
}

static void upload_tables(MelPlan & plan) {
    MelTables * tab = new MelTables();
    // Hann window exactly as the reference builds it: 0.5*(1 - cosf(2*pi*i/400)) with the angle in double
    // (src/whisper.cpp:3023-3032)
    for (int i = 0; i < kFrame; ++i) tab->hann[i] = (float) (0.5 * (1.0 - cosf((float) ((2.0 * M_PI * i) / kFrame))));
    for (int t = 0; t < 8; ++t)
        for (int k = 0; k < 25; ++k) {
            const double a = -2.0 * M_PI * (double) ((t * k) % 200) / 200.0;
            tab->tw200[t * 25 + k] = make_float2((float) cos(a), (float) sin(a));
        }
    WB_CUDA(cudaMalloc(&plan.d_tables, sizeof(MelTables)));
    WB_CUDA(cudaMemcpy(plan.d_tables, tab, sizeof(MelTables), cudaMemcpyHostToDevice));
    delete tab;
    float2 w25[25], w400[25], w16[8];
    for (int k = 0; k < 25; ++k) {
        const double a = -2.0 * M_PI * (double) k / 25.0, b = -2.0 * M_PI * (double) k / 400.0;
        w25[k] = make_float2((float) cos(a), (float) sin(a));
        w400[k] = make_float2((float) cos(b), (float) sin(b));
    }
    for (int k = 0; k < 8; ++k) {
        const double a = -2.0 * M_PI * (double) k / 16.0;
        w16[k] = make_float2((float) cos(a), (float) sin(a));
    }
    WB_CUDA(cudaMemcpyToSymbol(c_w25, w25, sizeof(w25)));
    WB_CUDA(cudaMemcpyToSymbol(c_w400, w400, sizeof(w400)));
    WB_CUDA(cudaMemcpyToSymbol(c_w16, w16, sizeof(w16)));
}

bool mel_plan_init(MelPlan & plan, const float * filters, int n_mel, int n_fft_bins) {
    if (n_fft_bins != kBins || n_mel <= 0 || n_mel > 256) return false;
    upload_tables(plan);
    const int n_groups = (n_mel + 7) / 8;
    std::vector<int> k0(8 * n_groups, 0), len(8 * n_groups, 0);
    for (int j = 0; j < n_mel; ++j) {
        const float * row = filters + (size_t) j * kBins;
        int lo = -1, hi = -1;
        for (int k = 0; k < kBins; ++k)
            if (row[k] != 0.0f) {
                if (lo < 0) lo = k;
                hi = k;
            }
        if (lo >= 0) {
            k0[j] = lo;
            len[j] = hi - lo + 1;
        }
    }
    std::vector<int> meta(2 * n_groups);
    std::vector<float> w;
    for (int g = 0; g < n_groups; ++g) {
        int glen = 0;
        for (int t = 0; t < 8; ++t) glen = std::max(glen, len[g * 8 + t]);
        meta[2 * g] = glen;
        meta[2 * g + 1] = (int) w.size();
        // every lane reads glen taps starting at its k_start: shift the start down so the window stays inside [0, 201)
        for (int t = 0; t < 8; ++t) k0[g * 8 + t] = std::min(k0[g * 8 + t], kBins - glen);
        for (int i = 0; i < glen; ++i)
            for (int t = 0; t < 8; ++t) {
                const int j = g * 8 + t;
                const int k = k0[j] + i;
                w.push_back(j < n_mel ? filters[(size_t) j * kBins + k] : 0.0f);
            }
    }
    std::vector<float> packed(2 * n_groups + 8 * n_groups + w.size());
    memcpy(packed.data(), meta.data(), meta.size() * 4);
    memcpy(packed.data() + 2 * n_groups, k0.data(), k0.size() * 4);
    memcpy(packed.data() + 10 * n_groups, w.data(), w.size() * 4);
    if (packed.size() * 4 > 64 * 1024) return false;      // a dense (non-triangular) filterbank is not supported
    plan.n_mel = n_mel;
    plan.n_groups = n_groups;
    plan.filt_floats = (int) ((packed.size() + 3) & ~size_t(3));
    packed.resize(plan.filt_floats, 0.0f);
    WB_CUDA(cudaMalloc(&plan.d_w, packed.size() * sizeof(float)));
    WB_CUDA(cudaMemcpy(plan.d_w, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice));
    plan.smem_bytes = (size_t) (kPcmSmem + kFramesPerCta * kPowStride + kTabFloats + plan.filt_floats + n_mel * kOutStride) * sizeof(float);
    WB_CUDA(cudaFuncSetAttribute(mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) plan.smem_bytes));
    return !cuda_failed();
}

MelGeometry mel_geometry(int n_samples) {
    MelGeometry g;
    // reference src/whisper.cpp:3206-3208 and the worker's loop bound at 3117
    g.n_len = (n_samples + 30 * 16000) / kHop;
    g.n_len_org = 1 + (n_samples + kFrame / 2 - kFrame) / kHop;
    const int n_fft = (n_samples + kFrame / 2) / kHop + 1;
    g.n_frames_fft = n_fft < g.n_len ? n_fft : g.n_len;
    g.stride = round_up(g.n_frames_fft, kFramesPerCta);
    return g;
}

void mel_launch(const MelPlan & plan, const MelStream * d_streams, int n_streams, int max_frames_fft,
                cudaStream_t stream) {
    if (n_streams <= 0 || max_frames_fft <= 0) return;
    dim3 grid(ceil_div(max_frames_fft, kFramesPerCta), n_streams);
    mel_kernel<<<grid, kThreads, plan.smem_bytes, stream>>>(d_streams, (const float *) plan.d_tables, (const float *) plan.d_w,
                                                             plan.filt_floats, plan.n_groups, plan.n_mel);
    WB_CUDA(cudaGetLastError());
}

void mel_finalize_launch(const float * raw, int raw_stride, int n_frames_fft, const unsigned * max_enc, float * out,
                         int n_len, int n_mel, cudaStream_t stream) {
    dim3 grid(ceil_div(n_len, 256), n_mel);
    mel_finalize_kernel<<<grid, 256, 0, stream>>>(raw, raw_stride, n_frames_fft, max_enc, out, n_len, n_mel);
    WB_CUDA(cudaGetLastError());
}

}  // namespace wb
