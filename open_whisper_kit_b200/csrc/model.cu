// GGML whisper model file -> device-resident weights in the layout the sm_100a kernels want.
//
// File format and validation rules follow the reference loader, whisper_model_load (src/whisper.cpp:1485-1956):
// magic, 11 hparams, mel filters, vocabulary (+ synthesised special tokens up to n_vocab), then tensor records
// {n_dims, name_len, ttype, ne[n_dims], name, data} until EOF, looked up by name (src/whisper-arch.h:42-109).
// Differences: no ggml contexts/buffers; F16/F32 payloads are converted once, on the device, to the 16-bit operand
// type of the tensor path; Q/K/V (and cross K/V) weights are concatenated so one GEMM produces them; conv weights
// are re-ordered to [out][k][in] so the conv stem runs as a GEMM over time-major activations.
#include "model.h"
#include "whisper_b200.h"

#include <math.h>
#include <stdarg.h>
#include <string.h>

namespace wb {

// ---- logging (reference: whisper_log_internal / g_state, src/whisper.cpp:954-960, 9000-9038) ---------------------
static void default_log(ggml_log_level, const char * text, void *) {
    fputs(text, stderr);
    fflush(stderr);
}
static ggml_log_callback g_log_cb = default_log;
static void * g_log_user = nullptr;

void wlog_set(ggml_log_callback cb, void * user) {
    g_log_cb = cb ? cb : default_log;
    g_log_user = user;
}
void wlog(ggml_log_level level, const char * fmt, ...) {
    if (!g_log_cb) return;
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_log_cb(level, buf, g_log_user);
}

// ---- languages (reference table g_lang, src/whisper.cpp:280-381; ids are the order of this list) --------------------
static const char * const kLang =
    "en:english,zh:chinese,de:german,es:spanish,ru:russian,ko:korean,fr:french,ja:japanese,pt:portuguese,tr:turkish,"
    "pl:polish,ca:catalan,nl:dutch,ar:arabic,sv:swedish,it:italian,id:indonesian,hi:hindi,fi:finnish,vi:vietnamese,"
    "he:hebrew,uk:ukrainian,el:greek,ms:malay,cs:czech,ro:romanian,da:danish,hu:hungarian,ta:tamil,no:norwegian,"
    "th:thai,ur:urdu,hr:croatian,bg:bulgarian,lt:lithuanian,la:latin,mi:maori,ml:malayalam,cy:welsh,sk:slovak,"
    "te:telugu,fa:persian,lv:latvian,bn:bengali,sr:serbian,az:azerbaijani,sl:slovenian,kn:kannada,et:estonian,"
    "mk:macedonian,br:breton,eu:basque,is:icelandic,hy:armenian,ne:nepali,mn:mongolian,bs:bosnian,kk:kazakh,"
    "sq:albanian,sw:swahili,gl:galician,mr:marathi,pa:punjabi,si:sinhala,km:khmer,sn:shona,yo:yoruba,so:somali,"
    "af:afrikaans,oc:occitan,ka:georgian,be:belarusian,tg:tajik,sd:sindhi,gu:gujarati,am:amharic,yi:yiddish,lo:lao,"
    "uz:uzbek,fo:faroese,ht:haitian creole,ps:pashto,tk:turkmen,nn:nynorsk,mt:maltese,sa:sanskrit,lb:luxembourgish,"
    "my:myanmar,bo:tibetan,tl:tagalog,mg:malagasy,as:assamese,tt:tatar,haw:hawaiian,ln:lingala,ha:hausa,ba:bashkir,"
    "jw:javanese,su:sundanese,yue:cantonese";

struct LangTable {
    std::vector<std::string> code, name;
    LangTable() {
        std::string s(kLang);
        size_t i = 0;
        while (i < s.size()) {
            size_t j = s.find(',', i);
            if (j == std::string::npos) j = s.size();
            const std::string item = s.substr(i, j - i);
            const size_t c = item.find(':');
            code.push_back(item.substr(0, c));
            name.push_back(item.substr(c + 1));
            i = j + 1;
        }
    }
};
static const LangTable & langs() {
    static LangTable t;
    return t;
}
int lang_max_id() { return (int) langs().code.size() - 1; }
const char * lang_str(int id) {
    if (id < 0 || id > lang_max_id()) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: unknown language id %d\n", __func__, id);
        return nullptr;
    }
    return langs().code[id].c_str();
}
const char * lang_str_full(int id) {
    if (id < 0 || id > lang_max_id()) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: unknown language id %d\n", __func__, id);
        return nullptr;
    }
    return langs().name[id].c_str();
}
int lang_id(const char * s) {
    if (!s) return -1;
    const auto & t = langs();
    for (size_t i = 0; i < t.code.size(); ++i)
        if (t.code[i] == s) return (int) i;
    for (size_t i = 0; i < t.name.size(); ++i)
        if (t.name[i] == s) return (int) i;
    wlog(GGML_LOG_LEVEL_ERROR, "%s: unknown language '%s'\n", __func__, s);
    return -1;
}

// ---- device conversion kernels -------------------------------------------------------------------------------
namespace {

template <typename S> __device__ __forceinline__ float load_f(const S * p, size_t i);
template <> __device__ __forceinline__ float load_f<float>(const float * p, size_t i) { return p[i]; }
template <> __device__ __forceinline__ float load_f<__half>(const __half * p, size_t i) { return __half2float(p[i]); }

// dst[r][c] (leading dim ld) = src[r][c]
template <typename S, typename T16>
__global__ void cvt_rows_kernel(const S * __restrict__ src, T16 * __restrict__ dst, int rows, int cols, int ld) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t) rows * cols) return;
    const int r = (int) (i / cols), c = (int) (i % cols);
    dst[(size_t) r * ld + c] = Half16<T16>::from_f(load_f<S>(src, i));
}
template <typename S> __global__ void cvt_f32_kernel(const S * __restrict__ src, float * __restrict__ dst, size_t n) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = load_f<S>(src, i);
}
// conv weight [out][in][3] -> [out][kpad] with K index k*in + c, zero padded
template <typename S, typename T16>
__global__ void cvt_conv_kernel(const S * __restrict__ src, T16 * __restrict__ dst, int n_out, int n_in, int kpad) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t) n_out * kpad) return;
    const int o = (int) (i / kpad), kk = (int) (i % kpad);
    float v = 0.0f;
    if (kk < 3 * n_in) {
        const int k = kk / n_in, c = kk % n_in;
        v = load_f<S>(src, ((size_t) o * n_in + c) * 3 + k);
    }
    dst[i] = Half16<T16>::from_f(v);
}

struct HostTensor {
    int ttype = 0;   // 0 f32, 1 f16
    int n_dims = 0;
    int ne[4] = {1, 1, 1, 1};
    std::vector<char> data;
    size_t nelem() const { return (size_t) ne[0] * ne[1] * ne[2] * ne[3]; }
};

// c[n] = sum_k gamma[k] W[n][k], b'[n] = bias[n] + sum_k beta[k] W[n][k] from the 16-bit weights as the GEMM sees them: one warp
// per output column, double accumulation (runs once per model load).
template <typename T16>
__global__ void ln_fold_kernel(const T16 * __restrict__ w, const float * __restrict__ gamma, const float * __restrict__ beta,
                               const float * __restrict__ bias, int N, int K, float * __restrict__ c_out, float * __restrict__ b_out) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    double c = 0.0, b = 0.0;
    for (int k = lane; k < K; k += 32) {
        const double wv = (double) Half16<T16>::to_f(w[(size_t) n * K + k]);
        c += wv * (double) gamma[k];
        b += wv * (double) beta[k];
    }
    for (int o = 16; o > 0; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
        c_out[n] = (float) c;
        b_out[n] = (float) (b + (bias ? (double) bias[n] : 0.0));
    }
}

struct Uploader {
    Model & m;
    void * d_tmp = nullptr;
    size_t tmp_cap = 0;
    bool ok = true;

    explicit Uploader(Model & mm) : m(mm) {}
    ~Uploader() {
        if (d_tmp) cudaFree(d_tmp);
    }
    void * dalloc(size_t bytes) {
        void * p = nullptr;
        bytes = round_up<size_t>(bytes ? bytes : 16, 256);
        WB_CUDA(cudaMalloc(&p, bytes));
        if (!p) {
            ok = false;
            return nullptr;
        }
        WB_CUDA(cudaMemset(p, 0, bytes));
        m.allocs.push_back(p);
        m.bytes_device += bytes;
        return p;
    }
    const void * stage(const HostTensor & t) {
        if (t.data.size() > tmp_cap) {
            if (d_tmp) cudaFree(d_tmp);
            tmp_cap = round_up<size_t>(t.data.size(), 1 << 20);
            WB_CUDA(cudaMalloc(&d_tmp, tmp_cap));
        }
        WB_CUDA(cudaMemcpy(d_tmp, t.data.data(), t.data.size(), cudaMemcpyHostToDevice));
        return d_tmp;
    }
    // 2-D weight [rows][cols] into dst (16-bit, leading dim ld) at row offset row0
    void put_rows(const HostTensor * t, void * dst, int row0, int rows, int cols, int ld) {
        if (!t) return;
        const void * s = stage(*t);
        const size_t n = (size_t) rows * cols;
        const int th = 256;
        const unsigned bl = (unsigned) ceil_div<size_t>(n, th);
        if (m.dtype == DType::F16) {
            __half * d = reinterpret_cast<__half *>(dst) + (size_t) row0 * ld;
            if (t->ttype == 1) cvt_rows_kernel<__half, __half><<<bl, th>>>((const __half *) s, d, rows, cols, ld);
            else cvt_rows_kernel<float, __half><<<bl, th>>>((const float *) s, d, rows, cols, ld);
        } else {
            __nv_bfloat16 * d = reinterpret_cast<__nv_bfloat16 *>(dst) + (size_t) row0 * ld;
            if (t->ttype == 1) cvt_rows_kernel<__half, __nv_bfloat16><<<bl, th>>>((const __half *) s, d, rows, cols, ld);
            else cvt_rows_kernel<float, __nv_bfloat16><<<bl, th>>>((const float *) s, d, rows, cols, ld);
        }
        WB_CUDA(cudaGetLastError());
        WB_CUDA(cudaDeviceSynchronize());
    }
    void ln_fold(const void * w, const float * gamma, const float * beta, const float * bias, int N, int K, float *& c, float *& b) {
        c = (float *) dalloc((size_t) N * 4);
        b = (float *) dalloc((size_t) N * 4);
        if (!w || !gamma || !beta || !c || !b) return;
        const unsigned bl = (unsigned) ceil_div(N, 8);
        if (m.dtype == DType::F16) ln_fold_kernel<__half><<<bl, 256>>>((const __half *) w, gamma, beta, bias, N, K, c, b);
        else ln_fold_kernel<__nv_bfloat16><<<bl, 256>>>((const __nv_bfloat16 *) w, gamma, beta, bias, N, K, c, b);
        WB_CUDA(cudaGetLastError());
    }
    void put_conv(const HostTensor * t, void * dst, int n_out, int n_in, int kpad) {
        if (!t) return;
        const void * s = stage(*t);
        const size_t n = (size_t) n_out * kpad;
        const int th = 256;
        const unsigned bl = (unsigned) ceil_div<size_t>(n, th);
        if (m.dtype == DType::F16) {
            if (t->ttype == 1) cvt_conv_kernel<__half, __half><<<bl, th>>>((const __half *) s, (__half *) dst, n_out, n_in, kpad);
            else cvt_conv_kernel<float, __half><<<bl, th>>>((const float *) s, (__half *) dst, n_out, n_in, kpad);
        } else {
            if (t->ttype == 1)
                cvt_conv_kernel<__half, __nv_bfloat16><<<bl, th>>>((const __half *) s, (__nv_bfloat16 *) dst, n_out, n_in, kpad);
            else
                cvt_conv_kernel<float, __nv_bfloat16><<<bl, th>>>((const float *) s, (__nv_bfloat16 *) dst, n_out, n_in, kpad);
        }
        WB_CUDA(cudaGetLastError());
        WB_CUDA(cudaDeviceSynchronize());
    }
    void put_f32(const HostTensor * t, float * dst, size_t off, size_t n) {
        if (!t) return;
        const void * s = stage(*t);
        const int th = 256;
        const unsigned bl = (unsigned) ceil_div<size_t>(n, th);
        if (t->ttype == 1) cvt_f32_kernel<__half><<<bl, th>>>((const __half *) s, dst + off, n);
        else cvt_f32_kernel<float><<<bl, th>>>((const float *) s, dst + off, n);
        WB_CUDA(cudaGetLastError());
        WB_CUDA(cudaDeviceSynchronize());
    }
};

// ---- ggml 32-element block formats -> f16 (host, at load time) ------------------------------------------------
// Layouts and arithmetic restate dequantize_row_q4_0 / q4_1 / q5_0 / q5_1 / q8_0 (reference ggml/src/ggml-quants.c:307-415,
// structs ggml/src/ggml-common.h): value = q * d (+ m) in f32, then rounded once to f16 (the type our GEMMs read).
int quant_block_bytes(int ttype) {
    switch (ttype) {
        case 2: return 18;   // q4_0: f16 d, 16 x 2 nibbles
        case 3: return 20;   // q4_1: f16 d, f16 m, 16 x 2 nibbles
        case 6: return 22;   // q5_0: f16 d, u32 fifth bits, 16 x 2 nibbles
        case 7: return 24;   // q5_1: f16 d, f16 m, u32 fifth bits, 16 x 2 nibbles
        case 8: return 34;   // q8_0: f16 d, 32 x i8
        default: return 0;
    }
}
// 256-element super-block formats ("K-quants", reference ggml/src/ggml-common.h:262-336)
int kquant_block_bytes(int ttype) {
    switch (ttype) {
        case 10: return 84;    // q2_K: scales[16], qs[64], f16 d, f16 dmin
        case 11: return 110;   // q3_K: hmask[32], qs[64], scales[12], f16 d
        case 12: return 144;   // q4_K: f16 d, f16 dmin, scales[12], qs[128]
        case 13: return 176;   // q5_K: f16 d, f16 dmin, scales[12], qh[32], qs[128]
        case 14: return 210;   // q6_K: ql[128], qh[64], i8 scales[16], f16 d
        default: return 0;
    }
}
// 6-bit scale / min pairs of q4_K and q5_K (get_scale_min_k4, ggml/src/ggml-quants.c:703-710)
static inline void scale_min_k4(int j, const unsigned char * q, int & sc, int & m) {
    if (j < 4) {
        sc = q[j] & 63;
        m = q[j + 4] & 63;
    } else {
        sc = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4);
        m = (q[j + 4] >> 4) | ((q[j] >> 6) << 4);
    }
}
// Restates dequantize_row_q2_K / q3_K / q4_K / q5_K / q6_K (ggml/src/ggml-quants.c:784-814, 1128-1176, 1352-1374, 1554-1579,
// 1762-1791): same f32 expressions in the same order, then one rounding to f16.
void dequantize_kblocks(int ttype, const unsigned char * raw, size_t n_blocks, __half * out) {
    const int bb = kquant_block_bytes(ttype);
    auto h2f = [](const unsigned char * p) {
        __half h;
        memcpy(&h, p, 2);
        return __half2float(h);
    };
    for (size_t b = 0; b < n_blocks; ++b) {
        const unsigned char * p = raw + b * bb;
        __half * y = out + b * 256;
        if (ttype == 10) {
            const unsigned char * scales = p, * q = p + 16;
            const float d = h2f(p + 80), mn = h2f(p + 82);
            int is = 0;
            for (int n = 0; n < 256; n += 128) {
                int shift = 0;
                for (int j = 0; j < 4; ++j) {
                    for (int half = 0; half < 2; ++half) {
                        const unsigned char sc = scales[is++];
                        const float dl = d * (sc & 0xF), ml = mn * (sc >> 4);
                        for (int l = 0; l < 16; ++l) *y++ = __float2half_rn(dl * (float) ((signed char) ((q[l + 16 * half] >> shift) & 3)) - ml);
                    }
                    shift += 2;
                }
                q += 32;
            }
        } else if (ttype == 11) {
            const unsigned char * hm = p, * q = p + 32;
            const float d_all = h2f(p + 108);
            uint32_t aux[4];
            memcpy(aux, p + 96, 12);
            const uint32_t kmask1 = 0x03030303, kmask2 = 0x0f0f0f0f, tmp = aux[2];
            aux[2] = ((aux[0] >> 4) & kmask2) | (((tmp >> 4) & kmask1) << 4);
            aux[3] = ((aux[1] >> 4) & kmask2) | (((tmp >> 6) & kmask1) << 4);
            aux[0] = (aux[0] & kmask2) | (((tmp >> 0) & kmask1) << 4);
            aux[1] = (aux[1] & kmask2) | (((tmp >> 2) & kmask1) << 4);
            const signed char * scales = reinterpret_cast<const signed char *>(aux);
            int is = 0;
            unsigned char m = 1;
            for (int n = 0; n < 256; n += 128) {
                int shift = 0;
                for (int j = 0; j < 4; ++j) {
                    for (int half = 0; half < 2; ++half) {
                        const float dl = d_all * (scales[is++] - 32);
                        for (int l = 0; l < 16; ++l) {
                            const int qv = (signed char) ((q[l + 16 * half] >> shift) & 3) - ((hm[l + 16 * half] & m) ? 0 : 4);
                            *y++ = __float2half_rn(dl * (float) qv);
                        }
                    }
                    shift += 2;
                    m <<= 1;
                }
                q += 32;
            }
        } else if (ttype == 12 || ttype == 13) {
            const float d = h2f(p), mn = h2f(p + 2);
            const unsigned char * scales = p + 4;
            const unsigned char * qh = p + 16;                       // q5_K only
            const unsigned char * q = ttype == 12 ? p + 16 : p + 48;
            int is = 0;
            unsigned char u1 = 1, u2 = 2;
            for (int j = 0; j < 256; j += 64) {
                int sc, m;
                scale_min_k4(is + 0, scales, sc, m);
                const float d1 = d * sc, m1 = mn * m;
                scale_min_k4(is + 1, scales, sc, m);
                const float d2 = d * sc, m2 = mn * m;
                for (int l = 0; l < 32; ++l) {
                    const int v = (q[l] & 0xF) + ((ttype == 13 && (qh[l] & u1)) ? 16 : 0);
                    *y++ = __float2half_rn(d1 * (float) v - m1);
                }
                for (int l = 0; l < 32; ++l) {
                    const int v = (q[l] >> 4) + ((ttype == 13 && (qh[l] & u2)) ? 16 : 0);
                    *y++ = __float2half_rn(d2 * (float) v - m2);
                }
                q += 32;
                is += 2;
                u1 <<= 2;
                u2 <<= 2;
            }
        } else {      // 14: q6_K
            const unsigned char * ql = p, * qh = p + 128;
            const signed char * sc = reinterpret_cast<const signed char *>(p + 192);
            const float d = h2f(p + 208);
            for (int n = 0; n < 256; n += 128) {
                for (int l = 0; l < 32; ++l) {
                    const int is = l / 16;
                    const signed char q1 = (signed char) ((ql[l + 0] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
                    const signed char q2 = (signed char) ((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
                    const signed char q3 = (signed char) ((ql[l + 0] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
                    const signed char q4 = (signed char) ((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
                    y[l + 0] = __float2half_rn(d * sc[is + 0] * q1);
                    y[l + 32] = __float2half_rn(d * sc[is + 2] * q2);
                    y[l + 64] = __float2half_rn(d * sc[is + 4] * q3);
                    y[l + 96] = __float2half_rn(d * sc[is + 6] * q4);
                }
                y += 128;
                ql += 64;
                qh += 32;
                sc += 8;
            }
        }
    }
}
void dequantize_blocks(int ttype, const unsigned char * raw, size_t n_blocks, __half * out) {
    const int bb = quant_block_bytes(ttype);
    auto h2f = [](const unsigned char * p) {
        __half h;
        memcpy(&h, p, 2);
        return __half2float(h);
    };
    for (size_t b = 0; b < n_blocks; ++b) {
        const unsigned char * p = raw + b * bb;
        __half * y = out + b * 32;
        const float d = h2f(p);
        if (ttype == 8) {
            const signed char * q = reinterpret_cast<const signed char *>(p + 2);
            for (int j = 0; j < 32; ++j) y[j] = __float2half_rn((float) q[j] * d);
            continue;
        }
        const bool has_m = ttype == 3 || ttype == 7, has_h = ttype == 6 || ttype == 7;
        const float m = has_m ? h2f(p + 2) : 0.0f;
        const unsigned char * q = p + 2 + (has_m ? 2 : 0);
        uint32_t qh = 0;
        if (has_h) {
            memcpy(&qh, q, 4);
            q += 4;
        }
        for (int j = 0; j < 16; ++j) {
            int x0 = q[j] & 0x0F, x1 = q[j] >> 4;
            if (has_h) {
                x0 |= ((qh >> j) << 4) & 0x10;
                x1 |= (qh >> (j + 12)) & 0x10;
            }
            if (!has_m) {               // symmetric formats are offset by half their range
                x0 -= has_h ? 16 : 8;
                x1 -= has_h ? 16 : 8;
            }
            y[j] = __float2half_rn(has_m ? (float) x0 * d + m : (float) x0 * d);
            y[j + 16] = __float2half_rn(has_m ? (float) x1 * d + m : (float) x1 * d);
        }
    }
}

template <typename T> bool read_pod(whisper_model_loader * l, T & v) { return l->read(l->context, &v, sizeof(T)) == sizeof(T); }

}  // namespace

Model::~Model() {
    for (void * p : allocs) cudaFree(p);
}

bool model_load(whisper_model_loader * loader, Model & m, DType dtype, int device) {
    m.dtype = dtype;
    m.device = device;
    auto & hp = m.hp;
    auto & vocab = m.vocab;

    uint32_t magic = 0;
    if (!read_pod(loader, magic) || magic != 0x67676d6c) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: invalid model data (bad magic)\n", __func__);
        return false;
    }
    int32_t h[11];
    for (int i = 0; i < 11; ++i)
        if (!read_pod(loader, h[i])) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: truncated header\n", __func__);
            return false;
        }
    hp.n_vocab = h[0]; hp.n_audio_ctx = h[1]; hp.n_audio_state = h[2]; hp.n_audio_head = h[3]; hp.n_audio_layer = h[4];
    hp.n_text_ctx = h[5]; hp.n_text_state = h[6]; hp.n_text_head = h[7]; hp.n_text_layer = h[8]; hp.n_mels = h[9];
    hp.ftype = h[10];
    switch (hp.n_audio_layer) {
        case 4: m.type = 1; break;
        case 6: m.type = 2; break;
        case 12: m.type = 3; break;
        case 24: m.type = 4; break;
        case 32: m.type = 5; break;
        default: m.type = 0;
    }
    hp.ftype %= 1000;   // GGML_QNT_VERSION_FACTOR
    // 0 f32, 1 f16, 2 q4_0, 3 q4_1, 7 q8_0, 8 q5_0, 9 q5_1 (ggml_ftype, ggml/include/ggml.h); the 32-element block formats are
    // and 10..14 q2_K..q6_K are expanded to 16-bit weights while loading (record by record, whatever its own type says).
    // IQ / MXFP4 formats are not.
    if (hp.ftype != 0 && hp.ftype != 1 && hp.ftype != 2 && hp.ftype != 3 && hp.ftype != 7 && hp.ftype != 8 && hp.ftype != 9 &&
        !(hp.ftype >= 10 && hp.ftype <= 14)) {
        wlog(GGML_LOG_LEVEL_ERROR,
             "%s: model file ftype %d is not supported by the B200 path (supported: f32, f16, q4_0, q4_1, q5_0, q5_1, q8_0, "
             "q2_K ... q6_K)\n",
             __func__, hp.ftype);
        return false;
    }
    const int d = hp.n_audio_state;
    if (hp.n_vocab > 53248 /* greedy selection kernel: two CTAs x 26624 logits per row */ || hp.n_text_state != d || d <= 0 ||
        d % 64 != 0 || d > 1280 || hp.n_audio_head * 64 != d ||
        hp.n_text_head * 64 != d || hp.n_audio_ctx != 1500 || hp.n_text_ctx <= 0 || hp.n_mels <= 0 || hp.n_vocab <= 0) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: unsupported geometry (d=%d, heads=%d/%d, audio_ctx=%d)\n", __func__, d,
             hp.n_audio_head, hp.n_text_head, hp.n_audio_ctx);
        return false;
    }
    wlog(GGML_LOG_LEVEL_INFO, "%s: n_vocab=%d n_audio_state=%d n_audio_layer=%d n_text_layer=%d n_mels=%d ftype=%d\n",
         __func__, hp.n_vocab, d, hp.n_audio_layer, hp.n_text_layer, hp.n_mels, hp.ftype);

    // mel filters
    if (!read_pod(loader, m.filt_n_mel) || !read_pod(loader, m.filt_n_fft) || m.filt_n_mel <= 0 || m.filt_n_fft <= 0 ||
        (size_t) m.filt_n_mel * m.filt_n_fft > (1u << 20)) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: bad mel filter header\n", __func__);
        return false;
    }
    m.filters.resize((size_t) m.filt_n_mel * m.filt_n_fft);
    if (loader->read(loader->context, m.filters.data(), m.filters.size() * sizeof(float)) != m.filters.size() * sizeof(float)) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: truncated mel filters\n", __func__);
        return false;
    }

    // vocabulary
    {
        int32_t n_vocab = 0;
        if (!read_pod(loader, n_vocab) || n_vocab < 0 || n_vocab > (1 << 20)) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: bad vocabulary size\n", __func__);
            return false;
        }
        vocab.id_to_token.assign(std::max(n_vocab, hp.n_vocab), std::string());
        std::vector<char> tmp;
        for (int i = 0; i < n_vocab; ++i) {
            uint32_t len = 0;
            if (!read_pod(loader, len) || len > (1u << 16)) {
                wlog(GGML_LOG_LEVEL_ERROR, "%s: bad vocabulary entry %d\n", __func__, i);
                return false;
            }
            std::string word;
            if (len > 0) {
                tmp.resize(len);
                if (loader->read(loader->context, tmp.data(), len) != len) return false;
                word.assign(tmp.data(), len);
            }
            vocab.token_to_id[word] = i;
            vocab.id_to_token[i] = word;
        }
        vocab.n_vocab = hp.n_vocab;
        if (vocab.is_multilingual()) {
            vocab.token_eot++;
            vocab.token_sot++;
            const int dt = vocab.num_languages() - 98;
            vocab.token_translate += dt;
            vocab.token_transcribe += dt;
            vocab.token_solm += dt;
            vocab.token_prev += dt;
            vocab.token_nosp += dt;
            vocab.token_not += dt;
            vocab.token_beg += dt;
        }
        for (int i = n_vocab; i < hp.n_vocab; ++i) {
            std::string word;
            if (i > vocab.token_beg) word = "[_TT_" + std::to_string(i - vocab.token_beg) + "]";
            else if (i == vocab.token_eot) word = "[_EOT_]";
            else if (i == vocab.token_sot) word = "[_SOT_]";
            else if (i == vocab.token_translate) word = "[_TRANSLATE_]";
            else if (i == vocab.token_transcribe) word = "[_TRANSCRIBE_]";
            else if (i == vocab.token_solm) word = "[_SOLM_]";
            else if (i == vocab.token_prev) word = "[_PREV_]";
            else if (i == vocab.token_nosp) word = "[_NOSP_]";
            else if (i == vocab.token_not) word = "[_NOT_]";
            else if (i == vocab.token_beg) word = "[_BEG_]";
            else if (i > vocab.token_sot && i <= vocab.token_sot + vocab.num_languages()) {
                const char * ls = (i - vocab.token_sot - 1) <= lang_max_id() ? lang_str(i - vocab.token_sot - 1) : nullptr;
                word = "[_LANG_" + std::string(ls ? ls : "?") + "]";
            } else word = "[_extra_token_" + std::to_string(i) + "]";
            vocab.token_to_id[word] = i;
            vocab.id_to_token[i] = word;
        }
    }

    // tensor records
    std::map<std::string, HostTensor> tensors;
    int n_quantised = 0;
    while (true) {
        int32_t n_dims = 0, length = 0, ttype = 0;
        if (!read_pod(loader, n_dims)) break;   // clean EOF
        if (!read_pod(loader, length) || !read_pod(loader, ttype)) break;
        const int kblock = kquant_block_bytes(ttype);     // 256-element super-blocks
        const int qblock = kblock ? kblock : quant_block_bytes(ttype);      // 0: not a block format
        const int qelems = kblock ? 256 : 32;
        if (n_dims < 1 || n_dims > 4 || length <= 0 || length > 256 || (ttype != 0 && ttype != 1 && qblock == 0)) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: malformed or unsupported tensor record (n_dims=%d, name_len=%d, type=%d)\n", __func__,
                 n_dims, length, ttype);
            return false;
        }
        HostTensor t;
        t.ttype = ttype;
        t.n_dims = n_dims;
        for (int i = 0; i < n_dims; ++i)
            if (!read_pod(loader, t.ne[i]) || t.ne[i] <= 0) return false;
        std::string name(length, '\0');
        if (loader->read(loader->context, &name[0], length) != (size_t) length) return false;
        if (qblock) {
            // quantised record: rows of ne[0] elements in blocks of 32 (reference ggml/src/ggml-quants.c:307-415); expand to f16
            if (t.ne[0] % qelems) {
                wlog(GGML_LOG_LEVEL_ERROR, "%s: tensor '%s': quantised row length %d is not a multiple of %d\n", __func__,
                     name.c_str(), t.ne[0], qelems);
                return false;
            }
            const size_t n_blocks = t.nelem() / qelems;
            std::vector<unsigned char> raw(n_blocks * qblock);
            if (loader->read(loader->context, raw.data(), raw.size()) != raw.size()) {
                wlog(GGML_LOG_LEVEL_ERROR, "%s: tensor '%s' is truncated\n", __func__, name.c_str());
                return false;
            }
            t.data.resize(t.nelem() * 2);
            if (kblock) dequantize_kblocks(ttype, raw.data(), n_blocks, reinterpret_cast<__half *>(t.data.data()));
            else dequantize_blocks(ttype, raw.data(), n_blocks, reinterpret_cast<__half *>(t.data.data()));
            t.ttype = 1;
            ++n_quantised;
            tensors[name] = std::move(t);
            continue;
        }
        const size_t bytes = t.nelem() * (ttype == 1 ? 2 : 4);
        t.data.resize(bytes);
        if (loader->read(loader->context, t.data.data(), bytes) != bytes) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: tensor '%s' is truncated\n", __func__, name.c_str());
            return false;
        }
        tensors[name] = std::move(t);
    }

    const int n_mels = hp.n_mels, n_vocab = hp.n_vocab, n_al = hp.n_audio_layer, n_tl = hp.n_text_layer;
    // expected directory: name -> ne[0..2] (ggml order, fastest first); reference src/whisper.cpp:1758-1842
    struct Want { int ne0, ne1, ne2; };
    std::map<std::string, Want> want;
    auto W = [&](const std::string & n, int a, int b = 1, int c = 1) { want[n] = {a, b, c}; };
    W("encoder.positional_embedding", d, hp.n_audio_ctx);
    W("encoder.conv1.weight", 3, n_mels, d);
    W("encoder.conv1.bias", 1, d);
    W("encoder.conv2.weight", 3, d, d);
    W("encoder.conv2.bias", 1, d);
    W("encoder.ln_post.weight", d);
    W("encoder.ln_post.bias", d);
    W("decoder.positional_embedding", d, hp.n_text_ctx);
    W("decoder.token_embedding.weight", d, n_vocab);
    W("decoder.ln.weight", d);
    W("decoder.ln.bias", d);
    auto block = [&](const std::string & p, bool cross) {
        W(p + ".attn_ln.weight", d); W(p + ".attn_ln.bias", d);
        W(p + ".attn.query.weight", d, d); W(p + ".attn.query.bias", d);
        W(p + ".attn.key.weight", d, d);
        W(p + ".attn.value.weight", d, d); W(p + ".attn.value.bias", d);
        W(p + ".attn.out.weight", d, d); W(p + ".attn.out.bias", d);
        if (cross) {
            W(p + ".cross_attn_ln.weight", d); W(p + ".cross_attn_ln.bias", d);
            W(p + ".cross_attn.query.weight", d, d); W(p + ".cross_attn.query.bias", d);
            W(p + ".cross_attn.key.weight", d, d);
            W(p + ".cross_attn.value.weight", d, d); W(p + ".cross_attn.value.bias", d);
            W(p + ".cross_attn.out.weight", d, d); W(p + ".cross_attn.out.bias", d);
        }
        W(p + ".mlp_ln.weight", d); W(p + ".mlp_ln.bias", d);
        W(p + ".mlp.0.weight", d, 4 * d); W(p + ".mlp.0.bias", 4 * d);
        W(p + ".mlp.2.weight", 4 * d, d); W(p + ".mlp.2.bias", d);
    };
    for (int i = 0; i < n_al; ++i) block("encoder.blocks." + std::to_string(i), false);
    for (int i = 0; i < n_tl; ++i) block("decoder.blocks." + std::to_string(i), true);

    for (const auto & kv : tensors) {
        auto it = want.find(kv.first);
        if (it == want.end()) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: unknown tensor '%s' in model file\n", __func__, kv.first.c_str());
            return false;
        }
        const HostTensor & t = kv.second;
        if (t.ne[0] != it->second.ne0 || t.ne[1] != it->second.ne1 || t.ne[2] != it->second.ne2) {
            wlog(GGML_LOG_LEVEL_ERROR, "%s: tensor '%s' has wrong shape in model file: got [%d, %d, %d], expected [%d, %d, %d]\n",
                 __func__, kv.first.c_str(), t.ne[0], t.ne[1], t.ne[2], it->second.ne0, it->second.ne1, it->second.ne2);
            return false;
        }
    }
    m.n_loaded = (int) tensors.size();
    if (m.n_loaded == 0) {
        wlog(GGML_LOG_LEVEL_WARN, "%s: WARN no tensors loaded from model file - assuming empty model for testing\n", __func__);
    } else if (m.n_loaded != (int) want.size()) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: ERROR not all tensors loaded from model file - expected %zu, got %d\n", __func__,
             want.size(), m.n_loaded);
        return false;
    }

    // ---- device placement ----
    cuda_clear_failure();
    WB_CUDA(cudaSetDevice(device));
    if (cuda_failed()) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: no usable CUDA device %d (this library has no CPU path)\n", __func__, device);
        return false;
    }
    Uploader up(m);
    auto T = [&](const std::string & n) -> const HostTensor * {
        auto it = tensors.find(n);
        return it == tensors.end() ? nullptr : &it->second;
    };
    auto f32 = [&](const std::string & n, size_t cnt) -> float * {
        float * p = (float *) up.dalloc(cnt * 4);
        if (p) up.put_f32(T(n), p, 0, cnt);
        return p;
    };
    auto w16 = [&](const std::string & n, int rows, int cols) -> void * {
        void * p = up.dalloc((size_t) rows * cols * 2);
        if (p) up.put_rows(T(n), p, 0, rows, cols, cols);
        return p;
    };

    m.e_pe = f32("encoder.positional_embedding", (size_t) hp.n_audio_ctx * d);
    m.conv1_kpad = round_up(3 * n_mels, 64);
    m.conv1_w = up.dalloc((size_t) d * m.conv1_kpad * 2);
    up.put_conv(T("encoder.conv1.weight"), m.conv1_w, d, n_mels, m.conv1_kpad);
    m.conv1_b = f32("encoder.conv1.bias", d);
    m.conv2_w = up.dalloc((size_t) d * 3 * d * 2);
    up.put_conv(T("encoder.conv2.weight"), m.conv2_w, d, d, 3 * d);
    m.conv2_b = f32("encoder.conv2.bias", d);
    m.e_ln_w = f32("encoder.ln_post.weight", d);
    m.e_ln_b = f32("encoder.ln_post.bias", d);

    auto fused_qkv = [&](const std::string & p, void *& w, float *& b) {
        w = up.dalloc((size_t) 3 * d * d * 2);
        b = (float *) up.dalloc((size_t) 3 * d * 4);
        if (!w || !b) return;
        up.put_rows(T(p + ".query.weight"), w, 0, d, d, d);
        up.put_rows(T(p + ".key.weight"), w, d, d, d, d);
        up.put_rows(T(p + ".value.weight"), w, 2 * d, d, d, d);
        up.put_f32(T(p + ".query.bias"), b, 0, d);
        up.put_f32(T(p + ".value.bias"), b, 2 * (size_t) d, d);
    };

    m.enc.resize(n_al);
    for (int i = 0; i < n_al && up.ok; ++i) {
        const std::string p = "encoder.blocks." + std::to_string(i);
        EncLayer & L = m.enc[i];
        L.ln1_w = f32(p + ".attn_ln.weight", d);
        L.ln1_b = f32(p + ".attn_ln.bias", d);
        fused_qkv(p + ".attn", L.wqkv, L.bqkv);
        L.wo = w16(p + ".attn.out.weight", d, d);
        L.bo = f32(p + ".attn.out.bias", d);
        L.ln2_w = f32(p + ".mlp_ln.weight", d);
        L.ln2_b = f32(p + ".mlp_ln.bias", d);
        L.w1 = w16(p + ".mlp.0.weight", 4 * d, d);
        L.b1 = f32(p + ".mlp.0.bias", 4 * (size_t) d);
        L.w2 = w16(p + ".mlp.2.weight", d, 4 * d);
        L.b2 = f32(p + ".mlp.2.bias", d);
    }

    m.d_pe = f32("decoder.positional_embedding", (size_t) hp.n_text_ctx * d);
    m.d_te = w16("decoder.token_embedding.weight", n_vocab, d);
    m.d_ln_w = f32("decoder.ln.weight", d);
    m.d_ln_b = f32("decoder.ln.bias", d);
    m.dec.resize(n_tl);
    for (int i = 0; i < n_tl && up.ok; ++i) {
        const std::string p = "decoder.blocks." + std::to_string(i);
        DecLayer & L = m.dec[i];
        L.ln1_w = f32(p + ".attn_ln.weight", d);
        L.ln1_b = f32(p + ".attn_ln.bias", d);
        fused_qkv(p + ".attn", L.wqkv, L.bqkv);
        L.wo = w16(p + ".attn.out.weight", d, d);
        L.bo = f32(p + ".attn.out.bias", d);
        L.lnx_w = f32(p + ".cross_attn_ln.weight", d);
        L.lnx_b = f32(p + ".cross_attn_ln.bias", d);
        L.wxq = w16(p + ".cross_attn.query.weight", d, d);
        L.bxq = f32(p + ".cross_attn.query.bias", d);
        L.wxkv = up.dalloc((size_t) 2 * d * d * 2);
        L.bxkv = (float *) up.dalloc((size_t) 2 * d * 4);
        if (L.wxkv && L.bxkv) {
            up.put_rows(T(p + ".cross_attn.key.weight"), L.wxkv, 0, d, d, d);
            up.put_rows(T(p + ".cross_attn.value.weight"), L.wxkv, d, d, d, d);
            up.put_f32(T(p + ".cross_attn.value.bias"), L.bxkv, d, d);
        }
        L.wxo = w16(p + ".cross_attn.out.weight", d, d);
        L.bxo = f32(p + ".cross_attn.out.bias", d);
        L.ln2_w = f32(p + ".mlp_ln.weight", d);
        L.ln2_b = f32(p + ".mlp_ln.bias", d);
        L.w1 = w16(p + ".mlp.0.weight", 4 * d, d);
        L.b1 = f32(p + ".mlp.0.bias", 4 * (size_t) d);
        L.w2 = w16(p + ".mlp.2.weight", d, 4 * d);
        L.b2 = f32(p + ".mlp.2.bias", d);
        up.ln_fold(L.wqkv, L.ln1_w, L.ln1_b, L.bqkv, 3 * d, d, L.qkv_c, L.qkv_b);
        up.ln_fold(L.wxq, L.lnx_w, L.lnx_b, L.bxq, d, d, L.xq_c, L.xq_b);
        up.ln_fold(L.w1, L.ln2_w, L.ln2_b, L.b1, 4 * d, d, L.m1_c, L.m1_b);
    }
    WB_CUDA(cudaDeviceSynchronize());
    if (!up.ok || cuda_failed()) {
        wlog(GGML_LOG_LEVEL_ERROR, "%s: failed to place the weights on CUDA device %d\n", __func__, device);
        return false;
    }
    if (n_quantised)
        wlog(GGML_LOG_LEVEL_INFO, "%s: %d quantised tensors (ftype %d) expanded to 16-bit weights at load\n", __func__, n_quantised, hp.ftype);
    wlog(GGML_LOG_LEVEL_INFO, "%s: %s weights on device %d: %.2f MB\n", __func__, dtype == DType::F16 ? "f16" : "bf16", device,
         m.bytes_device / 1e6);
    return true;
}

}  // namespace wb

extern "C" WB200_API long long whisper_b200_dequantize_blocks(int ggml_type, const void * raw, long long n_blocks, uint16_t * out16) {
    if (!raw || !out16 || n_blocks < 0) return -1;
    __half * out = reinterpret_cast<__half *>(out16);
    if (wb::kquant_block_bytes(ggml_type)) {
        wb::dequantize_kblocks(ggml_type, (const unsigned char *) raw, (size_t) n_blocks, out);
        return n_blocks * 256;
    }
    if (wb::quant_block_bytes(ggml_type)) {
        wb::dequantize_blocks(ggml_type, (const unsigned char *) raw, (size_t) n_blocks, out);
        return n_blocks * 32;
    }
    return -1;
}
