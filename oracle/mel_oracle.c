/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's log-mel front-end.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this; the product never does.
 * Pinned against the compiled reference (oracle/_ref, ref_log_mel) by tests/test_oracle_pinning.py and the
 * golden vectors in tests/golden/mel_*.npz.
 *
 * Follows, step for step (same fp32 operations in the same order, same table values):
 *   - tables:        whisper_global_cache            reference src/whisper.cpp:2998-3033
 *   - DFT-25 leaf:   dft()                            reference src/whisper.cpp:3038-3054
 *   - radix-2 steps: fft() (400->200->100->50->25)    reference src/whisper.cpp:3060-3102
 *   - frame loop:    log_mel_spectrogram_worker_thread reference src/whisper.cpp:3104-3167
 *   - pad/clamp:     log_mel_spectrogram              reference src/whisper.cpp:3170-3260
 * The recursion of the reference is restated iteratively: the 16 leaves are the decimated sequences
 * x[r + 16 m]; they are combined pairwise in four passes.  The arithmetic per output element is unchanged.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define N_FFT 400
#define HOP 160
#define N_BINS 201
#define PAD30 480000

static float g_sin[N_FFT], g_cos[N_FFT], g_hann[N_FFT];
static int g_ready = 0;

static void init_tables(void) {
    if (g_ready) return;
    for (int i = 0; i < N_FFT; i++) {
        double theta = (2 * M_PI * i) / N_FFT;
        g_sin[i] = sinf(theta);
        g_cos[i] = cosf(theta);
        g_hann[i] = 0.5 * (1.0 - cosf((2.0 * M_PI * i) / (N_FFT)));
    }
    g_ready = 1;
}

/* 25-point DFT of a real sequence taken with stride `stride` from `in` (table step 16) */
static void leaf_dft25(const float * in, int stride, float * out /* 50 floats */) {
    for (int k = 0; k < 25; k++) {
        float re = 0;
        float im = 0;
        for (int n = 0; n < 25; n++) {
            int idx = (k * n * 16) % N_FFT;
            re += in[n * stride] * g_cos[idx];
            im -= in[n * stride] * g_sin[idx];
        }
        out[2 * k] = re;
        out[2 * k + 1] = im;
    }
}

/* combine two length-h spectra (even, odd) into one of length 2h */
static void combine(const float * ev, const float * od, int h, float * out) {
    const int step = N_FFT / (2 * h);
    for (int k = 0; k < h; k++) {
        int idx = k * step;
        float re = g_cos[idx];
        float im = -g_sin[idx];
        float re_odd = od[2 * k];
        float im_odd = od[2 * k + 1];
        out[2 * k] = ev[2 * k] + re * re_odd - im * im_odd;
        out[2 * k + 1] = ev[2 * k + 1] + re * im_odd + im * re_odd;
        out[2 * (k + h)] = ev[2 * k] - re * re_odd + im * im_odd;
        out[2 * (k + h) + 1] = ev[2 * k + 1] - re * im_odd - im * re_odd;
    }
}

/* spectrum of x[off + stride*m], m < n, n in {25,50,100,200,400} */
static void fft_rec_free(const float * x, int off, int stride, int n, float * out, float * scratch) {
    if (n == 25) {
        leaf_dft25(x + off, stride, out);
        return;
    }
    float * ev = scratch;
    float * od = scratch + n;          /* each half needs 2*(n/2) = n floats */
    float * deeper = scratch + 2 * n;
    fft_rec_free(x, off, stride * 2, n / 2, ev, deeper);
    fft_rec_free(x, off + stride, stride * 2, n / 2, od, deeper);
    combine(ev, od, n / 2, out);
}

/* out: [n_mel][n_len] f32.  Returns n_len (or -1); n_len_org via pointer.  Single threaded. */
int oracle_log_mel(const float * samples, int n_samples, const float * filters /* [n_mel][201] */, int n_mel,
                   float * out, long out_cap, int * n_len_org) {
    init_tables();
    const long n_padded = (long) n_samples + PAD30 + N_FFT;   /* 200 reflect + audio + 30 s zeros + 200 zeros */
    const int n_len = (int) ((n_padded - N_FFT) / HOP);
    if (n_len_org) *n_len_org = 1 + (n_samples + N_FFT / 2 - N_FFT) / HOP;
    if (!out) return n_len;
    if (out_cap < (long) n_len * n_mel) return -1;

    float * pad = (float *) calloc((size_t) n_padded, sizeof(float));
    memcpy(pad + N_FFT / 2, samples, (size_t) n_samples * sizeof(float));
    for (int i = 0; i < N_FFT / 2; i++) pad[i] = samples[N_FFT / 2 - i];   /* reflect at the start only */

    const int n_in = n_samples + N_FFT / 2;
    int n_frames_fft = n_in / HOP + 1;
    if (n_frames_fft > n_len) n_frames_fft = n_len;

    float fft_in[N_FFT], spec[2 * N_FFT], scratch[4 * N_FFT], power[N_BINS];
    for (int i = 0; i < n_frames_fft; i++) {
        const int offset = i * HOP;
        int lim = n_in - offset;
        if (lim > N_FFT) lim = N_FFT;
        for (int j = 0; j < lim; j++) fft_in[j] = g_hann[j] * pad[offset + j];
        for (int j = lim < 0 ? 0 : lim; j < N_FFT; j++) fft_in[j] = 0.0f;
        fft_rec_free(fft_in, 0, 1, N_FFT, spec, scratch);
        for (int j = 0; j < N_BINS; j++) power[j] = spec[2 * j] * spec[2 * j] + spec[2 * j + 1] * spec[2 * j + 1];
        for (int j = 0; j < n_mel; j++) {
            const float * f = filters + (long) j * N_BINS;
            double sum = 0.0;
            int k = 0;
            for (k = 0; k < N_BINS - 3; k += 4) {
                sum += power[k] * f[k] + power[k + 1] * f[k + 1] + power[k + 2] * f[k + 2] + power[k + 3] * f[k + 3];
            }
            for (; k < N_BINS; k++) sum += power[k] * f[k];
            sum = log10(sum > 1e-10 ? sum : 1e-10);
            out[(long) j * n_len + i] = sum;
        }
    }
    const double floor_v = log10(1e-10);
    for (int i = n_frames_fft; i < n_len; i++)
        for (int j = 0; j < n_mel; j++) out[(long) j * n_len + i] = floor_v;
    free(pad);

    double mmax = -1e20;
    for (long i = 0; i < (long) n_mel * n_len; i++)
        if (out[i] > mmax) mmax = out[i];
    mmax -= 8.0;
    for (long i = 0; i < (long) n_mel * n_len; i++) {
        if (out[i] < mmax) out[i] = mmax;
        out[i] = (out[i] + 4.0) / 4.0;
    }
    return n_len;
}

/* Exact (float64, direct DFT) restatement used to put both the reference and the CUDA kernel on a common scale:
 * how far is each from the mathematically exact log-mel?  O(n_frames*201*400) -- small inputs only. */
int oracle_log_mel_f64(const float * samples, int n_samples, const float * filters, int n_mel, double * out,
                       long out_cap) {
    init_tables();
    const long n_padded = (long) n_samples + PAD30 + N_FFT;
    const int n_len = (int) ((n_padded - N_FFT) / HOP);
    if (out_cap < (long) n_len * n_mel) return -1;
    double * pad = (double *) calloc((size_t) n_padded, sizeof(double));
    for (int i = 0; i < n_samples; i++) pad[N_FFT / 2 + i] = samples[i];
    for (int i = 0; i < N_FFT / 2; i++) pad[i] = samples[N_FFT / 2 - i];
    const int n_in = n_samples + N_FFT / 2;
    int n_frames_fft = n_in / HOP + 1;
    if (n_frames_fft > n_len) n_frames_fft = n_len;
    static double c[N_FFT], s[N_FFT];
    for (int i = 0; i < N_FFT; i++) {
        c[i] = cos(2 * M_PI * i / N_FFT);
        s[i] = sin(2 * M_PI * i / N_FFT);
    }
    double w[N_FFT], power[N_BINS];
    for (int i = 0; i < n_frames_fft; i++) {
        for (int j = 0; j < N_FFT; j++) {
            const long p = (long) i * HOP + j;
            w[j] = (p < n_in ? pad[p] : 0.0) * (double) g_hann[j];
        }
        for (int k = 0; k < N_BINS; k++) {
            double re = 0, im = 0;
            for (int n = 0; n < N_FFT; n++) {
                const int idx = (int) (((long) k * n) % N_FFT);
                re += w[n] * c[idx];
                im -= w[n] * s[idx];
            }
            power[k] = re * re + im * im;
        }
        for (int j = 0; j < n_mel; j++) {
            double sum = 0;
            for (int k = 0; k < N_BINS; k++) sum += power[k] * (double) filters[(long) j * N_BINS + k];
            out[(long) j * n_len + i] = log10(sum > 1e-10 ? sum : 1e-10);
        }
    }
    for (int i = n_frames_fft; i < n_len; i++)
        for (int j = 0; j < n_mel; j++) out[(long) j * n_len + i] = -10.0;
    free(pad);
    double mmax = -1e20;
    for (long i = 0; i < (long) n_mel * n_len; i++)
        if (out[i] > mmax) mmax = out[i];
    mmax -= 8.0;
    for (long i = 0; i < (long) n_mel * n_len; i++) {
        if (out[i] < mmax) out[i] = mmax;
        out[i] = (out[i] + 4.0) / 4.0;
    }
    return n_len;
}
