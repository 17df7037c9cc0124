/*
 * precise_oracle.c -- plain-C restatement of the Precise streaming-inference hot path.
 * TEST INFRASTRUCTURE ONLY (second, independent oracle + a compiled CPU baseline).  The product never links it.
 *
 * Follows, like oracle/*.py:
 *   buffer_to_audio           precise/util.py:35-37
 *   Listener.update_vectors   precise/network_runner.py:125-146   (carry buffer re-featurised when >= window samples)
 *   sonopy.mfcc_spec          as called at precise/vectorization.py:36-39   ** PARITY UNPINNED (see oracle/mfcc.py) **
 *   Keras GRU + Dense         precise/model.py:77-82                         ** PARITY UNPINNED (see oracle/gru.py)  **
 *   ThresholdDecoder          precise/threshold_decoder.py:38-57, precise/functions.py:94-108
 *   TriggerDetector           runner/precise_runner/runner.py:115-142
 *
 * Arithmetic types as in the reference: MFCC in float64, network in float32, decode in float64.
 * Build: gcc -O2 -shared -fPIC -o _build/libprecise_oracle.so precise_oracle.c -lm   (oracle/c/Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PO_MAX_FILT 64
#define PO_MAX_FFT 4096

typedef struct {
    int sample_rate, window, hop, n_fft, n_filt, n_mfcc, n_features;
    int hidden;                 /* GRU units */
    int n_thresholds;
    double mu[8], sd[8], center;
    double sensitivity;
    int trigger_level;
    int chunk_samples;
} po_params;

typedef struct {
    po_params p;
    int n_bins, n_out;
    double* fbank;              /* [n_filt][n_bins] */
    double* dct;                /* [n_out][n_filt] */
    double* cd;                 /* CDF table */
    int cd_len, min_out, max_out;
    const float *kernel, *recurrent, *bias, *dense_w;   /* borrowed */
    float dense_b;
} po_model;

typedef struct {
    double* carry;              /* window_audio */
    int carry_len, carry_cap;
    double* mfccs;              /* [n_features][n_out], oldest first */
    int activation;
} po_stream;

static const double PO_EPS = 2.220446049250313e-16;

/* ------------------------------------------------------------------ tables */
static void build_filterbank(po_model* m) {
    const po_params* p = &m->p;
    int nb = m->n_bins, nf = p->n_filt, grid[PO_MAX_FILT + 2];
    double top = 1127.0 * log(1.0 + (double)p->sample_rate / 700.0);
    long shift = 0, prev = 0;
    for (int i = 0; i < nf + 2; ++i) {
        double mel = (i == nf + 1) ? top : (double)i * (top / (double)(nf + 1));
        double hz = 700.0 * (exp(mel / 1127.0) - 1.0);
        long raw = (long)(hz * (double)nb / (double)p->sample_rate);
        if (i == 0) prev = raw - 1;
        long s = shift + prev + 1 - raw;
        shift = s > 0 ? s : 0;
        grid[i] = (int)(raw + shift);
        prev = raw;
    }
    m->fbank = (double*)calloc((size_t)nf * nb, sizeof(double));
    for (int i = 0; i < nf; ++i) {
        int lo = grid[i], mid = grid[i + 1], hi = grid[i + 2];
        for (int k = lo; k < mid && k < nb; ++k) m->fbank[(size_t)i * nb + k] = (double)(k - lo) * (1.0 / (double)(mid - lo));
        for (int k = mid; k < hi && k < nb; ++k) m->fbank[(size_t)i * nb + k] = (double)(k - mid) * (-1.0 / (double)(hi - mid)) + 1.0;
    }
}

static void build_dct(po_model* m) {
    int nf = m->p.n_filt, no = m->n_out;
    m->dct = (double*)malloc((size_t)no * nf * sizeof(double));
    for (int k = 0; k < no; ++k)
        for (int n = 0; n < nf; ++n) {
            double v = cos(M_PI * k * (2 * n + 1) / (2.0 * nf)) * sqrt(2.0 / nf);
            if (k == 0) v *= sqrt(0.5);
            m->dct[(size_t)k * nf + n] = v;
        }
}

static void build_cdf(po_model* m) {
    const po_params* p = &m->p;
    double lo = 0, hi = 0;
    for (int i = 0; i < p->n_thresholds; ++i) {
        double a = p->mu[i] - 4 * p->sd[i], b = p->mu[i] + 4 * p->sd[i];
        if (i == 0 || a < lo) lo = a;
        if (i == 0 || b > hi) hi = b;
    }
    m->min_out = (int)lo; m->max_out = (int)hi;
    int range = m->max_out - m->min_out, num = 200 * range;
    m->cd_len = num > 0 ? num : 0;
    m->cd = (double*)calloc(num > 0 ? num : 1, sizeof(double));
    double step = num > 1 ? (double)range / (double)(num - 1) : 0.0, run = 0.0;
    for (int j = 0; j < num; ++j) {
        double x = (j == num - 1 && num > 1) ? (double)m->max_out : (double)j * step + (double)m->min_out, s = 0.0;
        for (int i = 0; i < p->n_thresholds; ++i) {
            double sd = p->sd[i], d = x - p->mu[i];
            double pd = sd == 0 ? 0.0 : (1.0 / (sd * sqrt(2 * M_PI))) * exp(-(d * d) / (2 * (sd * sd)));
            s = (i == 0) ? pd : s + pd;
        }
        run += s / (double)(200 * p->n_thresholds);
        m->cd[j] = run;
    }
}

po_model* po_model_create(const po_params* p, const float* kernel, const float* recurrent, const float* bias,
                          const float* dense_w, float dense_b) {
    po_model* m = (po_model*)calloc(1, sizeof(po_model));
    m->p = *p;
    m->n_bins = p->n_fft / 2 + 1;
    m->n_out = p->n_filt < p->n_mfcc ? p->n_filt : p->n_mfcc;
    build_filterbank(m); build_dct(m); build_cdf(m);
    m->kernel = kernel; m->recurrent = recurrent; m->bias = bias; m->dense_w = dense_w; m->dense_b = dense_b;
    return m;
}

void po_model_destroy(po_model* m) {
    if (!m) return;
    free(m->fbank); free(m->dct); free(m->cd); free(m);
}

/* ------------------------------------------------------------------ MFCC (float64) */
/* in-place iterative radix-2 complex FFT, n a power of two */
static void fft_c(double* re, double* im, int n) {
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        double ang = -2.0 * M_PI / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < len / 2; ++k) {
                double wr = cos(ang * k), wi = sin(ang * k);
                double ur = re[i + k], ui = im[i + k];
                double vr = re[i + k + len / 2] * wr - im[i + k + len / 2] * wi;
                double vi = re[i + k + len / 2] * wi + im[i + k + len / 2] * wr;
                re[i + k] = ur + vr; im[i + k] = ui + vi;
                re[i + k + len / 2] = ur - vr; im[i + k + len / 2] = ui - vi;
            }
    }
}

/* one frame: first min(window, n_fft) samples of `frame` -> out[n_out] */
static void mfcc_frame(const po_model* m, const double* frame, double* out) {
    const po_params* p = &m->p;
    double re[PO_MAX_FFT], im[PO_MAX_FFT], power[PO_MAX_FFT / 2 + 1], mels[PO_MAX_FILT];
    int used = p->window < p->n_fft ? p->window : p->n_fft, nb = m->n_bins;
    for (int i = 0; i < p->n_fft; ++i) { re[i] = i < used ? frame[i] : 0.0; im[i] = 0.0; }
    fft_c(re, im, p->n_fft);
    double tot = 0.0;
    for (int k = 0; k < nb; ++k) { power[k] = (re[k] * re[k] + im[k] * im[k]) / p->n_fft; tot += power[k]; }
    for (int j = 0; j < p->n_filt; ++j) {
        double s = 0.0;
        const double* f = m->fbank + (size_t)j * nb;
        for (int k = 0; k < nb; ++k) s += power[k] * f[k];
        mels[j] = log(s > PO_EPS ? s : PO_EPS);
    }
    for (int c = 0; c < m->n_out; ++c) {
        double s = 0.0;
        for (int j = 0; j < p->n_filt; ++j) s += m->dct[(size_t)c * p->n_filt + j] * mels[j];
        out[c] = s;
    }
    out[0] = log(tot > PO_EPS ? tot : PO_EPS);
}

/* vectorize_raw: returns the number of frames written to out[frames][n_out] */
int po_mfcc(const po_model* m, const double* audio, int n, double* out, int max_frames) {
    const po_params* p = &m->p;
    int nf = n < p->window ? 0 : (n - p->window) / p->hop + 1;
    if (nf > max_frames) nf = max_frames;
    for (int f = 0; f < nf; ++f) mfcc_frame(m, audio + (size_t)f * p->hop, out + (size_t)f * m->n_out);
    return nf;
}

/* ------------------------------------------------------------------ network (float32) */
static float hard_sigmoid(float x) { float v = 0.2f * x + 0.5f; return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); }

float po_gru(const po_model* m, const float* x /* [T][F] */, float* logit_out) {
    int H = m->p.hidden, F = m->n_out, T = m->p.n_features, H3 = 3 * H;
    float h[256] = {0}, a[768], hn[256];
    for (int t = 0; t < T; ++t) {
        const float* xt = x + (size_t)t * F;
        for (int j = 0; j < H3; ++j) {
            float s = 0.f;
            for (int f = 0; f < F; ++f) s += xt[f] * m->kernel[(size_t)f * H3 + j];
            a[j] = s + m->bias[j];
        }
        for (int j = 0; j < H; ++j) {
            float sz = 0.f, sr = 0.f;
            for (int k = 0; k < H; ++k) { sz += h[k] * m->recurrent[(size_t)k * H3 + j]; sr += h[k] * m->recurrent[(size_t)k * H3 + H + j]; }
            a[j] = hard_sigmoid(a[j] + sz);
            a[H + j] = hard_sigmoid(a[H + j] + sr) * h[j];          /* r * h */
        }
        for (int j = 0; j < H; ++j) {
            float s = 0.f;
            for (int k = 0; k < H; ++k) s += a[H + k] * m->recurrent[(size_t)k * H3 + 2 * H + j];
            float hh = a[2 * H + j] + s;
            hn[j] = a[j] * h[j] + (1.f - a[j]) * hh;
        }
        memcpy(h, hn, (size_t)H * sizeof(float));
    }
    float lg = 0.f;
    for (int j = 0; j < H; ++j) lg += h[j] * m->dense_w[j];
    lg += m->dense_b;
    if (logit_out) *logit_out = lg;
    return 1.f / (1.f + expf(-lg));
}

/* ------------------------------------------------------------------ decode / trigger */
double po_decode(const po_model* m, float raw) {
    double r = (double)raw, cp;
    if (raw == 1.0f || raw == 0.0f) return r;
    int range = m->max_out - m->min_out;
    if (range == 0) cp = r > (double)m->min_out ? 1.0 : 0.0;
    else {
        double ratio = (-log(1.0 / r - 1.0) - (double)m->min_out) / (double)range;
        ratio = ratio < 0.0 ? 0.0 : (ratio > 1.0 ? 1.0 : ratio);
        cp = m->cd[(int)(ratio * (double)(m->cd_len - 1) + 0.5)];
    }
    if (cp < m->p.center) return 0.5 * cp / m->p.center;
    return 0.5 + 0.5 * (cp - m->p.center) / (1 - m->p.center);
}

int po_trigger(const po_model* m, int* activation, double prob) {
    int hot = prob > 1.0 - m->p.sensitivity, fired = 0, a = *activation;
    if (hot || a < 0) {
        a += 1;
        fired = a > m->p.trigger_level;
        if (fired || (hot && a < 0)) {
            long bytes = 2L * m->p.chunk_samples, q = -(8 * 2048) / bytes;
            if ((-(8 * 2048)) % bytes != 0) q -= 1;                   /* python floor division */
            a = (int)q;
        }
    } else if (a > 0) a -= 1;
    *activation = a;
    return fired;
}

/* ------------------------------------------------------------------ streaming listener */
po_stream* po_stream_create(const po_model* m) {
    po_stream* s = (po_stream*)calloc(1, sizeof(po_stream));
    s->carry_cap = m->p.window + 2 * m->p.chunk_samples + 16;
    s->carry = (double*)malloc((size_t)s->carry_cap * sizeof(double));
    s->mfccs = (double*)calloc((size_t)m->p.n_features * m->n_out, sizeof(double));
    return s;
}
void po_stream_destroy(po_stream* s) { if (s) { free(s->carry); free(s->mfccs); free(s); } }

/* Listener.update on one int16 chunk; returns the decoded confidence, *raw_out the network output, *fired the trigger */
double po_update(const po_model* m, po_stream* s, const int16_t* chunk, int n, float* raw_out, int* fired) {
    const po_params* p = &m->p;
    int no = m->n_out, T = p->n_features;
    if (s->carry_len + n > s->carry_cap) {
        s->carry_cap = s->carry_len + n + p->window;
        s->carry = (double*)realloc(s->carry, (size_t)s->carry_cap * sizeof(double));
    }
    for (int i = 0; i < n; ++i) s->carry[s->carry_len + i] = (double)((float)chunk[i] / 32768.0f);   /* buffer_to_audio: float32 / 32768 */
    s->carry_len += n;
    if (s->carry_len >= p->window) {
        int nf = (s->carry_len - p->window) / p->hop + 1;
        double* feats = (double*)malloc((size_t)nf * no * sizeof(double));
        po_mfcc(m, s->carry, s->carry_len, feats, nf);
        int drop = nf * p->hop;
        memmove(s->carry, s->carry + drop, (size_t)(s->carry_len - drop) * sizeof(double));
        s->carry_len -= drop;
        const double* src = feats;
        int keep = nf;
        if (keep > T) { src += (size_t)(keep - T) * no; keep = T; }
        memmove(s->mfccs, s->mfccs + (size_t)keep * no, (size_t)(T - keep) * no * sizeof(double));
        memcpy(s->mfccs + (size_t)(T - keep) * no, src, (size_t)keep * no * sizeof(double));
        free(feats);
    }
    float x[64 * 64];
    for (int i = 0; i < T * no; ++i) x[i] = (float)s->mfccs[i];
    float raw = po_gru(m, x, 0);
    if (raw_out) *raw_out = raw;
    double conf = po_decode(m, raw);
    int f = po_trigger(m, &s->activation, conf);
    if (fired) *fired = f;
    return conf;
}

/* S independent streams over pcm[S][n_chunks * chunk]; the caller shards streams over threads.  raw/conf/fired: [S][n_chunks] (may be NULL).
 * Returns the number of detections. */
long po_run_streams(const po_model* m, const int16_t* pcm, int n_streams, int n_chunks, float* raw, double* conf, unsigned char* fired) {
    long total = 0;
    int chunk = m->p.chunk_samples;
    /* callers shard the streams over threads (ctypes releases the GIL); no OpenMP runtime is assumed */
    for (int s = 0; s < n_streams; ++s) {
        po_stream* st = po_stream_create(m);
        for (int k = 0; k < n_chunks; ++k) {
            float r; int f;
            double c = po_update(m, st, pcm + ((size_t)s * n_chunks + k) * chunk, chunk, &r, &f);
            if (raw) raw[(size_t)s * n_chunks + k] = r;
            if (conf) conf[(size_t)s * n_chunks + k] = c;
            if (fired) fired[(size_t)s * n_chunks + k] = (unsigned char)f;
            total += f;
        }
        po_stream_destroy(st);
    }
    return total;
}

int po_mfcc_width(const po_model* m) { return m->n_out; }
