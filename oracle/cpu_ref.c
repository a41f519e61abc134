/*
 * oracle/cpu_ref.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the numeric kernels the reference's `sink` reaches on
 * the hot path.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline may load this file; libsignalops_cuda.so never does.
 *
 * The kernels live in DSP.jl 0.6.10 (pinned at docs/Manifest.toml:48-52 of the
 * reference), which is NOT vendored under /root/reference and cannot be run here
 * (no Julia): they are restated from the published algorithm as recorded in
 * SURVEY.md Appendix B.  PARITY UNPINNED for the absolute output of Filt /
 * ToFramerate: the reference ships no golden vectors for them;
 * the restatement is cross-checked against scipy.signal in tests/.
 *
 * Reference call sites restated here:
 *   oracle_sos_filt        DSP `filt!(out, DF2TFilter{SOS}, x)`   src/filters.jl:252-255
 *   oracle_fir_filt        DSP `filt!(out, FIRFilter{...}, x)`    src/filters.jl:252-255
 *                          (filter built at src/reformatting.jl:92-99)
 *   oracle_iir_amplify_batch  the whole `Filt |> Amplify |> sink` pull loop:
 *                          src/sink.jl:225-267 + src/filters.jl:221-262 +
 *                          src/mapsignal.jl:249-255, one signal per thread
 *   oracle_resample_batch  `ToFramerate |> sink` the same way
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* DSP.jl `_filt!(out, si, f::SecondOrderSections, x, col)`, SURVEY.md App. B.2.
 * coef: M rows [b0 b1 b2 a1 a2]; si: M rows [s1 s2], carried between calls. */
void oracle_sos_filt(double* out, const double* x, int64_t n, const double* coef, int M,
                     double g, double* si) {
    for (int64_t i = 0; i < n; ++i) {
        double yi = x[i];
        for (int f = 0; f < M; ++f) {
            const double* c = coef + 5 * f;
            double* s = si + 2 * f;
            const double xi = yi;
            yi = s[0] + c[0] * xi;
            s[0] = s[1] + c[1] * xi - c[3] * yi;
            s[1] = c[2] * xi - c[4] * yi;
        }
        out[i] = yi * g;
    }
}

/* ---- FIR kernels (DSP.jl Filters/stream_filt.jl) ---------- */
enum { FIR_STANDARD = 0, FIR_INTERPOLATOR = 1, FIR_DECIMATOR = 2, FIR_RATIONAL = 3, FIR_ARBITRARY = 4 };

typedef struct {
    int32_t kind;
    int32_t n_phi;          /* Nphi                                     */
    int64_t taps_per_phi;   /* rows of pfb; hLen for standard/decimator  */
    int64_t h_len;
    int32_t interpolation, decimation, phi_step;
    int32_t phi_idx;        /* 1-based                                   */
    int64_t input_deficit;  /* 1-based                                   */
    int64_t x_idx;
    double rate, delta, phi_acc, alpha;
    const double* pfb;      /* Julia layout pfb[row + col*taps_per_phi]; h reversed for standard/decimator */
    const double* dpfb;
    double* history;        /* taps_per_phi-1 samples                    */
    int32_t simd;           /* 1: dot products reassociate into 8 lanes like DSP.jl's `@simd` loops
                               (the timing baseline); 0: strict left-to-right sums (the parity checker) */
} oracle_fir;

/* DSP.jl's unsafe_dot loops carry `@simd`, i.e. the compiler may reassociate the sum into vector
 * lanes (SURVEY.md §8c iii).  The timed CPU baseline does the same: eight partial sums, one clone per
 * instruction set picked at load time, so the baseline is not handicapped by a scalar dependency chain. */
__attribute__((target_clones("avx512f", "avx2", "default")))
static double dot_lanes(const double* a, const double* w, int64_t n) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t i = 0;
    for (; i + 8 <= n; i += 8)
        for (int j = 0; j < 8; ++j) acc[j] += a[i + j] * w[i + j];
    double d = 0.0;
    for (; i < n; ++i) d += a[i] * w[i];
    return (((acc[0] + acc[4]) + (acc[1] + acc[5])) + ((acc[2] + acc[6]) + (acc[3] + acc[7]))) + d;
}

/* unsafe_dot(pfb, col, history, x, xLastIdx) / unsafe_dot(pfb, col, x, xLastIdx); 1-based xLastIdx */
static double dot_window(const double* a, int64_t alen, const double* hist, const double* x, int64_t last, int simd) {
    double d = 0.0;
    if (simd && last >= alen) return dot_lanes(a, x + (last - alen), alen);
    if (last < alen) {
        const int64_t nh = alen - last;              /* taken from the end of history (length alen-1) */
        const double* h = hist + (alen - 1 - nh);
        for (int64_t i = 0; i < nh; ++i) d += a[i] * h[i];
        for (int64_t i = 0; i < last; ++i) d += a[nh + i] * x[i];
    } else {
        const double* w = x + (last - alen);
        for (int64_t i = 0; i < alen; ++i) d += a[i] * w[i];
    }
    return d;
}

static void shiftin(double* hist, int64_t hlen, const double* x, int64_t xlen) {
    if (hlen <= 0) return;
    if (xlen >= hlen) memcpy(hist, x + (xlen - hlen), (size_t)hlen * sizeof(double));
    else {
        memmove(hist, hist + xlen, (size_t)(hlen - xlen) * sizeof(double));
        memcpy(hist + (hlen - xlen), x, (size_t)xlen * sizeof(double));
    }
}

/* Returns the number of samples written (DSP's filt! return value, which the
 * reference discards at src/filters.jl:252-255 — see SURVEY.md App. C-2). */
int64_t oracle_fir_filt(double* buf, int64_t buflen, oracle_fir* k, const double* x, int64_t xlen) {
    const int64_t T = k->taps_per_phi;
    int64_t out = 0;
    if (xlen < k->input_deficit) {
        shiftin(k->history, T - 1, x, xlen);
        k->input_deficit -= xlen;
        return 0;
    }
    if (k->kind == FIR_ARBITRARY) {
        k->x_idx = k->input_deficit;
        while (k->x_idx <= xlen) {
            if (out >= buflen) return -1;
            const double* p = k->pfb + (int64_t)(k->phi_idx - 1) * T;
            const double* dp = k->dpfb + (int64_t)(k->phi_idx - 1) * T;
            const double lower = dot_window(p, T, k->history, x, k->x_idx, k->simd);
            const double upper = dot_window(dp, T, k->history, x, k->x_idx, k->simd);
            buf[out++] = lower + upper * k->alpha;
            /* update(kernel) */
            k->phi_acc += k->delta;
            if (k->phi_acc > k->n_phi) {
                k->x_idx += (int64_t)floor((k->phi_acc - 1.0) / k->n_phi);
                k->phi_acc = fmod(k->phi_acc - 1.0, (double)k->n_phi) + 1.0;
            }
            k->phi_idx = (int32_t)floor(k->phi_acc);
            k->alpha = k->phi_acc - k->phi_idx;
        }
        k->input_deficit = k->x_idx - xlen;
    } else if (k->kind == FIR_RATIONAL || k->kind == FIR_INTERPOLATOR) {
        int64_t idx = k->input_deficit;
        while (idx <= xlen) {
            if (out >= buflen) return -1;
            buf[out++] = dot_window(k->pfb + (int64_t)(k->phi_idx - 1) * T, T, k->history, x, idx, k->simd);
            idx += (k->phi_idx + k->decimation - 1) / k->interpolation;
            const int32_t v = k->phi_idx + k->phi_step;
            k->phi_idx = v > k->interpolation ? v - k->interpolation : v;
        }
        k->input_deficit = idx - xlen;
    } else { /* standard (decimation 1) / decimator */
        int64_t idx = k->input_deficit;
        const int64_t step = k->kind == FIR_DECIMATOR ? k->decimation : 1;
        while (idx <= xlen) {
            if (out >= buflen) return -1;
            buf[out++] = dot_window(k->pfb, T, k->history, x, idx, k->simd);
            idx += step;
        }
        k->input_deficit = idx - xlen;
    }
    shiftin(k->history, T - 1, x, xlen);
    return out;
}

/* ---- whole-pipeline restatements used as the CPU baseline -------------------------- */

/* No OpenMP runtime in this image: a plain pthread fan-out, signal s -> thread s % nthreads. */
#include <pthread.h>

typedef struct {
    int tid, nthreads;
    const double* in; double* out;
    int64_t ninst, nframes, n_in, n_out, blocksize;
    int nch, M;
    const double* coef; double g, amp;
    const oracle_fir* proto;
} job_t;

/* `Filt(x, design) |> Amplify(amp) |> sink` for one signal of nframes x nch,
 * channel-planar.  Follows the pull loop: for each 4096-frame block the child is
 * sunk into `input` (src/filters.jl:240-244), each channel is filtered into
 * `output` with its own DF2T state (:252-255), then the outer `sink_helper!`
 * reads frames from `output` and multiplies (src/mapsignal.jl:249-255). */
static void* iir_worker(void* arg) {
    job_t* j = (job_t*)arg;
    const int64_t nframes = j->nframes, blocksize = j->blocksize;
    const int nch = j->nch, M = j->M;
    double* input = (double*)malloc((size_t)blocksize * nch * sizeof(double));
    double* output = (double*)malloc((size_t)blocksize * nch * sizeof(double));
    double* si = (double*)malloc((size_t)2 * M * nch * sizeof(double));
    for (int64_t s = j->tid; s < j->ninst; s += j->nthreads) {
        const double* x = j->in + s * nframes * nch;
        double* y = j->out + s * nframes * nch;
        memset(si, 0, (size_t)2 * M * nch * sizeof(double));
        for (int64_t off = 0; off < nframes; off += blocksize) {
            const int64_t len = nframes - off < blocksize ? nframes - off : blocksize;
            for (int c = 0; c < nch; ++c)
                for (int64_t i = 0; i < blocksize; ++i)   /* Pad(x.signal,zero) */
                    input[c * blocksize + i] = i < len ? x[c * nframes + off + i] : 0.0;
            for (int c = 0; c < nch; ++c)
                oracle_sos_filt(output + c * blocksize, input + c * blocksize, blocksize, j->coef, M, j->g,
                                si + (size_t)2 * M * c);
            for (int64_t i = 0; i < len; ++i)             /* sink_helper!: frame by frame */
                for (int c = 0; c < nch; ++c) y[c * nframes + off + i] = output[c * blocksize + i] * j->amp;
        }
    }
    free(input); free(output); free(si);
    return 0;
}

/* `ToFramerate(x, fs) |> sink`, one-shot semantics;
 * `proto` holds the kernel state right after setphase!. */
static void* resample_worker(void* arg) {
    job_t* j = (job_t*)arg;
    const oracle_fir* proto = j->proto;
    const int64_t T = proto->taps_per_phi, blocksize = j->blocksize, n_in = j->n_in, n_out = j->n_out;
    const int64_t outcap = (int64_t)(blocksize * (proto->rate > 1 ? proto->rate : 1) + 64) * 2;
    double* hist = (double*)malloc((size_t)(T > 1 ? T - 1 : 1) * sizeof(double));
    double* input = (double*)malloc((size_t)blocksize * sizeof(double));
    double* output = (double*)malloc((size_t)outcap * sizeof(double));
    for (int64_t s = j->tid; s < j->ninst; s += j->nthreads) {
        for (int c = 0; c < j->nch; ++c) {
            oracle_fir k = *proto;
            k.history = hist;
            memset(hist, 0, (size_t)(T > 1 ? T - 1 : 1) * sizeof(double));
            const double* x = j->in + (s * j->nch + c) * n_in;
            double* y = j->out + (s * j->nch + c) * n_out;
            int64_t written = 0, off = 0;
            while (written < n_out) {
                for (int64_t i = 0; i < blocksize; ++i) input[i] = off + i < n_in ? x[off + i] : 0.0;
                off += blocksize;
                const int64_t w = oracle_fir_filt(output, outcap, &k, input, blocksize);
                if (w < 0) break;
                const int64_t take = w < n_out - written ? w : n_out - written;
                memcpy(y + written, output, (size_t)take * sizeof(double));
                written += take;
            }
        }
    }
    free(hist); free(input); free(output);
    return 0;
}

static void fan_out(void* (*fn)(void*), job_t* proto_job, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    pthread_t* th = (pthread_t*)malloc((size_t)nthreads * sizeof(pthread_t));
    job_t* jobs = (job_t*)malloc((size_t)nthreads * sizeof(job_t));
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = *proto_job;
        jobs[t].tid = t;
        jobs[t].nthreads = nthreads;
        pthread_create(&th[t], 0, fn, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
    free(th); free(jobs);
}

void oracle_iir_amplify_batch(const double* in, double* out, int64_t ninst, int64_t nframes,
                              int nch, const double* coef, int M, double g, double amp,
                              int64_t blocksize, int nthreads) {
    job_t j;
    memset(&j, 0, sizeof j);
    j.in = in; j.out = out; j.ninst = ninst; j.nframes = nframes; j.nch = nch;
    j.coef = coef; j.M = M; j.g = g; j.amp = amp; j.blocksize = blocksize;
    fan_out(iir_worker, &j, nthreads);
}

void oracle_resample_batch(const double* in, double* out, int64_t ninst, int64_t n_in, int64_t n_out,
                           int nch, const oracle_fir* proto, int64_t blocksize, int nthreads) {
    job_t j;
    memset(&j, 0, sizeof j);
    j.in = in; j.out = out; j.ninst = ninst; j.n_in = n_in; j.n_out = n_out; j.nch = nch;
    j.proto = proto; j.blocksize = blocksize;
    fan_out(resample_worker, &j, nthreads);
}
