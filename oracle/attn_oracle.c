/*
 * C restatement of the attention forward for one sequence.  TEST INFRASTRUCTURE ONLY: used by tests/
 * to cross-check oracle/attention_oracle.py and by bench.py as the CPU-baseline "port" when the
 * compiled reference (oracle/_ref) is not available.  The product never links or loads this file.
 *
 * Semantics restated (see oracle/attention_oracle.py for the full citation list):
 *   scores/masks/bias/softcap  reference include/mat_mul.h:82-157
 *   softmax, output, LSE       reference include/softmax.h:80-95, kernel/fused_mha_forward.cu:220-223
 *   GQA head map               reference include/template.h:71-73
 *   RoPE                       reference include/rotary.h:95-137 (same fmaf / rounding order)
 * Structure follows the reference's own CPU loop `cpu_attention`
 * (utils/sass/mma_swizzle/forward_kernel.cu:346-370): per row, scores -> max -> exp -> normalise -> PV.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NEG_SENTINEL (-1e30f)

#include <pthread.h>

typedef struct {
    const float *q, *k, *v;
    float *out, *lse;
    int Sq, Sk, H, Hk, D;
    float scale;
    int wl, wr;
    const float* slopes;
    float softcap;
    int tid, nthreads;
} oracle_job_t;

static void* oracle_worker(void* arg) {
    const oracle_job_t* J = (const oracle_job_t*)arg;
    const int Sq = J->Sq, Sk = J->Sk, H = J->H, Hk = J->Hk, D = J->D;
    const int g = H / Hk;
    const int off = Sk - Sq;
    double* s = (double*)malloc(sizeof(double) * (size_t)(Sk > 0 ? Sk : 1));
    for (long idx = J->tid; idx < (long)H * Sq; idx += J->nthreads) {
        const int h = (int)(idx / Sq), i = (int)(idx % Sq);
        const int hk = h / g;
        const float* qi = J->q + ((size_t)i * H + h) * D;
        int lo = 0, hi = Sk; /* visible keys [lo, hi) */
        if (J->wr >= 0 && i + off + J->wr + 1 < hi) hi = i + off + J->wr + 1;
        if (J->wl >= 0 && i + off - J->wl > lo) lo = i + off - J->wl;
        double m = -INFINITY;
        for (int j = lo; j < hi; ++j) {
            const float* kj = J->k + ((size_t)j * Hk + hk) * D;
            double acc = 0.0;
            for (int d = 0; d < D; ++d) acc += (double)qi[d] * (double)kj[d];
            double x = acc * (double)J->scale;
            if (J->slopes) x -= (double)J->slopes[h] * fabs((double)(i + off - j));
            if (J->softcap > 0.f) x = (double)J->softcap * tanh(x / (double)J->softcap);
            s[j] = x;
            if (x > m) m = x;
        }
        float* oi = J->out + ((size_t)i * H + h) * D;
        if (hi <= lo) {
            memset(oi, 0, sizeof(float) * (size_t)D);
            J->lse[(size_t)h * Sq + i] = NEG_SENTINEL;
            continue;
        }
        double l = 0.0;
        for (int j = lo; j < hi; ++j) {
            s[j] = exp(s[j] - m);
            l += s[j];
        }
        for (int d = 0; d < D; ++d) {
            double acc = 0.0;
            for (int j = lo; j < hi; ++j) acc += s[j] * (double)J->v[((size_t)j * Hk + hk) * D + d];
            oi[d] = (float)(acc / l);
        }
        J->lse[(size_t)h * Sq + i] = (float)(m + log(l));
    }
    free(s);
    return NULL;
}

/* q:[Sq,H,D]  k,v:[Sk,Hk,D]  out:[Sq,H,D]  lse:[H,Sq]; wl/wr = -1 for unbounded (causal: wr = 0).
 * (head,row) pairs are dealt round-robin to `nthreads` pthreads. */
void oracle_attention(const float* q, const float* k, const float* v, float* out, float* lse, int Sq, int Sk,
                      int H, int Hk, int D, float scale, int wl, int wr, const float* slopes, float softcap,
                      int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    oracle_job_t jobs[256];
    for (int t = 0; t < nthreads; ++t) {
        oracle_job_t j = {q, k, v, out, lse, Sq, Sk, H, Hk, D, scale, wl, wr, slopes, softcap, t, nthreads};
        jobs[t] = j;
        pthread_create(&th[t], NULL, oracle_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
}

/* Rotate one head vector (fp32 values that are exactly representable in the 16-bit dtype).
 * y is fp32; the caller rounds to the 16-bit dtype. */
void oracle_rope(const float* x, float* y, const float* cosr, const float* sinr, int D, int rot, int interleaved) {
    const int half = rot / 2;
    for (int d = 0; d < D; ++d) y[d] = x[d];
    for (int a = 0; a < half; ++a) {
        const int i0 = interleaved ? 2 * a : a;
        const int i1 = interleaved ? 2 * a + 1 : a + half;
        const float x0 = x[i0], x1 = x[i1], c = cosr[a], s = sinr[a];
        y[i0] = fmaf(x0, c, -(x1 * s));
        y[i1] = fmaf(x0, s, x1 * c);
    }
}
