/* Plain-C float64 restatement of CTC loss + gradient w.r.t. logits.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): never linked into or called
 * from the product library.  Built by oracle/Makefile into oracle/_build/.
 *
 * Algorithm: Graves et al. 2006 as written out in SURVEY.md Appendix B, with
 * torch's convention that alpha and beta both include the emission lp[t, l'_s].
 * The reference repository has no CTC source of its own; the op
 * restated here is torch.nn.functional.ctc_loss applied to
 * log_softmax(logits) with batch-major [B,T,V] logits
 * (/root/reference/Predictor/Utils/loss.py:10), int64 targets padded with 0
 * (/root/reference/data/data_loader/ai_shell_1.py:75-88), blank = 0
 * (/root/reference/Predictor/data_handler/vocab.py:10,17).
 *
 * One utterance at a time, everything in double; utterances are independent so a
 * small pthread pool pulls utterance indices from an atomic counter.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static double lse2(double a, double b) {
    double m = a > b ? a : b;
    if (m == -INFINITY) return -INFINITY;
    return m + log(exp(a - m) + exp(b - m));
}
static double lse3(double a, double b, double c) {
    double m = a > b ? a : b;
    if (c > m) m = c;
    if (m == -INFINITY) return -INFINITY;
    return m + log(exp(a - m) + exp(b - m) + exp(c - m));
}

typedef struct {
    const float *logits; const int64_t *targets; const int64_t *toff;
    const int64_t *in_len; const int64_t *tgt_len; int B, T, V, blank, zero_infinity;
    const double *scale; double *nll; double *grad; int next; int fail;
} job_t;

static void one_utterance(job_t *J, int b) {
    const int T = J->T, V = J->V, blank = J->blank, zero_infinity = J->zero_infinity;
    const int Tb = (int)J->in_len[b], U = (int)J->tgt_len[b], S = 2 * U + 1;
    const int64_t *y = J->targets + J->toff[b];
    const float *x = J->logits + (size_t)b * T * V;
    double *g = J->grad ? J->grad + (size_t)b * T * V : NULL;
    double *nll = J->nll;
    const double sc = J->scale ? J->scale[b] : 1.0;
    if (g) memset(g, 0, sizeof(double) * (size_t)T * V);
    if (Tb <= 0) { nll[b] = (U == 0) ? 0.0 : (zero_infinity ? 0.0 : INFINITY); return; }
    double *lse = (double *)malloc(sizeof(double) * (size_t)Tb);
    double *lpe = (double *)malloc(sizeof(double) * (size_t)Tb * S);
    double *al = (double *)malloc(sizeof(double) * (size_t)Tb * S);
    double *be = (double *)malloc(sizeof(double) * (size_t)Tb * S);
    if (!lse || !lpe || !al || !be) { J->fail = 1; free(lse); free(lpe); free(al); free(be); return; }
    for (int t = 0; t < Tb; ++t) {
        const float *r = x + (size_t)t * V;
        double m = r[0];
        for (int v = 1; v < V; ++v) if (r[v] > m) m = r[v];
        double s = 0.0;
        for (int v = 0; v < V; ++v) s += exp((double)r[v] - m);
        lse[t] = m + log(s);
        for (int k = 0; k < S; ++k) {
            int c = (k & 1) ? (int)y[k >> 1] : blank;
            lpe[(size_t)t * S + k] = (double)r[c] - lse[t];
        }
    }
    for (size_t i = 0; i < (size_t)Tb * S; ++i) al[i] = be[i] = -INFINITY;
    al[0] = lpe[0];
    if (S > 1) al[1] = lpe[1];
    for (int t = 1; t < Tb; ++t) {
        const double *p = al + (size_t)(t - 1) * S;
        double *a = al + (size_t)t * S;
        for (int k = 0; k < S; ++k) {
            double a0 = p[k], a1 = k >= 1 ? p[k - 1] : -INFINITY, a2 = -INFINITY;
            if ((k & 1) && k >= 3 && y[k >> 1] != y[(k >> 1) - 1]) a2 = p[k - 2];
            a[k] = lpe[(size_t)t * S + k] + lse3(a0, a1, a2);
        }
    }
    const double *aT = al + (size_t)(Tb - 1) * S;
    double ll = S > 1 ? lse2(aT[S - 1], aT[S - 2]) : aT[0];
    double *bT = be + (size_t)(Tb - 1) * S;
    bT[S - 1] = lpe[(size_t)(Tb - 1) * S + S - 1];
    if (S > 1) bT[S - 2] = lpe[(size_t)(Tb - 1) * S + S - 2];
    for (int t = Tb - 2; t >= 0; --t) {
        const double *n = be + (size_t)(t + 1) * S;
        double *bb = be + (size_t)t * S;
        for (int k = 0; k < S; ++k) {
            double b0 = n[k], b1 = k + 1 < S ? n[k + 1] : -INFINITY, b2 = -INFINITY;
            if ((k & 1) && k + 2 < S && y[k >> 1] != y[(k >> 1) + 1]) b2 = n[k + 2];
            bb[k] = lpe[(size_t)t * S + k] + lse3(b0, b1, b2);
        }
    }
    const int infeasible = (ll == -INFINITY);
    nll[b] = infeasible ? (zero_infinity ? 0.0 : INFINITY) : -ll;
    if (g && !(infeasible && zero_infinity)) {
        for (int t = 0; t < Tb; ++t) {
            const float *r = x + (size_t)t * V;
            double *gr = g + (size_t)t * V;
            if (infeasible) { for (int v = 0; v < V; ++v) gr[v] = NAN; continue; }
            for (int v = 0; v < V; ++v) gr[v] = sc * exp((double)r[v] - lse[t]);
            for (int k = 0; k < S; ++k) {
                double a = al[(size_t)t * S + k], bb = be[(size_t)t * S + k];
                if (a == -INFINITY || bb == -INFINITY) continue;
                int c = (k & 1) ? (int)y[k >> 1] : blank;
                gr[c] -= sc * exp(a + bb - lpe[(size_t)t * S + k] - ll);
            }
        }
    }
    free(lse); free(lpe); free(al); free(be);
}

static void *worker(void *arg) {
    job_t *J = (job_t *)arg;
    for (;;) {
        int b = __atomic_fetch_add(&J->next, 1, __ATOMIC_RELAXED);
        if (b >= J->B) break;
        one_utterance(J, b);
    }
    return NULL;
}

/* targets_stride == 0: 1-D concatenated targets; otherwise row stride of [B,Umax].
 * scale[b]: gradient scale of utterance b (grad_out * reduction factor); may be NULL (=1).
 * grad may be NULL (loss only).  nthreads <= 1: scalar.  Returns 0, or -1 on allocation failure. */
int ctc_oracle_f64(const float *logits, const int64_t *targets, int64_t targets_stride,
                   const int64_t *in_len, const int64_t *tgt_len, int B, int T, int V,
                   int blank, int zero_infinity, const double *scale, double *nll,
                   double *grad, int nthreads) {
    int64_t *toff = (int64_t *)malloc(sizeof(int64_t) * (size_t)(B + 1));
    if (!toff) return -1;
    toff[0] = 0;
    for (int b = 0; b < B; ++b)
        toff[b + 1] = toff[b] + (targets_stride ? targets_stride : tgt_len[b]);
    job_t J = {logits, targets, toff, in_len, tgt_len, B, T, V, blank, zero_infinity,
               scale, nll, grad, 0, 0};
    if (nthreads > B) nthreads = B;
    if (nthreads <= 1) {
        worker(&J);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
        int started = 0;
        if (th) {
            for (int i = 0; i < nthreads; ++i)
                if (pthread_create(&th[i], NULL, worker, &J) == 0) th[started++] = th[i];
        }
        worker(&J);
        for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    free(toff);
    return J.fail ? -1 : 0;
}
