/* oracle/ibdgem_oracle.c — TEST INFRASTRUCTURE ONLY (see ibdgem_oracle.h).
 *
 * Plain-C restatement of the IBDGem scoring path.  It is written from the behaviour of the
 * reference (file:line citations relative to /root/reference), not from its text, and is
 * compiled with -ffp-contract=off -fno-builtin-pow so that every double operation rounds the
 * way the as-shipped (-O0, SSE2) reference binary does: same libm pow(), same evaluation
 * order, no fused multiply-add.
 *
 * Parity status: PINNED (tests/test_oracle_golden.py: 18 shipped golden files + outputs of
 * oracle/_ref for --LD, -v, -D, -B, -A, -F/-f, -M and hiddengem).
 */
#include "ibdgem_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* M1 — src/ibd-math.c:5-43.  The reference fills nCk[i][j] by the recursion
 * C(n,k) = (n * C(n-1,k-1)) / k in unsigned long with an unsigned-int n that wraps below zero.
 * Unrolled deepest-level-first this is v_0 = 1, v_j = ((n-k+j) * v_{j-1}) / j, j = 1..k, with
 * the factor (n-k+j) reduced modulo 2^32; for k > n one factor is 0, so the entry is 0. */
static unsigned long orc_binom(unsigned int n, unsigned int k) {
    unsigned long v = 1;
    unsigned int base = n - k; /* wraps for k > n, exactly like the recursion's n-1 chain */
    for (unsigned int j = 1; j <= k; j++) {
        unsigned int factor = base + j;
        v = (factor * v) / j;
    }
    return v;
}

unsigned long **orc_init_nCk(unsigned int n) {
    unsigned long **tab = malloc((size_t)(n + 1) * sizeof *tab);
    for (unsigned int i = 0; i <= n; i++) {
        tab[i] = malloc((size_t)(n + 1) * sizeof **tab);
        for (unsigned int j = 0; j <= n; j++) tab[i][j] = orc_binom(i, j);
    }
    return tab;
}

unsigned long orc_retrieve_nCk(unsigned long **nCk, unsigned int n, unsigned int k) {
    return nCk[n][k]; /* src/ibd-math.c:26-31 (the k>n warning has no numeric effect) */
}

int orc_destroy_nCk(unsigned long **nCk, unsigned int n) {
    if (!nCk) return 0;
    for (unsigned int i = 0; i <= n; i++) free(nCk[i]);
    free(nCk);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* M2 — src/ibd-math.c:46-81.  P(D|G): binomial with per-base error epsilon for the two
 * homozygotes, p = 1/2 for the heterozygote; (coefficient * first power) * second power. */
double orc_find_pDgG(unsigned long **nCk, double epsilon, unsigned short A0, unsigned short A1,
                     unsigned int n_ref, unsigned int n_alt) {
    if (n_ref == 0 && n_alt == 0) return 1.0; /* :50-53 */
    double coef = (double)orc_retrieve_nCk(nCk, n_ref + n_alt, n_ref);
    int g = (int)A0 + (int)A1;
    double p;
    if (A0 > 1 || A1 > 1) exit(1); /* :71-74 invalid genotype is fatal */
    if (g == 0) {
        p = coef * pow(1 - epsilon, n_ref);
        p = p * pow(epsilon, n_alt); /* :57-59 */
    } else if (g == 2) {
        p = coef * pow(1 - epsilon, n_alt);
        p = p * pow(epsilon, n_ref); /* :60-62 */
    } else {
        p = coef * pow(0.5, n_ref);
        p = p * pow(0.5, n_alt); /* :68-70 */
    }
    if (p == 0.0) p = DBL_MIN; /* :77-79 */
    return p;
}

/* M3 — src/ibd-math.c:84-101.  Hardy-Weinberg mixture; exactly 1 whenever any input is 1. */
double orc_find_pDgf(double f, double pD_g_00, double pD_g_01, double pD_g_11) {
    if (pD_g_00 == 1 || pD_g_01 == 1 || pD_g_11 == 1) return 1.0; /* :88-90 */
    double t0 = pow(1 - f, 2.0) * pD_g_00;
    double t1 = 2 * (1 - f) * f * pD_g_01; /* ((2*(1-f))*f)*P01 */
    double t2 = pow(f, 2.0) * pD_g_11;
    double p = (t0 + t1) + t2; /* :93-95 */
    if (p == 0.0) p = DBL_MIN;
    return p;
}

/* M4 — src/ibd-math.c:104-142.  One allele shared with the genotyped individual, the other
 * drawn at population frequency f.  Unknown alleles leave the value at 1. */
double orc_find_pDgIBD1(unsigned short A0, unsigned short A1, double f, double pD_g_00,
                        double pD_g_01, double pD_g_11) {
    double p = 1.0;
    if (A0 <= 1 && A1 <= 1) {
        int g = A0 + A1;
        if (g == 0) {
            p = (f * pD_g_01) + ((1 - f) * pD_g_00); /* :115-117 */
        } else if (g == 1) {
            double a = 0.5 * pD_g_01;
            double b = 0.5 * (1 - f) * pD_g_00;
            double c = 0.5 * f * pD_g_11;
            p = (a + b) + c; /* :122-126 */
        } else {
            p = ((1 - f) * pD_g_01) + (f * pD_g_11); /* :131-133 */
        }
    }
    if (p == 0.0) p = DBL_MIN; /* :138-140 */
    return p;
}

/* A1 — src/ibd-parse.c:91-99: count of '1' alleles over ALL 2N haplotypes / (2N). */
double orc_find_f(const uint8_t *hap_row, int n_indiv) {
    double n_alt = 0;
    for (int h = 0; h < 2 * n_indiv; h++)
        if (hap_row[h] == 1) n_alt++;
    return n_alt / (n_indiv * 2);
}

/* F3 — src/ibdgem.c:126-137: per-base Bernoulli thinning with the C library rand(). */
static unsigned int orc_cull(unsigned int count, double cull_p) {
    if (cull_p == 1.0) return count;
    unsigned int kept = 0;
    for (unsigned int b = 0; b < count; b++)
        if ((rand() / (double)RAND_MAX) < cull_p) kept++;
    return kept;
}

/* three-way select used all over src/ibdgem.c:643-713 */
static inline double orc_pick(int a, int b, const double P[3]) { return P[a + b]; }

static long double orc_lse(const long double *x, int n) {
    if (n <= 0) return -INFINITY;
    long double m = x[0];
    for (int i = 1; i < n; i++)
        if (x[i] > m) m = x[i];
    if (isinf(m)) return m;
    long double s = 0;
    for (int i = 0; i < n; i++) s += expl(x[i] - m);
    return m + logl(s);
}

/* ------------------------------------------------------------------------------------------ */
/* One target of compare_impute — src/ibdgem.c:550-768. */
int orc_compare_target(const orc_params *p, int64_t n_sites, int32_t n_indiv, const uint64_t *pos,
                       const uint8_t *host_keep, const double *af_user, const uint8_t *n_ref_in,
                       const uint8_t *n_alt_in, const uint8_t *hap, int32_t target,
                       const int32_t *bg, int32_t n_bg, int32_t max_windows, int reseed,
                       orc_result *out) {
    if (reseed) srand(1);
    const int H = 2 * n_indiv;
    unsigned long **nCk = orc_init_nCk(p->max_cov); /* src/ibdgem.c:1168 */
    const int ld = p->ld_mode;

    double *lin2 = NULL, *lin1 = NULL;       /* sum_ibd2_ref[n], sum_ibd1_ref[4n..] :563 */
    long double *log2v = NULL, *log1v = NULL; /* the same chains in log space */
    if (ld) {
        lin2 = malloc(sizeof(double) * (size_t)n_bg);
        lin1 = malloc(sizeof(double) * 4 * (size_t)n_bg);
        log2v = malloc(sizeof(long double) * (size_t)n_bg);
        log1v = malloc(sizeof(long double) * 4 * (size_t)n_bg);
    }

    out->processed = out->skipped = out->final_total_cov = 0;
    for (uint32_t c = 0; c <= p->max_cov; c++) out->final_dist[c] = 0;
    out->n_windows = 0;
    memset(out->status, 0, (size_t)n_sites);

    int64_t i = 0;
    int at_eof = 0;
    while (!at_eof) { /* :558 one iteration = one window */
        int snp_count = 0;
        uint64_t w_first = 0, w_last = 0, prev_pos = 0;
        double lin[3] = {1, 1, 1}; /* :562 */
        long double lg[3] = {0, 0, 0};
        if (ld)
            for (int n = 0; n < n_bg; n++) {
                lin2[n] = 1;
                log2v[n] = 0;
                for (int q = 0; q < 4; q++) {
                    lin1[4 * n + q] = 1;
                    log1v[4 * n + q] = 0;
                }
            }

        while (snp_count < p->window) { /* :572 */
            if (i >= n_sites) { /* :575-578 */
                at_eof = 1;
                w_last = prev_pos;
                break;
            }
            const int64_t s = i++;
            const uint8_t *row = hap + (size_t)s * H;
            const int A0 = row[2 * target], A1 = row[2 * target + 1];
            if (p->opt_v && A0 == 0 && A1 == 0) { /* :584-587 */
                out->skipped++;
                continue;
            }
            if (!host_keep[s]) { /* :589-608 */
                out->skipped++;
                continue;
            }
            double f = orc_find_f(row, n_indiv);                   /* :609 */
            if (af_user && !isnan(af_user[s])) f = af_user[s];      /* :610-615 */
            if (f > p->max_af || f < p->min_af) {                   /* :616-620 */
                out->skipped++;
                continue;
            }
            unsigned int nr = n_ref_in[s], na = n_alt_in[s];
            if (nr + na > p->max_cov) { /* :623-626 */
                out->skipped++;
                continue;
            }
            nr = orc_cull(nr, p->cull_p); /* :627-628, order matters for rand() */
            na = orc_cull(na, p->cull_p);
            out->final_total_cov += nr + na;
            out->final_dist[nr + na]++;

            double P[3];
            P[0] = orc_find_pDgG(nCk, p->epsilon, 0, 0, nr, na); /* :632-634 */
            P[1] = orc_find_pDgG(nCk, p->epsilon, 0, 1, nr, na);
            P[2] = orc_find_pDgG(nCk, p->epsilon, 1, 1, nr, na);
            double v0 = orc_find_pDgf(f, P[0], P[1], P[2]);                /* :641 */
            double v1 = orc_find_pDgIBD1(A0, A1, f, P[0], P[1], P[2]);     /* :642 */
            double v2 = (A0 <= 1 && A1 <= 1) ? orc_pick(A0, A1, P) : 1.0;  /* :643-651 */

            out->f[s] = f;
            out->n_ref[s] = (uint8_t)nr;
            out->n_alt[s] = (uint8_t)na;
            out->ibd0[s] = v0;
            out->ibd1[s] = v1;
            out->ibd2[s] = v2;
            out->processed++;
            if (nr + na < 1) { /* :657-663 zero-data bypass */
                out->status[s] = 2;
                continue;
            }
            out->status[s] = 1;

            lin[0] *= v0; /* :665-667 */
            lin[1] *= v1;
            lin[2] *= v2;
            lg[0] += logl(v0);
            lg[1] += logl(v1);
            lg[2] += logl(v2);

            if (ld) { /* :673-721 */
                for (int n = 0; n < n_bg; n++) {
                    const int b = bg[n];
                    if (b == p->pu_idx || b == target) continue; /* :714 */
                    const int r0 = row[2 * b], r1 = row[2 * b + 1];
                    const double e = orc_pick(r0, r1, P);
                    const double x00 = orc_pick(A0, r0, P), x01 = orc_pick(A0, r1, P);
                    const double x10 = orc_pick(A1, r0, P), x11 = orc_pick(A1, r1, P);
                    lin2[n] *= e;
                    lin1[4 * n + 0] *= x00;
                    lin1[4 * n + 1] *= x01;
                    lin1[4 * n + 2] *= x10;
                    lin1[4 * n + 3] *= x11;
                    log2v[n] += logl(e);
                    log1v[4 * n + 0] += logl(x00);
                    log1v[4 * n + 1] += logl(x01);
                    log1v[4 * n + 2] += logl(x10);
                    log1v[4 * n + 3] += logl(x11);
                }
            }
            snp_count++; /* :723-730 */
            prev_pos = pos[s];
            if (snp_count == 1) w_first = pos[s];
            if (snp_count == p->window) w_last = prev_pos;
        }

        if (snp_count > 0) { /* :736-759 */
            const int w = out->n_windows;
            if (w >= max_windows) {
                orc_destroy_nCk(nCk, p->max_cov);
                free(lin2); free(lin1); free(log2v); free(log1v);
                return 2;
            }
            out->w_start[w] = w_first;
            out->w_end[w] = w_last;
            out->w_nsites[w] = snp_count;
            if (ld) {
                int n_refpanel = n_bg;
                double s0 = 0, s1 = 0;
                long double *x0 = malloc(sizeof(long double) * (size_t)(n_bg ? n_bg : 1));
                long double *x1 = malloc(sizeof(long double) * 4 * (size_t)(n_bg ? n_bg : 1));
                int m = 0;
                for (int n = 0; n < n_bg; n++) {
                    if (bg[n] != p->pu_idx && bg[n] != target) { /* :742 */
                        s0 += lin2[n];
                        s1 += (lin1[4 * n] + lin1[4 * n + 1] + lin1[4 * n + 2] + lin1[4 * n + 3]);
                        x0[m] = log2v[n];
                        for (int q = 0; q < 4; q++) x1[4 * m + q] = log1v[4 * n + q];
                        m++;
                    } else {
                        n_refpanel--;
                    }
                }
                out->w_lin[3 * w + 0] = s0 / n_refpanel;       /* :751-752 */
                out->w_lin[3 * w + 1] = s1 / (n_refpanel * 4);
                out->w_lin[3 * w + 2] = lin[2];
                if (m > 0) {
                    out->w_log[3 * w + 0] = (double)(orc_lse(x0, m) - logl((long double)m));
                    out->w_log[3 * w + 1] = (double)(orc_lse(x1, 4 * m) - logl(4.0L * m));
                } else {
                    out->w_log[3 * w + 0] = NAN;
                    out->w_log[3 * w + 1] = NAN;
                }
                out->w_log[3 * w + 2] = (double)lg[2];
                free(x0);
                free(x1);
            } else {
                for (int k = 0; k < 3; k++) {
                    out->w_lin[3 * w + k] = lin[k]; /* :755-756 */
                    out->w_log[3 * w + k] = (double)lg[k];
                }
            }
            out->n_windows = w + 1;
        }
    }
    orc_destroy_nCk(nCk, p->max_cov);
    free(lin2); free(lin1); free(log2v); free(log1v);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* H1-H3 — src/hiddengem.c:51-147, 246-257.  Three-state Viterbi over normalised window
 * likelihoods, running products in x87 long double, strict-'>' argmax (lowest index wins ties
 * and NaNs). */
static int orc_argmax3(const long double v[3]) { /* src/hiddengem.c:91-99 */
    int best = 0;
    for (int k = 1; k < 3; k++)
        if (v[k] > v[best]) best = k;
    return best;
}

int orc_hiddengem(const double *l, int32_t n_bins, double p01, double p02, double p12,
                  int32_t *state, double *score_log, long double *score_ld) {
    if (n_bins <= 0) return 1;
    const size_t n = (size_t)n_bins;
    long double *sc = malloc(sizeof(long double) * 3 * n);
    uint8_t *from = malloc(3 * n);
    double pen[3][3] = {{1, p01, p02}, {p01, 1, p12}, {p02, p12, 1}}; /* :13-15, 122-137 */

    for (size_t i = 0; i < n; i++) {
        const double tot = l[3 * i] + l[3 * i + 1] + l[3 * i + 2]; /* :74-76 */
        double nrm[3];
        for (int s = 0; s < 3; s++) nrm[s] = l[3 * i + s] / tot;
        for (int s = 0; s < 3; s++) {
            if (i == 0) { /* :112-118 */
                sc[s] = nrm[s];
                from[s] = (uint8_t)s;
                continue;
            }
            long double cand[3];
            for (int k = 0; k < 3; k++) {
                /* (score * nrm) * penalty, the diagonal has no third factor :122-137 */
                long double c = sc[3 * (i - 1) + k] * nrm[s];
                if (k != s) c = c * pen[k][s];
                cand[k] = c;
            }
            const int k = orc_argmax3(cand);
            sc[3 * i + s] = cand[k];
            from[3 * i + s] = (uint8_t)k;
        }
    }
    long double last[3] = {sc[3 * (n - 1)], sc[3 * (n - 1) + 1], sc[3 * (n - 1) + 2]};
    int cur = orc_argmax3(last); /* :246-249 */
    for (size_t i = n; i-- > 0;) { /* :252-257 */
        state[i] = cur;
        cur = from[3 * i + cur];
    }
    for (size_t j = 0; j < 3 * n; j++) {
        if (score_ld) score_ld[j] = sc[j];
        score_log[j] = (double)logl(sc[j]);
    }
    free(sc);
    free(from);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* CPU-baseline helper: the LD inner loop of src/ibdgem.c:673-721 as the reference runs it
 * (linear fp64 products, per-site P(D|G) from the table, one thread), without text I/O. */
uint64_t orc_ld_loop_bench(const orc_params *p, int64_t n_sites, int32_t n_indiv,
                           const uint8_t *n_ref, const uint8_t *n_alt, const uint8_t *hap,
                           const int32_t *targets, int32_t n_targets, const int32_t *bg,
                           int32_t n_bg, double *sink) {
    const int H = 2 * n_indiv;
    unsigned long **nCk = orc_init_nCk(p->max_cov);
    double *acc = malloc(sizeof(double) * 5 * (size_t)n_bg);
    uint64_t mults = 0;
    double total = 0;
    for (int t = 0; t < n_targets; t++) {
        const int tg = targets[t];
        for (int n = 0; n < 5 * n_bg; n++) acc[n] = 1;
        int in_window = 0;
        for (int64_t s = 0; s < n_sites; s++) {
            const unsigned nr = n_ref[s], na = n_alt[s];
            if (nr + na < 1 || nr + na > p->max_cov) continue;
            const uint8_t *row = hap + (size_t)s * H;
            double P[3];
            P[0] = orc_find_pDgG(nCk, p->epsilon, 0, 0, nr, na);
            P[1] = orc_find_pDgG(nCk, p->epsilon, 0, 1, nr, na);
            P[2] = orc_find_pDgG(nCk, p->epsilon, 1, 1, nr, na);
            const int A0 = row[2 * tg], A1 = row[2 * tg + 1];
            for (int n = 0; n < n_bg; n++) {
                const int b = bg[n];
                if (b == p->pu_idx || b == tg) continue;
                const int r0 = row[2 * b], r1 = row[2 * b + 1];
                acc[5 * n] *= P[r0 + r1];
                acc[5 * n + 1] *= P[A0 + r0];
                acc[5 * n + 2] *= P[A0 + r1];
                acc[5 * n + 3] *= P[A1 + r0];
                acc[5 * n + 4] *= P[A1 + r1];
                mults += 4;
            }
            if (++in_window == p->window) {
                for (int n = 0; n < 5 * n_bg; n++) {
                    total += acc[n];
                    acc[n] = 1;
                }
                in_window = 0;
            }
        }
        for (int n = 0; n < 5 * n_bg; n++) total += acc[n];
    }
    if (sink) *sink = total;
    free(acc);
    orc_destroy_nCk(nCk, p->max_cov);
    return mults;
}
