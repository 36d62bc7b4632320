/* oracle/ibdgem_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the IBDGem scoring path, used as the parity checker for the CUDA
 * engine.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (ibdgem_b200/) never links or calls it.
 *
 * Parity status: PINNED — tests/test_oracle_golden.py checks this restatement against the 18
 * golden files the reference ships (supplementary/ibdgem-test/output) and against outputs of
 * the reference itself compiled into oracle/_ref (LD, -v, -D, -B, hiddengem; fixtures under
 * tests/golden/, generator tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 */
#ifndef IBDGEM_ORACLE_H
#define IBDGEM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- M1: binomial table, src/ibd-math.c:5-43 ------------------------------------------------ */
unsigned long **orc_init_nCk(unsigned int n);
unsigned long orc_retrieve_nCk(unsigned long **nCk, unsigned int n, unsigned int k);
int orc_destroy_nCk(unsigned long **nCk, unsigned int n);

/* ---- M2..M4: per-site closed forms, src/ibd-math.c:46-142 (same signatures as ibd-math.h) ---- */
double orc_find_pDgG(unsigned long **nCk, double epsilon, unsigned short A0, unsigned short A1,
                     unsigned int n_ref, unsigned int n_alt);
double orc_find_pDgf(double f, double pD_g_00, double pD_g_01, double pD_g_11);
double orc_find_pDgIBD1(unsigned short A0, unsigned short A1, double f, double pD_g_00,
                        double pD_g_01, double pD_g_11);

/* ---- A1: allele frequency over all 2N haplotypes, src/ibd-parse.c:91-99 ---------------------- */
double orc_find_f(const uint8_t *hap_row, int n_indiv);

/* Parameters of one comparison run (the file-scope statics of src/ibdgem.c:21-38). */
typedef struct {
    double epsilon;      /* -e, default 0.02 */
    uint32_t max_cov;    /* -M, default 20 */
    int32_t window;      /* -w, default 100 */
    double min_af;       /* -f */
    double max_af;       /* -F */
    int32_t ld_mode;     /* --LD */
    int32_t opt_v;       /* -v */
    double cull_p;       /* 1.0 = no down-sampling (-D), src/ibdgem.c:83-106 */
    int32_t pu_idx;      /* ordinal of the panel member named by -N, or -1 (src/ibdgem.c:501-506) */
} orc_params;

/* Result of comparing the pileup against ONE target (one iteration of the loop at
 * src/ibdgem.c:522).  All arrays are caller-allocated. */
typedef struct {
    /* per panel line (n_sites) */
    uint8_t *status;   /* 0 = skipped, 1 = processed & informative, 2 = processed, zero data */
    double *f;         /* allele frequency used (valid where status != 0) */
    uint8_t *n_ref;    /* counts after down-sampling */
    uint8_t *n_alt;
    double *ibd0;      /* per-site likelihoods exactly as the reference's doubles */
    double *ibd1;
    double *ibd2;
    /* per window (capacity max_windows) */
    int32_t n_windows;
    uint64_t *w_start; /* sgmt_start */
    uint64_t *w_end;   /* sgmt_end */
    int32_t *w_nsites; /* snp_count */
    double *w_lin;     /* [3*w]: the reference's linear fp64 values as printed (may underflow) */
    double *w_log;     /* [3*w]: natural-log of the same quantities, long-double log-space */
    /* counters, src/ibdgem.c:537-545, 761-768 */
    uint64_t processed;
    uint64_t skipped;
    uint64_t final_total_cov;
    uint64_t *final_dist; /* [max_cov+1] */
} orc_result;

/* Restatement of the per-target body of compare_impute, src/ibdgem.c:550-768, on pre-parsed
 * arrays.
 *   host_keep[i] : legend line parsed && is_snp && pileup line found && (-p) membership
 *                  (src/ibdgem.c:589-608) — everything that does not depend on the panel row.
 *   af_user[i]   : NaN, or the -A override found by bsearch (src/ibdgem.c:609-614); may be NULL.
 *   hap          : n_sites rows of 2*n_indiv bytes, each 0 or 1 (alleles of the .hap line).
 *   bg           : background individual ordinals (refids), length n_bg.
 *   target       : ordinal of the compared individual (cmp_idx/2).
 *   reseed       : if non-zero call srand(1) first (fresh-process rand() state for -D).
 */
int orc_compare_target(const orc_params *p, int64_t n_sites, int32_t n_indiv, const uint64_t *pos,
                       const uint8_t *host_keep, const double *af_user, const uint8_t *n_ref,
                       const uint8_t *n_alt, const uint8_t *hap, int32_t target,
                       const int32_t *bg, int32_t n_bg, int32_t max_windows, int reseed,
                       orc_result *out);

/* ---- H1..H3: hiddengem, src/hiddengem.c:51-147, 246-283 ------------------------------------- */
/* l[3*i+s] are the three likelihoods of bin i as parsed from the summary file.
 * Outputs: state[i] (0/1/2), score[3*i+s] = natural log of the long-double running product
 * (−inf for 0, NaN propagated), score_ld (optional, may be NULL) = the long double values. */
int orc_hiddengem(const double *l, int32_t n_bins, double p01, double p02, double p12,
                  int32_t *state, double *score_log, long double *score_ld);

/* Throughput helper for bench.py's cpu_baseline leg: runs the LD inner loop of
 * src/ibdgem.c:673-721 (linear fp64 products, single thread) over the given arrays for all
 * targets and returns the number of cross-term multiplies performed (site*target*bg*4). */
uint64_t orc_ld_loop_bench(const orc_params *p, int64_t n_sites, int32_t n_indiv,
                           const uint8_t *n_ref, const uint8_t *n_alt, const uint8_t *hap,
                           const int32_t *targets, int32_t n_targets, const int32_t *bg,
                           int32_t n_bg, double *sink);

#ifdef __cplusplus
}
#endif
#endif
