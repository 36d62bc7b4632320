// hidden.cu — batched hiddengem: three-state Viterbi over per-window likelihoods (H1-H3,
// src/hiddengem.c:51-147, 246-283), one summary table per thread, in fp64 log space.
//
// The reference keeps running PRODUCTS of normalised likelihoods in x87 long double
// (src/hiddengem.c:43, 110-141); here the same recurrence runs on log values, so
//   score[s][i] = max_k ( (score[k][i-1] + ln nrm[s][i]) + ln pen(k,s) )
// with the reference's strict-'>' argmax (lowest index wins ties, NaN never wins,
// src/hiddengem.c:91-99).  Zeros become -inf and NaN rows (0/0, src/hiddengem.c:74-76)
// propagate exactly as they do through the products.
#include <math.h>

#include "engine.h"

namespace ibdgem {

__device__ __forceinline__ int argmax3(double c0, double c1, double c2) {
    int b = 0;
    double v = c0;
    if (c1 > v) { b = 1; v = c1; }
    if (c2 > v) { b = 2; }
    return b;
}

__global__ void __launch_bounds__(128)
viterbi_kernel(int n_tables, const int64_t *__restrict__ off, const double *__restrict__ lik, int is_log,
               double lp01, double lp02, double lp12, uint8_t *__restrict__ state,
               double *__restrict__ score, long long *__restrict__ counts) {
    const int tb = blockIdx.x * blockDim.x + threadIdx.x;
    if (tb >= n_tables) return;
    const int64_t b0 = off[tb], b1 = off[tb + 1];
    const int64_t n = b1 - b0;
    if (n <= 0) {
        counts[tb * 3 + 0] = counts[tb * 3 + 1] = counts[tb * 3 + 2] = 0;
        return;
    }
    double s0 = 0, s1 = 0, s2 = 0;
    for (int64_t i = 0; i < n; i++) {
        const double *L = lik + (b0 + i) * 3;
        double n0, n1, n2;
        if (is_log) {  // ln nrm = ll - logsumexp(ll)
            const double m = fmax(L[0], fmax(L[1], L[2]));
            if (m == -INFINITY) {
                n0 = n1 = n2 = __longlong_as_double(0x7ff8000000000000LL);  // 0/0 in the reference
            } else {
                const double z = m + log(exp(L[0] - m) + exp(L[1] - m) + exp(L[2] - m));
                n0 = L[0] - z; n1 = L[1] - z; n2 = L[2] - z;
            }
        } else {  // src/hiddengem.c:74-76: l_s / (l0 + l1 + l2) in fp64, then the log
            const double tot = __dadd_rn(__dadd_rn(L[0], L[1]), L[2]);
            n0 = log(__ddiv_rn(L[0], tot));
            n1 = log(__ddiv_rn(L[1], tot));
            n2 = log(__ddiv_rn(L[2], tot));
        }
        uint8_t from;
        if (i == 0) {  // src/hiddengem.c:112-118
            s0 = n0; s1 = n1; s2 = n2;
            from = 0 | (1 << 2) | (2 << 4);
        } else {
            // candidates (prev_k + ln nrm_s) + ln pen(k, s), diagonal has no penalty factor
            const int k0 = argmax3(s0 + n0, (s1 + n0) + lp01, (s2 + n0) + lp02);
            const int k1 = argmax3((s0 + n1) + lp01, s1 + n1, (s2 + n1) + lp12);
            const int k2 = argmax3((s0 + n2) + lp02, (s1 + n2) + lp12, s2 + n2);
            const double p[3] = {s0, s1, s2};
            const double t0 = (k0 == 0) ? p[0] + n0 : (p[k0] + n0) + (k0 == 1 ? lp01 : lp02);
            const double t1 = (k1 == 1) ? p[1] + n1 : (p[k1] + n1) + (k1 == 0 ? lp01 : lp12);
            const double t2 = (k2 == 2) ? p[2] + n2 : (p[k2] + n2) + (k2 == 0 ? lp02 : lp12);
            s0 = t0; s1 = t1; s2 = t2;
            from = (uint8_t)(k0 | (k1 << 2) | (k2 << 4));
        }
        state[b0 + i] = from;
        double *o = score + (b0 + i) * 3;
        o[0] = s0; o[1] = s1; o[2] = s2;
    }
    int cur = argmax3(s0, s1, s2);  // src/hiddengem.c:246-249
    long long c[3] = {0, 0, 0};
    for (int64_t i = n - 1; i >= 0; i--) {  // src/hiddengem.c:252-257
        const uint8_t from = state[b0 + i];
        state[b0 + i] = (uint8_t)cur;
        c[cur]++;
        cur = (from >> (2 * cur)) & 3;
    }
    counts[tb * 3 + 0] = c[0];
    counts[tb * 3 + 1] = c[1];
    counts[tb * 3 + 2] = c[2];
}

}  // namespace ibdgem

using namespace ibdgem;

extern "C" int hiddengem_viterbi_batch(ibdgem_engine *e, int32_t n_tables, const int64_t *bin_offsets,
                                       const double *lik, int32_t is_log, double p01, double p02,
                                       double p12, uint8_t *state, double *score_log,
                                       int64_t *state_counts) {
    if (!e || n_tables <= 0 || !bin_offsets || !lik) {
        set_error("[::] ERROR in hiddengem_viterbi_batch(): bad arguments.");
        return 1;
    }
    IBD_CUDA(cudaSetDevice(e->device));
    const int64_t nb = bin_offsets[n_tables];
    if (nb <= 0) {
        set_error("[::] ERROR parsing likelihood data; make sure input is valid.");
        return 1;
    }
    double *d_lik, *d_score;
    int64_t *d_off;
    uint8_t *d_state;
    long long *d_counts;
    if (scratch(e, SC_HG_LIK, (size_t)nb * 24, (void **)&d_lik) || scratch(e, SC_HG_OFF, (size_t)(n_tables + 1) * 8, (void **)&d_off) ||
        scratch(e, SC_HG_STATE, (size_t)nb, (void **)&d_state) || scratch(e, SC_HG_SCORE, (size_t)nb * 24, (void **)&d_score) ||
        scratch(e, SC_HG_COUNTS, (size_t)n_tables * 24, (void **)&d_counts))
        return 1;
    IBD_CUDA(cudaMemcpyAsync(d_lik, lik, (size_t)nb * 24, cudaMemcpyHostToDevice, e->stream));
    IBD_CUDA(cudaMemcpyAsync(d_off, bin_offsets, (size_t)(n_tables + 1) * 8, cudaMemcpyHostToDevice, e->stream));
    {
        LaunchScope ls(e, K_VITERBI);
        viterbi_kernel<<<(n_tables + 127) / 128, 128, 0, e->stream>>>(n_tables, d_off, d_lik, is_log, log(p01), log(p02),
                                                                     log(p12), d_state, d_score, d_counts);
    }
    IBD_CUDA(cudaGetLastError());
    if (state) IBD_CUDA(cudaMemcpyAsync(state, d_state, (size_t)nb, cudaMemcpyDeviceToHost, e->stream));
    if (score_log) IBD_CUDA(cudaMemcpyAsync(score_log, d_score, (size_t)nb * 24, cudaMemcpyDeviceToHost, e->stream));
    if (state_counts) IBD_CUDA(cudaMemcpyAsync(state_counts, d_counts, (size_t)n_tables * 24, cudaMemcpyDeviceToHost, e->stream));
    IBD_CUDA(cudaStreamSynchronize(e->stream));
    resolve_timers(e);
    return 0;
}
