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

#include <algorithm>

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

// ---------------------------------------------------------------------------------------------
// Batched pipeline for many tables (C4: 10,000 tables x 10,000 bins).  The recursion is sequential
// in bins and independent across tables, so the natural thread is a table — but the caller's
// arrays are table-major and a thread-per-table walk over them is uncoalesced.  Four kernels, all
// streaming: (1) ln nrm per bin, transposed to bin-major through shared memory; (2) forward pass,
// thread = table, every load and store coalesced across tables, loads issued UNROLL bins ahead of
// the recurrence; (3) back-trace, same layout; (4) scores and states transposed back to the
// caller's table-major layout.  Arithmetic is operation-for-operation that of viterbi_kernel.
__device__ __forceinline__ void ln_nrm(const double *L, int is_log, double &n0, double &n1, double &n2) {
    if (is_log) {
        const double m = fmax(L[0], fmax(L[1], L[2]));
        if (m == -INFINITY) {
            n0 = n1 = n2 = __longlong_as_double(0x7ff8000000000000LL);
        } else {
            const double z = m + log(exp(L[0] - m) + exp(L[1] - m) + exp(L[2] - m));
            n0 = L[0] - z; n1 = L[1] - z; n2 = L[2] - z;
        }
    } else {
        const double tot = __dadd_rn(__dadd_rn(L[0], L[1]), L[2]);
        n0 = log(__ddiv_rn(L[0], tot));
        n1 = log(__ddiv_rn(L[1], tot));
        n2 = log(__ddiv_rn(L[2], tot));
    }
}

// (1) block = 32 tables x 32 bins.  Reads are contiguous runs of 32 bins of one table; writes are
// runs of 32 tables of one (bin, state).
__global__ void __launch_bounds__(1024)
viterbi_norm_kernel(int n_tables, int64_t maxbins, const int64_t *__restrict__ off, const double *__restrict__ lik,
                    int is_log, double *__restrict__ nrmT /*[maxbins][3][n_tables]*/) {
    __shared__ double tile[3][32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int t_in = blockIdx.x * 32 + ty;
    const int64_t i_in = (int64_t)blockIdx.y * 32 + tx;
    double n0 = 0, n1 = 0, n2 = 0;
    if (t_in < n_tables) {
        const int64_t b0 = off[t_in], n = off[t_in + 1] - b0;
        if (i_in < n) {
            const double *L = lik + (b0 + i_in) * 3;
            const double v[3] = {L[0], L[1], L[2]};
            ln_nrm(v, is_log, n0, n1, n2);
        }
    }
    tile[0][tx][ty] = n0;
    tile[1][tx][ty] = n1;
    tile[2][tx][ty] = n2;
    __syncthreads();
    const int t_out = blockIdx.x * 32 + tx;
    const int64_t i_out = (int64_t)blockIdx.y * 32 + ty;
    if (t_out < n_tables && i_out < maxbins) {
#pragma unroll
        for (int s = 0; s < 3; s++) nrmT[(i_out * 3 + s) * n_tables + t_out] = tile[s][ty][tx];
    }
}

// (2) thread = table
constexpr int VIT_UNROLL = 8;
__global__ void __launch_bounds__(64)
viterbi_forward_kernel(int n_tables, const int64_t *__restrict__ off, const double *__restrict__ nrmT, double lp01,
                       double lp02, double lp12, uint8_t *__restrict__ fromT /*[maxbins][n_tables]*/,
                       double *__restrict__ scoreT /*[maxbins][3][n_tables]*/, double *__restrict__ last /*[n_tables][3]*/) {
    const int tb = blockIdx.x * blockDim.x + threadIdx.x;
    if (tb >= n_tables) return;
    const int64_t n = off[tb + 1] - off[tb];
    const size_t nT = (size_t)n_tables;
    double s0 = 0, s1 = 0, s2 = 0;
    // two register batches: the loads of batch k + 1 are in flight while batch k runs the recurrence
    double nx[VIT_UNROLL][3];
    auto fetch = [&](int64_t ib) {
#pragma unroll
        for (int u = 0; u < VIT_UNROLL; u++)
            if (ib + u < n) {
#pragma unroll
                for (int s = 0; s < 3; s++) nx[u][s] = __ldcs(nrmT + ((size_t)(ib + u) * 3 + s) * nT + tb);
            }
    };
    fetch(0);
    for (int64_t ib = 0; ib < n; ib += VIT_UNROLL) {
        double v[VIT_UNROLL][3];
#pragma unroll
        for (int u = 0; u < VIT_UNROLL; u++) {
            v[u][0] = nx[u][0]; v[u][1] = nx[u][1]; v[u][2] = nx[u][2];
        }
        if (ib + VIT_UNROLL < n) fetch(ib + VIT_UNROLL);
#pragma unroll
        for (int u = 0; u < VIT_UNROLL; u++) {
            const int64_t i = ib + u;
            if (i >= n) break;
            const double n0 = v[u][0], n1 = v[u][1], n2 = v[u][2];
            uint8_t from;
            if (i == 0) {
                s0 = n0; s1 = n1; s2 = n2;
                from = 0 | (1 << 2) | (2 << 4);
            } else {
                const int k0 = argmax3(s0 + n0, (s1 + n0) + lp01, (s2 + n0) + lp02);
                const int k1 = argmax3((s0 + n1) + lp01, s1 + n1, (s2 + n1) + lp12);
                const int k2 = argmax3((s0 + n2) + lp02, (s1 + n2) + lp12, s2 + n2);
                const double p0 = s0, p1 = s1, p2 = s2;
                const double q0 = k0 == 0 ? p0 : (k0 == 1 ? p1 : p2), q1 = k1 == 0 ? p0 : (k1 == 1 ? p1 : p2),
                             q2 = k2 == 0 ? p0 : (k2 == 1 ? p1 : p2);
                s0 = (k0 == 0) ? q0 + n0 : (q0 + n0) + (k0 == 1 ? lp01 : lp02);
                s1 = (k1 == 1) ? q1 + n1 : (q1 + n1) + (k1 == 0 ? lp01 : lp12);
                s2 = (k2 == 2) ? q2 + n2 : (q2 + n2) + (k2 == 0 ? lp02 : lp12);
                from = (uint8_t)(k0 | (k1 << 2) | (k2 << 4));
            }
            fromT[(size_t)i * nT + tb] = from;
            __stcs(scoreT + ((size_t)i * 3 + 0) * nT + tb, s0);
            __stcs(scoreT + ((size_t)i * 3 + 1) * nT + tb, s1);
            __stcs(scoreT + ((size_t)i * 3 + 2) * nT + tb, s2);
        }
    }
    last[(size_t)tb * 3 + 0] = s0;
    last[(size_t)tb * 3 + 1] = s1;
    last[(size_t)tb * 3 + 2] = s2;
}

// (3) back-trace; the state overwrites the back-pointer byte in place
__global__ void __launch_bounds__(64)
viterbi_back_kernel(int n_tables, const int64_t *__restrict__ off, uint8_t *__restrict__ fromT, const double *__restrict__ last,
                    long long *__restrict__ counts) {
    const int tb = blockIdx.x * blockDim.x + threadIdx.x;
    if (tb >= n_tables) return;
    const int64_t n = off[tb + 1] - off[tb];
    const size_t nT = (size_t)n_tables;
    long long c0 = 0, c1 = 0, c2 = 0;
    if (n > 0) {
        int cur = argmax3(last[(size_t)tb * 3], last[(size_t)tb * 3 + 1], last[(size_t)tb * 3 + 2]);
        constexpr int BU = 32;
        for (int64_t ib = n; ib > 0; ib -= BU) {
            uint8_t f[BU];
#pragma unroll
            for (int u = 0; u < BU; u++)
                if (ib - 1 - u >= 0) f[u] = fromT[(size_t)(ib - 1 - u) * nT + tb];
#pragma unroll
            for (int u = 0; u < BU; u++) {
                const int64_t i = ib - 1 - u;
                if (i < 0) break;
                fromT[(size_t)i * nT + tb] = (uint8_t)cur;
                c0 += cur == 0; c1 += cur == 1; c2 += cur == 2;
                cur = (f[u] >> (2 * cur)) & 3;
            }
        }
    }
    counts[tb * 3 + 0] = c0;
    counts[tb * 3 + 1] = c1;
    counts[tb * 3 + 2] = c2;
}

// (4) bin-major -> the caller's table-major layout
__global__ void __launch_bounds__(1024)
viterbi_out_kernel(int n_tables, const int64_t *__restrict__ off, const uint8_t *__restrict__ stateT,
                   const double *__restrict__ scoreT, uint8_t *__restrict__ state, double *__restrict__ score) {
    __shared__ double tile[3][32][33];
    __shared__ uint8_t st[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int t_in = blockIdx.x * 32 + tx;
    const int64_t i_in = (int64_t)blockIdx.y * 32 + ty;
    const size_t nT = (size_t)n_tables;
    if (t_in < n_tables && i_in < off[t_in + 1] - off[t_in]) {
#pragma unroll
        for (int s = 0; s < 3; s++) tile[s][ty][tx] = __ldcs(scoreT + ((size_t)i_in * 3 + s) * nT + t_in);
        st[ty][tx] = stateT[(size_t)i_in * nT + t_in];
    }
    __syncthreads();
    const int t_out = blockIdx.x * 32 + ty;
    const int64_t i_out = (int64_t)blockIdx.y * 32 + tx;
    if (t_out < n_tables) {
        const int64_t b0 = off[t_out], n = off[t_out + 1] - b0;
        if (i_out < n) {
            double *o = score + (b0 + i_out) * 3;
            o[0] = tile[0][tx][ty];
            o[1] = tile[1][tx][ty];
            o[2] = tile[2][tx][ty];
            state[b0 + i_out] = st[tx][ty];
        }
    }
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
    int64_t maxbins = 0;
    for (int t = 0; t < n_tables; t++) maxbins = std::max<int64_t>(maxbins, bin_offsets[t + 1] - bin_offsets[t]);
    const bool batched = n_tables >= 64 && (double)maxbins * n_tables <= 2.0 * (double)nb;
    if (batched) {
        double *d_nrmT, *d_scoreT, *d_last;
        uint8_t *d_fromT;
        const size_t cells = (size_t)maxbins * n_tables;
        if (scratch(e, SC_HG_NRMT, cells * 24, (void **)&d_nrmT) || scratch(e, SC_HG_SCORET, cells * 24, (void **)&d_scoreT) ||
            scratch(e, SC_HG_FROMT, cells, (void **)&d_fromT) || scratch(e, SC_HG_LAST, (size_t)n_tables * 24, (void **)&d_last))
            return 1;
        const dim3 tiles((unsigned)((n_tables + 31) / 32), (unsigned)((maxbins + 31) / 32));
        {
            LaunchScope ls(e, K_VITERBI_NORM);
            viterbi_norm_kernel<<<tiles, 1024, 0, e->stream>>>(n_tables, maxbins, d_off, d_lik, is_log, d_nrmT);
        }
        {
            LaunchScope ls(e, K_VITERBI);
            viterbi_forward_kernel<<<(n_tables + 63) / 64, 64, 0, e->stream>>>(n_tables, d_off, d_nrmT, log(p01), log(p02), log(p12),
                                                                             d_fromT, d_scoreT, d_last);
        }
        {
            LaunchScope ls(e, K_VITERBI_BACK);
            viterbi_back_kernel<<<(n_tables + 63) / 64, 64, 0, e->stream>>>(n_tables, d_off, d_fromT, d_last, d_counts);
        }
        {
            LaunchScope ls(e, K_VITERBI_OUT);
            viterbi_out_kernel<<<tiles, 1024, 0, e->stream>>>(n_tables, d_off, d_fromT, d_scoreT, d_state, d_score);
        }
    } else {
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
