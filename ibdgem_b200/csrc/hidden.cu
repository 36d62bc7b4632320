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
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "engine.h"
#include "tc_common.cuh"

namespace ibdgem {

__device__ __forceinline__ int argmax3(double c0, double c1, double c2) {
    int b = 0;
    double v = c0;
    if (c1 > v) { b = 1; v = c1; }
    if (c2 > v) { b = 2; }
    return b;
}

// Near-tie guard.  The reference multiplies x87 long doubles (src/hiddengem.c:110-141); sums of fp64 logs carry a
// rounding error of up to ~i * 2^-53 * |score| after i bins, so an arg-max whose winner leads the runner-up by less
// than that could come out differently.  Such tables (and, on the text path, tables whose scores leave the range
// of a normal long double, where the reference's products lose bits and then vanish) are flagged and re-evaluated
// on the host in long double.  NaN / -inf candidates never flag: their comparisons are exact in both worlds.
constexpr double LN_LDBL_MIN = -11355.137111933024;  // ln(3.3621e-4932)
__device__ __forceinline__ double tie_scale(int64_t i) { return (double)(int)(i + 1) * 4.440892098500626e-16; }  // 4 (i + 1) 2^-53
__device__ __forceinline__ bool near_tie(double c0, double c1, double c2, int best, double scale) {
    const double w = best == 0 ? c0 : (best == 1 ? c1 : c2);
    const double r = best == 0 ? fmax(c1, c2) : (best == 1 ? fmax(c0, c2) : fmax(c0, c1));
    return (w - r) < fma(scale, fabs(w), 1e-12);  // false for NaN and for inf - inf
}
__device__ __forceinline__ bool near_tie(double c0, double c1, double c2, int best, int64_t i) {
    return near_tie(c0, c1, c2, best, tie_scale(i));
}
// a finite score below the smallest normal long double: the reference's product is a denormal (or zero) there
__device__ __forceinline__ bool below_ldbl(double s0, double s1, double s2) {
    return (s0 < LN_LDBL_MIN && s0 > -INFINITY) || (s1 < LN_LDBL_MIN && s1 > -INFINITY) || (s2 < LN_LDBL_MIN && s2 > -INFINITY);
}

__global__ void __launch_bounds__(128)
viterbi_kernel(int n_tables, const int64_t *__restrict__ start, const int64_t *__restrict__ len, const double *__restrict__ lik, int is_log,
               double lp01, double lp02, double lp12, uint8_t *__restrict__ state,
               double *__restrict__ score, long long *__restrict__ counts, uint8_t *__restrict__ flag) {
    const int tb = blockIdx.x * blockDim.x + threadIdx.x;
    if (tb >= n_tables) return;
    const int64_t b0 = start[tb];
    const int64_t n = len[tb];
    if (n <= 0) {
        counts[tb * 3 + 0] = counts[tb * 3 + 1] = counts[tb * 3 + 2] = 0;
        flag[tb] = 0;
        return;
    }
    double s0 = 0, s1 = 0, s2 = 0;
    bool tie = false;
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
            const double a0 = s0 + n0, a1 = (s1 + n0) + lp01, a2 = (s2 + n0) + lp02;
            const double b0_ = (s0 + n1) + lp01, b1 = s1 + n1, b2 = (s2 + n1) + lp12;
            const double c0 = (s0 + n2) + lp02, c1 = (s1 + n2) + lp12, c2 = s2 + n2;
            const int k0 = argmax3(a0, a1, a2), k1 = argmax3(b0_, b1, b2), k2 = argmax3(c0, c1, c2);
            const double ts = tie_scale(i);
            tie |= near_tie(a0, a1, a2, k0, ts) | near_tie(b0_, b1, b2, k1, ts) | near_tie(c0, c1, c2, k2, ts);
            s0 = k0 == 0 ? a0 : (k0 == 1 ? a1 : a2);
            s1 = k1 == 0 ? b0_ : (k1 == 1 ? b1 : b2);
            s2 = k2 == 0 ? c0 : (k2 == 1 ? c1 : c2);
            from = (uint8_t)(k0 | (k1 << 2) | (k2 << 4));
        }
        if (!is_log) tie |= below_ldbl(s0, s1, s2);
        state[b0 + i] = from;
        double *o = score + (b0 + i) * 3;
        o[0] = s0; o[1] = s1; o[2] = s2;
    }
    int cur = argmax3(s0, s1, s2);  // src/hiddengem.c:246-249
    tie |= near_tie(s0, s1, s2, cur, n);
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
    flag[tb] = tie ? 1 : 0;
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
        } else {  // exp_nonpos: arguments are <= 0, degree-12 polynomial, error ~1 ulp (half the instructions of exp())
            const double z = m + log(exp_nonpos(L[0] - m) + exp_nonpos(L[1] - m) + exp_nonpos(L[2] - m));
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
viterbi_norm_kernel(int n_tables, int64_t maxbins, const int64_t *__restrict__ start, const int64_t *__restrict__ len,
                    const double *__restrict__ lik, int is_log, double *__restrict__ nrmT /*[maxbins][3][n_tables]*/) {
    __shared__ double tile[3][32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int t_in = blockIdx.x * 32 + ty;
    const int64_t i_in = (int64_t)blockIdx.y * 32 + tx;
    double n0 = 0, n1 = 0, n2 = 0;
    if (t_in < n_tables) {
        const int64_t b0 = start[t_in], n = len[t_in];
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
viterbi_forward_kernel(int n_tables, const int64_t *__restrict__ len, int is_log, const double *__restrict__ nrmT, double lp01,
                       double lp02, double lp12, uint8_t *__restrict__ fromT /*[maxbins][n_tables]*/,
                       double *__restrict__ scoreT /*[maxbins][3][n_tables]*/, double *__restrict__ last /*[n_tables][3]*/,
                       uint8_t *__restrict__ flag) {
    const int tb = blockIdx.x * blockDim.x + threadIdx.x;
    if (tb >= n_tables) return;
    const int64_t n = len[tb];
    bool tie = false;
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
                const double a0 = s0 + n0, a1 = (s1 + n0) + lp01, a2 = (s2 + n0) + lp02;
                const double b0 = (s0 + n1) + lp01, b1 = s1 + n1, b2 = (s2 + n1) + lp12;
                const double c0 = (s0 + n2) + lp02, c1 = (s1 + n2) + lp12, c2 = s2 + n2;
                const int k0 = argmax3(a0, a1, a2), k1 = argmax3(b0, b1, b2), k2 = argmax3(c0, c1, c2);
                const double ts = tie_scale(i);
                tie |= near_tie(a0, a1, a2, k0, ts) | near_tie(b0, b1, b2, k1, ts) | near_tie(c0, c1, c2, k2, ts);
                s0 = k0 == 0 ? a0 : (k0 == 1 ? a1 : a2);
                s1 = k1 == 0 ? b0 : (k1 == 1 ? b1 : b2);
                s2 = k2 == 0 ? c0 : (k2 == 1 ? c1 : c2);
                from = (uint8_t)(k0 | (k1 << 2) | (k2 << 4));
            }
            if (!is_log) tie |= below_ldbl(s0, s1, s2);
            fromT[(size_t)i * nT + tb] = from;
            __stcs(scoreT + ((size_t)i * 3 + 0) * nT + tb, s0);
            __stcs(scoreT + ((size_t)i * 3 + 1) * nT + tb, s1);
            __stcs(scoreT + ((size_t)i * 3 + 2) * nT + tb, s2);
        }
    }
    last[(size_t)tb * 3 + 0] = s0;
    last[(size_t)tb * 3 + 1] = s1;
    last[(size_t)tb * 3 + 2] = s2;
    if (n > 0) tie |= near_tie(s0, s1, s2, argmax3(s0, s1, s2), n);
    flag[tb] = tie ? 1 : 0;
}

// (3) back-trace; the state overwrites the back-pointer byte in place
__global__ void __launch_bounds__(64)
viterbi_back_kernel(int n_tables, const int64_t *__restrict__ len, uint8_t *__restrict__ fromT, const double *__restrict__ last,
                    long long *__restrict__ counts) {
    const int tb = blockIdx.x * blockDim.x + threadIdx.x;
    if (tb >= n_tables) return;
    const int64_t n = len[tb];
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
viterbi_out_kernel(int n_tables, const int64_t *__restrict__ start, const int64_t *__restrict__ len, const uint8_t *__restrict__ stateT,
                   const double *__restrict__ scoreT, uint8_t *__restrict__ state, double *__restrict__ score) {
    __shared__ double tile[3][32][33];
    __shared__ uint8_t st[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int t_in = blockIdx.x * 32 + tx;
    const int64_t i_in = (int64_t)blockIdx.y * 32 + ty;
    const size_t nT = (size_t)n_tables;
    if (t_in < n_tables && i_in < len[t_in]) {
#pragma unroll
        for (int s = 0; s < 3; s++) tile[s][ty][tx] = __ldcs(scoreT + ((size_t)i_in * 3 + s) * nT + t_in);
        st[ty][tx] = stateT[(size_t)i_in * nT + t_in];
    }
    __syncthreads();
    const int t_out = blockIdx.x * 32 + ty;
    const int64_t i_out = (int64_t)blockIdx.y * 32 + tx;
    if (t_out < n_tables) {
        const int64_t b0 = start[t_out], n = len[t_out];
        if (i_out < n) {
            double *o = score + (b0 + i_out) * 3;
            o[0] = tile[0][tx][ty];
            o[1] = tile[1][tx][ty];
            o[2] = tile[2][tx][ty];
            state[b0 + i_out] = st[tx][ty];
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Fused forward pass for large batches (C4: 10,000 tables x 10,000 bins): one read of the likelihoods, one write of
// the scores.  A block takes 32 tables.  Four worker warps move 16-bin chunks between global memory and shared
// memory with coalesced accesses (a table's chunk is 384 contiguous bytes) and turn the likelihoods into ln nrm
// there — that is where the fp64 exp / log work is, and it is parallel over bins; ONE solver warp (lane = table)
// walks the chunk sequentially, which the recursion demands, touching shared memory only.  Chunks are double
// buffered, so loads, normalisation, recursion and stores of neighbouring chunks overlap.  The arithmetic is that
// of viterbi_kernel, operation for operation.  Back-pointers leave bin-major (coalesced over tables) for
// viterbi_back_kernel; scores leave in the caller's table-major layout.
constexpr int VF_TABLES = 32, VF_BINS = 16, VF_ROW = VF_BINS * 3 + 1;  // padded row: lanes hit distinct banks
constexpr int VF_WORKERS = 256, VF_THREADS = VF_WORKERS + 32;
constexpr int VF_SMEM = 4 * VF_TABLES * VF_ROW * 8 + 2 * VF_TABLES * 8 + 2 * VF_TABLES * 3 * 8 + VF_TABLES * 4 + 2 * VF_BINS * VF_TABLES;
__global__ void __launch_bounds__(VF_THREADS, 3)
viterbi_fused_kernel(int n_tables, int64_t maxbins, const int64_t *__restrict__ start, const int64_t *__restrict__ len,
                     const double *__restrict__ lik, int is_log, double lp01, double lp02, double lp12, double *__restrict__ score,
                     uint8_t *__restrict__ fromT /*[maxbins][n_tables]*/, double *__restrict__ last /*[n_tables][3]*/,
                     uint8_t *__restrict__ flag) {
    extern __shared__ double vf_smem[];
    double (*nrm)[VF_TABLES][VF_ROW] = reinterpret_cast<double (*)[VF_TABLES][VF_ROW]>(vf_smem);
    double (*sco)[VF_TABLES][VF_ROW] = reinterpret_cast<double (*)[VF_TABLES][VF_ROW]>(vf_smem + 2 * VF_TABLES * VF_ROW);
    int64_t *s_start = reinterpret_cast<int64_t *>(vf_smem + 4 * VF_TABLES * VF_ROW), *s_len = s_start + VF_TABLES;
    double (*prevlast)[VF_TABLES][3] = reinterpret_cast<double (*)[VF_TABLES][3]>(s_len + VF_TABLES);  // last bin's scores of a chunk
    int *tieflag = reinterpret_cast<int *>(&prevlast[2][0][0]);
    uint8_t (*frm)[VF_BINS][VF_TABLES] = reinterpret_cast<uint8_t (*)[VF_BINS][VF_TABLES]>(tieflag + VF_TABLES);
    const int tb0 = blockIdx.x * VF_TABLES;
    if (threadIdx.x < VF_TABLES) {
        const int tb = tb0 + threadIdx.x;
        s_start[threadIdx.x] = tb < n_tables ? start[tb] : 0;
        s_len[threadIdx.x] = tb < n_tables ? len[tb] : 0;
        tieflag[threadIdx.x] = 0;
    }
    __syncthreads();
    int64_t blk_bins = 0;
    for (int k = 0; k < VF_TABLES; k++) blk_bins = max(blk_bins, s_len[k]);
    const int nchunk = (int)((blk_bins + VF_BINS - 1) / VF_BINS);
    // named barriers: 1 + b = chunk in buffer b is normalised (workers arrive, solver syncs);
    //                 3 + b = chunk in buffer b is solved (solver arrives, workers sync); 5 = workers only
    if (threadIdx.x >= 32) {
        // ===== workers =====
        const int w = threadIdx.x - 32;
        constexpr int NLD = VF_TABLES * VF_BINS * 3 / VF_WORKERS;
        // raw likelihoods of the NEXT chunk are fetched (coalesced: 32 tables x 48 doubles) before this chunk's
        // exp / log work starts, so DRAM latency hides under it
        double pre[NLD];
        auto fetch = [&](int c) {
            const int64_t i0 = (int64_t)c * VF_BINS;
#pragma unroll
            for (int k = 0; k < NLD; k++) {
                const int idx = k * VF_WORKERS + w;
                const int t = idx / (VF_BINS * 3), d = idx % (VF_BINS * 3);
                pre[k] = (i0 + d / 3 < s_len[t]) ? __ldcs(lik + (s_start[t] + i0) * 3 + d) : 0.0;
            }
        };
        if (nchunk > 0) fetch(0);
        for (int c = 0; c <= nchunk; c++) {
            const int b = c & 1;
            if (c < nchunk) {
                const int64_t i0 = (int64_t)c * VF_BINS;
#pragma unroll
                for (int k = 0; k < NLD; k++) {
                    const int idx = k * VF_WORKERS + w;
                    nrm[b][idx / (VF_BINS * 3)][idx % (VF_BINS * 3)] = pre[k];
                }
                if (c + 1 < nchunk) fetch(c + 1);
                asm volatile("bar.sync 5, %0;" ::"n"(VF_WORKERS) : "memory");
                // ln nrm in place: 512 bins, 4 per worker
#pragma unroll
                for (int k = 0; k < VF_TABLES * VF_BINS / VF_WORKERS; k++) {
                    const int idx = k * VF_WORKERS + w;
                    const int t = idx / VF_BINS, i = idx % VF_BINS;
                    if (i0 + i < s_len[t]) {
                        const double v[3] = {nrm[b][t][i * 3], nrm[b][t][i * 3 + 1], nrm[b][t][i * 3 + 2]};
                        double n0, n1, n2;
                        ln_nrm(v, is_log, n0, n1, n2);
                        nrm[b][t][i * 3] = n0;
                        nrm[b][t][i * 3 + 1] = n1;
                        nrm[b][t][i * 3 + 2] = n2;
                    }
                }
                __threadfence_block();
                if (b == 0) asm volatile("bar.arrive 1, %0;" ::"n"(VF_THREADS) : "memory");
                else asm volatile("bar.arrive 2, %0;" ::"n"(VF_THREADS) : "memory");
            }
            if (c >= 1) {  // drain the chunk the solver has just finished
                const int pb = (c - 1) & 1;
                if (pb == 0) asm volatile("bar.sync 3, %0;" ::"n"(VF_THREADS) : "memory");
                else asm volatile("bar.sync 4, %0;" ::"n"(VF_THREADS) : "memory");
                const int64_t i0 = (int64_t)(c - 1) * VF_BINS;
                // Near-tie and long-double-range tests, moved off the solver's critical path: with the scores of bin
                // i - 1 and ln nrm of bin i both still in shared memory, every bin's nine candidates can be rebuilt
                // independently (the same expressions, hence the same doubles) — 512 bins over 256 workers.
#pragma unroll
                for (int k = 0; k < VF_TABLES * VF_BINS / VF_WORKERS; k++) {
                    const int idx = k * VF_WORKERS + w;
                    const int t = idx / VF_BINS, i = idx % VF_BINS;
                    const int64_t gi = i0 + i;
                    if (gi < s_len[t]) {
                        const double *cur = &sco[pb][t][i * 3];
                        bool tie = !is_log && below_ldbl(cur[0], cur[1], cur[2]);
                        if (gi > 0) {
                            const double *pv = i > 0 ? &sco[pb][t][(i - 1) * 3] : &prevlast[pb ^ 1][t][0];
                            const double p0 = pv[0], p1 = pv[1], p2 = pv[2];
                            const double n0 = nrm[pb][t][i * 3], n1 = nrm[pb][t][i * 3 + 1], n2 = nrm[pb][t][i * 3 + 2];
                            const double ts = tie_scale(gi);
                            const double a0 = p0 + n0, a1 = (p1 + n0) + lp01, a2 = (p2 + n0) + lp02;
                            const double b0 = (p0 + n1) + lp01, b1 = p1 + n1, b2 = (p2 + n1) + lp12;
                            const double c0 = (p0 + n2) + lp02, c1 = (p1 + n2) + lp12, c2 = p2 + n2;
                            tie |= near_tie(a0, a1, a2, argmax3(a0, a1, a2), ts) | near_tie(b0, b1, b2, argmax3(b0, b1, b2), ts) |
                                   near_tie(c0, c1, c2, argmax3(c0, c1, c2), ts);
                        }
                        if (tie) tieflag[t] = 1;
                        if (i == VF_BINS - 1) { prevlast[pb][t][0] = cur[0]; prevlast[pb][t][1] = cur[1]; prevlast[pb][t][2] = cur[2]; }
                    }
                }
#pragma unroll
                for (int k = 0; k < VF_TABLES * VF_BINS * 3 / VF_WORKERS; k++) {
                    const int idx = k * VF_WORKERS + w;
                    const int t = idx / (VF_BINS * 3), d = idx % (VF_BINS * 3);
                    if (i0 + d / 3 < s_len[t]) __stcs(score + (s_start[t] + i0) * 3 + d, sco[pb][t][d]);
                }
#pragma unroll
                for (int k = 0; k < VF_TABLES * VF_BINS / VF_WORKERS; k++) {
                    const int idx = k * VF_WORKERS + w;
                    const int i = idx / VF_TABLES, t = idx % VF_TABLES;
                    if (tb0 + t < n_tables && i0 + i < s_len[t]) fromT[(size_t)(i0 + i) * n_tables + tb0 + t] = frm[pb][i][t];
                }
                asm volatile("bar.sync 5, %0;" ::"n"(VF_WORKERS) : "memory");  // buffer pb is free for chunk c + 1
            }
        }
    } else {
        // ===== solver: lane = table.  A single warp walks the bins, so its time is (instructions per bin) x (issue
        // latency): the step is branch-free and starts from s = 0, which makes bin 0 an ordinary step (the diagonal
        // candidate wins there, s_k = ln nrm_k exactly; its back-pointers are never followed). =====
        const int lane = threadIdx.x;
        const int64_t n = s_len[lane];
        double s0 = 0, s1 = 0, s2 = 0;
        for (int c = 0; c < nchunk; c++) {
            const int b = c & 1;
            if (b == 0) asm volatile("bar.sync 1, %0;" ::"n"(VF_THREADS) : "memory");
            else asm volatile("bar.sync 2, %0;" ::"n"(VF_THREADS) : "memory");
            const int64_t i0 = (int64_t)c * VF_BINS;
            const double *nr = &nrm[b][lane][0];
            double *so = &sco[b][lane][0];
            const int nq = (int)min((int64_t)VF_BINS, n - i0);  // bins of this table in the chunk (<= 0: none)
#pragma unroll 4
            for (int q = 0; q < VF_BINS; q++) {
                const double n0 = nr[q * 3], n1 = nr[q * 3 + 1], n2 = nr[q * 3 + 2];
                // candidates (prev_k + ln nrm_s) + ln pen(k, s); strict '>' arg-max: lowest index wins, NaN never wins
                const double a0 = s0 + n0, a1 = (s1 + n0) + lp01, a2 = (s2 + n0) + lp02;
                const double b0 = (s0 + n1) + lp01, b1 = s1 + n1, b2 = (s2 + n1) + lp12;
                const double c0 = (s0 + n2) + lp02, c1 = (s1 + n2) + lp12, c2 = s2 + n2;
                const bool ga = a1 > a0, gb = b1 > b0, gc = c1 > c0;
                const double ma = ga ? a1 : a0, mb = gb ? b1 : b0, mc = gc ? c1 : c0;
                const bool ha = a2 > ma, hb = b2 > mb, hc = c2 > mc;
                const bool live = q < nq;
                s0 = live ? (ha ? a2 : ma) : s0;
                s1 = live ? (hb ? b2 : mb) : s1;
                s2 = live ? (hc ? c2 : mc) : s2;
                so[q * 3] = s0;
                so[q * 3 + 1] = s1;
                so[q * 3 + 2] = s2;
                frm[b][q][lane] = (uint8_t)((ha ? 2 : (int)ga) | ((hb ? 2 : (int)gb) << 2) | ((hc ? 2 : (int)gc) << 4));
            }
            __threadfence_block();
            if (b == 0) asm volatile("bar.arrive 3, %0;" ::"n"(VF_THREADS) : "memory");
            else asm volatile("bar.arrive 4, %0;" ::"n"(VF_THREADS) : "memory");
        }
        const int tb = tb0 + lane;
        if (tb < n_tables) {
            last[(size_t)tb * 3 + 0] = s0;
            last[(size_t)tb * 3 + 1] = s1;
            last[(size_t)tb * 3 + 2] = s2;
        }
    }
    __syncthreads();  // the workers' last drain has set the tie flags
    if (threadIdx.x < VF_TABLES) {
        const int tb = tb0 + threadIdx.x;
        if (tb < n_tables) {
            const int64_t n = s_len[threadIdx.x];
            bool tie = tieflag[threadIdx.x] != 0;
            if (n > 0) {
                const double l0 = last[(size_t)tb * 3], l1 = last[(size_t)tb * 3 + 1], l2 = last[(size_t)tb * 3 + 2];
                tie |= near_tie(l0, l1, l2, argmax3(l0, l1, l2), n);
            }
            flag[tb] = tie ? 1 : 0;
        }
    }
}

// states of the back-trace (bin-major) -> the caller's table-major layout
__global__ void __launch_bounds__(1024)
viterbi_state_out_kernel(int n_tables, const int64_t *__restrict__ start, const int64_t *__restrict__ len, const uint8_t *__restrict__ stateT,
                         uint8_t *__restrict__ state) {
    __shared__ uint8_t st[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int t_in = blockIdx.x * 32 + tx;
    const int64_t i_in = (int64_t)blockIdx.y * 32 + ty;
    if (t_in < n_tables && i_in < len[t_in]) st[ty][tx] = stateT[(size_t)i_in * n_tables + t_in];
    __syncthreads();
    const int t_out = blockIdx.x * 32 + ty;
    const int64_t i_out = (int64_t)blockIdx.y * 32 + tx;
    if (t_out < n_tables && i_out < len[t_out]) state[start[t_out] + i_out] = st[tx][ty];
}

}  // namespace ibdgem

using namespace ibdgem;

// ---------------------------------------------------------------------------------------------
// host side
namespace {

// The reference's own recurrence for one table, in x87 long double (src/hiddengem.c:74-76, 108-147, 246-257): the
// arbiter for tables the device flagged (near-tie at an arg-max, or scores outside the normal long double range).
// is_log = 1 (the engine's natural-log window scores) runs the same recurrence on long double logarithms.
void viterbi_long_double(const double *lik, int64_t n, int is_log, double p01, double p02, double p12, uint8_t *state,
                         double *score_log, int64_t counts[3]) {
    counts[0] = counts[1] = counts[2] = 0;
    if (n <= 0) return;
    std::vector<uint8_t> from((size_t)n);
    long double s[3] = {0, 0, 0};
    const long double pen[3][3] = {{1.0L, (long double)p01, (long double)p02},
                                   {(long double)p01, 1.0L, (long double)p12},
                                   {(long double)p02, (long double)p12, 1.0L}};
    long double lpen[3][3];
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) lpen[a][b] = a == b ? 0.0L : logl(pen[a][b]);
    for (int64_t i = 0; i < n; i++) {
        const double *L = lik + i * 3;
        long double nrm[3];
        if (is_log) {
            const long double m = fmaxl(L[0], fmaxl(L[1], L[2]));
            if (m == -INFINITY) {
                nrm[0] = nrm[1] = nrm[2] = NAN;
            } else {
                const long double z = m + logl(expl(L[0] - m) + expl(L[1] - m) + expl(L[2] - m));
                for (int k = 0; k < 3; k++) nrm[k] = (long double)L[k] - z;
            }
        } else {
            volatile double tot = L[0] + L[1];
            tot = tot + L[2];
            for (int k = 0; k < 3; k++) {
                volatile double q = L[k] / tot;  // fp64 quotient, as the reference stores it
                nrm[k] = q;
            }
        }
        if (i == 0) {
            for (int k = 0; k < 3; k++) s[k] = nrm[k];
            from[0] = (uint8_t)(0 | (1 << 2) | (2 << 4));
        } else {
            long double ns[3];
            uint8_t f = 0;
            for (int j = 0; j < 3; j++) {
                long double c[3];
                for (int k = 0; k < 3; k++) {
                    if (is_log)
                        c[k] = k == j ? s[k] + nrm[j] : (s[k] + nrm[j]) + lpen[k][j];
                    else
                        c[k] = k == j ? s[k] * nrm[j] : s[k] * nrm[j] * pen[k][j];
                }
                int best = 0;
                for (int k = 0; k < 3; k++)
                    if (c[k] > c[best]) best = k;  // find_max_idx: strict >, lowest index wins, NaN never wins
                ns[j] = c[best];
                f |= (uint8_t)(best << (2 * j));
            }
            for (int k = 0; k < 3; k++) s[k] = ns[k];
            from[(size_t)i] = f;
        }
        for (int k = 0; k < 3; k++) score_log[i * 3 + k] = is_log ? (double)s[k] : (double)logl(s[k]);
    }
    int cur = 0;
    for (int k = 0; k < 3; k++)
        if (s[k] > s[cur]) cur = k;
    for (int64_t i = n - 1; i >= 0; i--) {
        state[i] = (uint8_t)cur;
        counts[cur]++;
        cur = (from[(size_t)i] >> (2 * cur)) & 3;
    }
}

// One batch of tables, everything in device memory.  start / len describe where each table's bins sit in lik /
// state / score; counts and flag are per table.  Kernels only (no host synchronisation).
int viterbi_on_device(ibdgem_engine *e, int n_tables, int64_t maxbins, int64_t total_bins, const int64_t *d_start, const int64_t *d_len,
                      const double *d_lik, int is_log, double p01, double p02, double p12, uint8_t *d_state, double *d_score,
                      long long *d_counts, uint8_t *d_flag) {
    const bool batched = n_tables >= 64 && (double)maxbins * n_tables <= 2.0 * (double)total_bins;
    if (batched) {
        double *d_last;
        uint8_t *d_fromT;
        const size_t cells = (size_t)maxbins * n_tables;
        static const int fused_env = [] { const char *s = getenv("IBDGEM_VITERBI_FUSED"); return s ? atoi(s) : 1; }();
        // the fused kernel starts from s = 0 and treats bin 0 as an ordinary step, which is the reference's bin 0 only
        // while no switch is rewarded (penalties <= 1)
        const bool fused = fused_env && p01 <= 1.0 && p02 <= 1.0 && p12 <= 1.0;
        if (scratch(e, SC_HG_FROMT, cells, (void **)&d_fromT) || scratch(e, SC_HG_LAST, (size_t)n_tables * 24, (void **)&d_last)) return 1;
        const dim3 tiles((unsigned)((n_tables + 31) / 32), (unsigned)((maxbins + 31) / 32));
        if (fused) {
            {
                LaunchScope ls(e, K_VITERBI);
                IBD_CUDA(cudaFuncSetAttribute(viterbi_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VF_SMEM));
                viterbi_fused_kernel<<<(n_tables + VF_TABLES - 1) / VF_TABLES, VF_THREADS, VF_SMEM, e->stream>>>(
                    n_tables, maxbins, d_start, d_len, d_lik, is_log, log(p01), log(p02), log(p12), d_score, d_fromT, d_last, d_flag);
            }
            {
                LaunchScope ls(e, K_VITERBI_BACK);
                viterbi_back_kernel<<<(n_tables + 63) / 64, 64, 0, e->stream>>>(n_tables, d_len, d_fromT, d_last, d_counts);
            }
            {
                LaunchScope ls(e, K_VITERBI_OUT);
                viterbi_state_out_kernel<<<tiles, 1024, 0, e->stream>>>(n_tables, d_start, d_len, d_fromT, d_state);
            }
        } else {  // the four-kernel pipeline with bin-major intermediates (kept for A/B measurement)
            double *d_nrmT, *d_scoreT;
            if (scratch(e, SC_HG_NRMT, cells * 24, (void **)&d_nrmT) || scratch(e, SC_HG_SCORET, cells * 24, (void **)&d_scoreT)) return 1;
            {
                LaunchScope ls(e, K_VITERBI_NORM);
                viterbi_norm_kernel<<<tiles, 1024, 0, e->stream>>>(n_tables, maxbins, d_start, d_len, d_lik, is_log, d_nrmT);
            }
            {
                LaunchScope ls(e, K_VITERBI);
                viterbi_forward_kernel<<<(n_tables + 63) / 64, 64, 0, e->stream>>>(n_tables, d_len, is_log, d_nrmT, log(p01), log(p02), log(p12),
                                                                                 d_fromT, d_scoreT, d_last, d_flag);
            }
            {
                LaunchScope ls(e, K_VITERBI_BACK);
                viterbi_back_kernel<<<(n_tables + 63) / 64, 64, 0, e->stream>>>(n_tables, d_len, d_fromT, d_last, d_counts);
            }
            {
                LaunchScope ls(e, K_VITERBI_OUT);
                viterbi_out_kernel<<<tiles, 1024, 0, e->stream>>>(n_tables, d_start, d_len, d_fromT, d_scoreT, d_state, d_score);
            }
        }
    } else {
        LaunchScope ls(e, K_VITERBI);
        viterbi_kernel<<<(n_tables + 127) / 128, 128, 0, e->stream>>>(n_tables, d_start, d_len, d_lik, is_log, log(p01), log(p02), log(p12),
                                                                     d_state, d_score, d_counts, d_flag);
    }
    IBD_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

// Host buffers in, host buffers out.  The tables go through in batches: while batch b is scored, batch b + 1 is on its
// way up and the results of batch b - 1 on their way down (three streams; page-locked buffers make the copies
// asynchronous, pageable ones are simply staged by the driver).
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
    int64_t *d_start;
    uint8_t *d_state, *d_flag;
    long long *d_counts;
    if (scratch(e, SC_HG_LIK, (size_t)nb * 24, (void **)&d_lik) || scratch(e, SC_HG_OFF, (size_t)n_tables * 16, (void **)&d_start) ||
        scratch(e, SC_HG_STATE, (size_t)nb + (size_t)n_tables, (void **)&d_state) || scratch(e, SC_HG_SCORE, (size_t)nb * 24, (void **)&d_score) ||
        scratch(e, SC_HG_COUNTS, (size_t)n_tables * 24, (void **)&d_counts))
        return 1;
    d_flag = d_state + nb;
    int64_t *d_len = d_start + n_tables;
    std::vector<int64_t> h_sl((size_t)n_tables * 2);
    for (int t = 0; t < n_tables; t++) {
        h_sl[(size_t)t] = bin_offsets[t];
        h_sl[(size_t)n_tables + t] = bin_offsets[t + 1] - bin_offsets[t];
    }
    if (!e->copy_stream) {
        IBD_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        IBD_CUDA(cudaEventCreateWithFlags(&e->ev_order, cudaEventDisableTiming));
    }
    if (!e->d2h_stream) IBD_CUDA(cudaStreamCreateWithFlags(&e->d2h_stream, cudaStreamNonBlocking));
    IBD_CUDA(cudaMemcpyAsync(d_start, h_sl.data(), h_sl.size() * 8, cudaMemcpyHostToDevice, e->stream));
    // earlier work on the engine stream may still use the scratch buffers: the side streams start after it
    IBD_CUDA(cudaEventRecord(e->ev_order, e->stream));
    IBD_CUDA(cudaStreamWaitEvent(e->copy_stream, e->ev_order, 0));
    IBD_CUDA(cudaStreamWaitEvent(e->d2h_stream, e->ev_order, 0));
    // Batches of whole tables, at least 64 so the batched kernels apply.  A batch's kernels take the time of one block
    // (the recursion is sequential in bins) however few tables it holds, so batches are few: about six per call —
    // copy-bound with ~1/6 of the upload exposed at the front and ~1/6 of the download at the back — and never below
    // ~192 MB of likelihoods.
    const int64_t target_bins = std::max<int64_t>((int64_t)8 << 20, nb / 6);
    std::vector<int> cuts{0};
    for (int t = 0; t < n_tables;) {
        int t1 = t;
        while (t1 < n_tables && (t1 - t < 64 || bin_offsets[t1] - bin_offsets[t] < target_bins)) t1++;
        if (n_tables - t1 < 64) t1 = n_tables;  // no runt batch at the end
        cuts.push_back(t1);
        t = t1;
    }
    const size_t nbatch = cuts.size() - 1;
    std::vector<cudaEvent_t> ev_up(nbatch), ev_done(nbatch);
    for (size_t b = 0; b < nbatch; b++) {
        IBD_CUDA(cudaEventCreateWithFlags(&ev_up[b], cudaEventDisableTiming));
        IBD_CUDA(cudaEventCreateWithFlags(&ev_done[b], cudaEventDisableTiming));
    }
    int rc = 0;
    for (size_t b = 0; b < nbatch && !rc; b++) {
        const int t0 = cuts[b], t1 = cuts[b + 1];
        const int64_t b0 = bin_offsets[t0], b1 = bin_offsets[t1];
        IBD_CUDA(cudaMemcpyAsync(d_lik + b0 * 3, lik + b0 * 3, (size_t)(b1 - b0) * 24, cudaMemcpyHostToDevice, e->copy_stream));
        IBD_CUDA(cudaEventRecord(ev_up[b], e->copy_stream));
    }
    for (size_t b = 0; b < nbatch && !rc; b++) {
        const int t0 = cuts[b], t1 = cuts[b + 1];
        const int64_t b0 = bin_offsets[t0], b1 = bin_offsets[t1];
        int64_t maxbins = 0;
        for (int t = t0; t < t1; t++) maxbins = std::max<int64_t>(maxbins, bin_offsets[t + 1] - bin_offsets[t]);
        IBD_CUDA(cudaStreamWaitEvent(e->stream, ev_up[b], 0));
        rc = viterbi_on_device(e, t1 - t0, maxbins, b1 - b0, d_start + t0, d_len + t0, d_lik, is_log, p01, p02, p12, d_state, d_score,
                               d_counts + (size_t)t0 * 3, d_flag + t0);
        if (rc) break;
        IBD_CUDA(cudaEventRecord(ev_done[b], e->stream));
        IBD_CUDA(cudaStreamWaitEvent(e->d2h_stream, ev_done[b], 0));
        if (state) IBD_CUDA(cudaMemcpyAsync(state + b0, d_state + b0, (size_t)(b1 - b0), cudaMemcpyDeviceToHost, e->d2h_stream));
        if (score_log) IBD_CUDA(cudaMemcpyAsync(score_log + b0 * 3, d_score + b0 * 3, (size_t)(b1 - b0) * 24, cudaMemcpyDeviceToHost, e->d2h_stream));
    }
    std::vector<uint8_t> h_flag((size_t)n_tables, 0);
    std::vector<long long> h_counts((size_t)n_tables * 3);
    if (!rc) {
        IBD_CUDA(cudaMemcpyAsync(h_counts.data(), d_counts, (size_t)n_tables * 24, cudaMemcpyDeviceToHost, e->stream));
        IBD_CUDA(cudaMemcpyAsync(h_flag.data(), d_flag, (size_t)n_tables, cudaMemcpyDeviceToHost, e->stream));
    }
    cudaStreamSynchronize(e->copy_stream);
    cudaStreamSynchronize(e->stream);
    cudaStreamSynchronize(e->d2h_stream);
    for (size_t b = 0; b < nbatch; b++) {
        cudaEventDestroy(ev_up[b]);
        cudaEventDestroy(ev_done[b]);
    }
    if (rc) return 1;
    IBD_CUDA(cudaGetLastError());
    settle_timers(e);
    // flagged tables: the reference's long double recurrence decides (tables are independent: a few host threads)
    std::vector<int> flagged;
    for (int t = 0; t < n_tables; t++) {
        if (state_counts) {
            state_counts[(size_t)t * 3] = h_counts[(size_t)t * 3];
            state_counts[(size_t)t * 3 + 1] = h_counts[(size_t)t * 3 + 1];
            state_counts[(size_t)t * 3 + 2] = h_counts[(size_t)t * 3 + 2];
        }
        if (h_flag[(size_t)t]) flagged.push_back(t);
    }
    e->hg_flagged = (int64_t)flagged.size();
    auto redo = [&](size_t k0, size_t step) {
        std::vector<uint8_t> st_tmp;
        std::vector<double> sc_tmp;
        for (size_t k = k0; k < flagged.size(); k += step) {
            const int t = flagged[k];
            const int64_t b0 = bin_offsets[t], n = bin_offsets[t + 1] - b0;
            int64_t c[3];
            st_tmp.resize((size_t)n);
            sc_tmp.resize((size_t)n * 3);
            viterbi_long_double(lik + b0 * 3, n, is_log, p01, p02, p12, st_tmp.data(), sc_tmp.data(), c);
            if (state) memcpy(state + b0, st_tmp.data(), (size_t)n);
            if (score_log) memcpy(score_log + b0 * 3, sc_tmp.data(), (size_t)n * 24);
            if (state_counts) {
                state_counts[(size_t)t * 3] = c[0];
                state_counts[(size_t)t * 3 + 1] = c[1];
                state_counts[(size_t)t * 3 + 2] = c[2];
            }
        }
    };
    const size_t nthr = std::max<size_t>(1, std::min<size_t>({flagged.size(), (size_t)16, (size_t)std::thread::hardware_concurrency()}));
    if (nthr <= 1) {
        redo(0, 1);
    } else {
        std::vector<std::thread> pool;
        for (size_t k = 0; k < nthr; k++) pool.emplace_back(redo, k, nthr);
        for (auto &th : pool) th.join();
    }
    return 0;
}

// Device buffers in, device buffers out: the hiddengem front-end for scores that never left the GPU.
//   table_stride = 0: tables are packed, table t holds bins [bin_offsets[t], bin_offsets[t+1]) of d_lik / d_state / d_score;
//   table_stride > 0: table t starts at bin t * table_stride (the layout of ibdgem_scores.w_loglik_device with
//                     table_stride = max_windows) and has bin_offsets[t+1] - bin_offsets[t] bins.
// bin_offsets is a HOST array.  Flagged tables (near-ties) are re-evaluated on the host in long double and patched
// back into the device buffers.
extern "C" int hiddengem_viterbi_batch_device(ibdgem_engine *e, int32_t n_tables, const int64_t *bin_offsets, int64_t table_stride,
                                              const double *d_lik, int32_t is_log, double p01, double p02, double p12,
                                              uint8_t *d_state, double *d_score_log, int64_t *d_state_counts) {
    if (!e || n_tables <= 0 || !bin_offsets || !d_lik || !d_state || !d_score_log || !d_state_counts || table_stride < 0) {
        set_error("[::] ERROR in hiddengem_viterbi_batch_device(): bad arguments.");
        return 1;
    }
    IBD_CUDA(cudaSetDevice(e->device));
    std::vector<int64_t> h_sl((size_t)n_tables * 2);
    int64_t maxbins = 0, total = 0;
    for (int t = 0; t < n_tables; t++) {
        const int64_t n = bin_offsets[t + 1] - bin_offsets[t];
        if (n < 0 || (table_stride > 0 && n > table_stride)) {
            set_error("[::] ERROR in hiddengem_viterbi_batch_device(): table %d has %lld bins (stride %lld).", t, (long long)n, (long long)table_stride);
            return 1;
        }
        h_sl[(size_t)t] = table_stride > 0 ? (int64_t)t * table_stride : bin_offsets[t];
        h_sl[(size_t)n_tables + t] = n;
        maxbins = std::max(maxbins, n);
        total += n;
    }
    if (total <= 0) {
        set_error("[::] ERROR parsing likelihood data; make sure input is valid.");
        return 1;
    }
    int64_t *d_start;
    uint8_t *d_flag;
    if (scratch(e, SC_HG_OFF, (size_t)n_tables * 16, (void **)&d_start) || scratch(e, SC_HG_STATE, (size_t)n_tables, (void **)&d_flag)) return 1;
    IBD_CUDA(cudaMemcpyAsync(d_start, h_sl.data(), h_sl.size() * 8, cudaMemcpyHostToDevice, e->stream));
    static_assert(sizeof(long long) == sizeof(int64_t), "state counts are 64-bit");
    if (viterbi_on_device(e, n_tables, maxbins, total, d_start, d_start + n_tables, d_lik, is_log, p01, p02, p12, d_state, d_score_log,
                          reinterpret_cast<long long *>(d_state_counts), d_flag))
        return 1;
    std::vector<uint8_t> h_flag((size_t)n_tables);
    IBD_CUDA(cudaMemcpyAsync(h_flag.data(), d_flag, (size_t)n_tables, cudaMemcpyDeviceToHost, e->stream));
    IBD_CUDA(cudaStreamSynchronize(e->stream));
    settle_timers(e);
    e->hg_flagged = 0;
    for (int t = 0; t < n_tables; t++) {
        if (!h_flag[(size_t)t]) continue;
        const int64_t b0 = h_sl[(size_t)t], n = h_sl[(size_t)n_tables + t];
        std::vector<double> l((size_t)n * 3), sc((size_t)n * 3);
        std::vector<uint8_t> st((size_t)n);
        int64_t c[3];
        IBD_CUDA(cudaMemcpy(l.data(), d_lik + b0 * 3, (size_t)n * 24, cudaMemcpyDeviceToHost));
        viterbi_long_double(l.data(), n, is_log, p01, p02, p12, st.data(), sc.data(), c);
        IBD_CUDA(cudaMemcpy(d_state + b0, st.data(), (size_t)n, cudaMemcpyHostToDevice));
        IBD_CUDA(cudaMemcpy(d_score_log + b0 * 3, sc.data(), (size_t)n * 24, cudaMemcpyHostToDevice));
        IBD_CUDA(cudaMemcpy(d_state_counts + (size_t)t * 3, c, 24, cudaMemcpyHostToDevice));
        e->hg_flagged++;
    }
    return 0;
}

// Tables of the last hiddengem call that were re-evaluated on the host in long double (near-tie guard).
extern "C" int64_t hiddengem_last_flagged(ibdgem_engine *e) { return e ? e->hg_flagged : -1; }
