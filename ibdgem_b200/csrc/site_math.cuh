// site_math.cuh — per-site closed forms of the IBD0/IBD1/IBD2 models on the device.
//
// These follow src/ibd-math.c:84-142 of the reference operation by operation, with explicit
// round-to-nearest intrinsics so that nvcc cannot contract a multiply and an add into an FMA:
// the linear per-site values then round exactly like the reference's -O0 SSE2 doubles (the
// only remaining difference is pow(x, 2.0), evaluated here as x*x).
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace ibdgem {

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// find_pDgf, src/ibd-math.c:84-101
__device__ __forceinline__ double lik_ibd0(double f, double P0, double P1, double P2) {
    if (P0 == 1.0 || P1 == 1.0 || P2 == 1.0) return 1.0;
    const double q = __dsub_rn(1.0, f);
    const double t0 = mul_rn(mul_rn(q, q), P0);
    const double t1 = mul_rn(mul_rn(mul_rn(2.0, q), f), P1);
    const double t2 = mul_rn(mul_rn(f, f), P2);
    double p = add_rn(add_rn(t0, t1), t2);
    if (p == 0.0) p = DBL_MIN;
    return p;
}

// find_pDgIBD1, src/ibd-math.c:104-142, g = A0 + A1 (the reference treats 0|1 and 1|0 alike)
__device__ __forceinline__ double lik_ibd1(int g, double f, double P0, double P1, double P2) {
    const double q = __dsub_rn(1.0, f);
    double p;
    if (g == 0) {
        p = add_rn(mul_rn(f, P1), mul_rn(q, P0));
    } else if (g == 1) {
        const double a = mul_rn(0.5, P1);
        const double b = mul_rn(mul_rn(0.5, q), P0);
        const double c = mul_rn(mul_rn(0.5, f), P2);
        p = add_rn(add_rn(a, b), c);
    } else {
        p = add_rn(mul_rn(q, P1), mul_rn(f, P2));
    }
    if (p == 0.0) p = DBL_MIN;
    return p;
}

// Genotype of individual `indiv` at a site row of the site-major bit panel: haplotypes 2i, 2i+1
// always share a 32-bit word.  Returns a0 | a1 << 1.
__device__ __forceinline__ uint32_t hap_pair(const uint32_t *__restrict__ row, int indiv) {
    const uint32_t w = __ldg(row + (indiv >> 4));
    return (w >> ((indiv & 15) * 2)) & 3u;
}

// Per-target view of one site: everything the reference's filter chain and -D thinning decide
// (src/ibdgem.c:584-630).  Returns IBDGEM_SITE_* and the counts/genotype actually used.
struct SiteView {
    const uint8_t *keep;       // [S] target-independent filters (incl. max-cov on original counts)
    const uint8_t *nref;       // [S]
    const uint8_t *nalt;       // [S]
    const uint32_t *bits;      // [S][Wh]; may be NULL for a target-independent pass (vflag = 0)
    int64_t Wh;
    int64_t S;
    const uint8_t *tgt_counts; // NULL or [T][S][2]
    int vflag;
};

__device__ __forceinline__ int site_eval(const SiteView &v, int tslot, int indiv, int64_t s, int &r,
                                         int &a, int &g) {
    if (!v.keep[s]) return 0;
    const uint32_t pr = v.bits ? hap_pair(v.bits + s * v.Wh, indiv) : 0u;  // bits == NULL: genotype not needed
    g = (int)(pr & 1u) + (int)(pr >> 1);
    if (v.vflag && g == 0) return 0;  // src/ibdgem.c:584-587
    if (v.tgt_counts) {
        const uint8_t *c = v.tgt_counts + ((int64_t)tslot * v.S + s) * 2;
        r = c[0];
        a = c[1];
    } else {
        r = v.nref[s];
        a = v.nalt[s];
    }
    return (r + a >= 1) ? 1 : 2;  // src/ibdgem.c:657
}

}  // namespace ibdgem
