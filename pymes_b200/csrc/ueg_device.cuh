// Device-side evaluation of one UEG two-electron integral, shared by the dense block writer
// (ueg_build.cu) and the generated-operand producers of the contraction kernel
// (cc_contract.cu).  Every floating-point operation is spelled with an explicit rounding
// intrinsic so that both users produce bit-identical values.
#pragma once
#include "common.cuh"

namespace pmb {

// The four table words of one element; issued together, consumed by ueg_combine.  The
// contraction kernel's producers load them one tile ahead of their use.
struct UegWords {
    double w0, w1, s0, s1;
};
// PIN: the loads are volatile asm, which the compiler neither deletes nor sinks to the point
// of use -- a prefetch into registers stays where it was written (plain __ldg loads of
// read-only memory may legally be moved next to their consumer, which puts the whole memory
// round trip back on the critical path).
template <bool PIN>
__device__ __forceinline__ double ueg_ld(const double *ptr) {
    if (PIN) {
        double v;
        asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(ptr));
        return v;
    }
    return __ldg(ptr);
}
template <bool PIN = false>
__device__ __forceinline__ UegWords ueg_load_words(int nP, const double *__restrict__ W0a,
                                                   const double *__restrict__ W1a, const double *__restrict__ W0s,
                                                   int p, int q, int r, int s) {
    const int pr = p * nP + r;
    UegWords w = {0.0, 0.0, 0.0, 0.0};
    if (W0a) w.w0 = ueg_ld<PIN>(W0a + pr);
    if (W1a) w.w1 = ueg_ld<PIN>(W1a + pr);
    if (W0s) {
        w.s0 = ueg_ld<PIN>(W0s + pr);
        w.s1 = ueg_ld<PIN>(W0s + q * nP + s);
    }
    return w;
}

// V[p,q,r,s] from its table words; kp = [nP][3] plane-wave vectors (global or a shared copy).
// Every value is consumed unconditionally (the conditions are selects, not branches) so the
// compiler cannot sink a load behind another load's result.
__device__ __forceinline__ double ueg_combine(const UegWords &t, bool has_w1, bool has_sym, const double *kp,
                                              int p, int r, int s) {
    double dot = 0.0;
    if (has_w1) {
        const double rx = kp[3 * r], ry = kp[3 * r + 1], rz = kp[3 * r + 2];
        const double dx = __dsub_rn(rx, kp[3 * p]), dy = __dsub_rn(ry, kp[3 * p + 1]), dz = __dsub_rn(rz, kp[3 * p + 2]);
        const double ex = __dsub_rn(rx, kp[3 * s]), ey = __dsub_rn(ry, kp[3 * s + 1]), ez = __dsub_rn(rz, kp[3 * s + 2]);
        dot = __fma_rn(ez, dz, __fma_rn(ey, dy, __dmul_rn(ex, dx)));
    }
    const double f = __fma_rn(t.w1, dot, t.w0);
    double w = (t.w1 != 0.0) ? f : t.w0;         // reference: the term exists only where w1 != 0
    const double sym = __dadd_rn(w, __dmul_rn(0.5, __dadd_rn(t.s0, t.s1)));
    if (has_sym) w = sym;
    return w;
}

// V[p,q,r,s] for s = s*(p,q,r), reference pymes/model/ueg.py:411-513:
//   W0a[p,r] + W1a[p,r] * (k_r - k_s).(k_r - k_p) + 1/2 (W0s[p,r] + W0s[q,s])
__device__ __forceinline__ double ueg_value(const pmb_ueg_t &u, const double *__restrict__ W0a,
                                            const double *__restrict__ W1a, const double *__restrict__ W0s,
                                            int p, int q, int r, int s) {
    const UegWords t = ueg_load_words(u.n_orb, W0a, W1a, W0s, p, q, r, s);
    return ueg_combine(t, W1a != nullptr, W0s != nullptr, u.kp, p, r, s);
}

}  // namespace pmb
