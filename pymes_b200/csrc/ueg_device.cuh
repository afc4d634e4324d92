// Device-side evaluation of one UEG two-electron integral, shared by the dense block writer
// (ueg_build.cu) and the generated-operand producers of the contraction kernel
// (cc_contract.cu).  Every floating-point operation is spelled with an explicit rounding
// intrinsic so that both users produce bit-identical values.
#pragma once
#include "common.cuh"

namespace pmb {

// V[p,q,r,s] for s = s*(p,q,r), reference pymes/model/ueg.py:411-513:
//   W0a[p,r] + W1a[p,r] * (k_r - k_s).(k_r - k_p) + 1/2 (W0s[p,r] + W0s[q,s])
__device__ __forceinline__ double ueg_value(const pmb_ueg_t &u, const double *__restrict__ W0a,
                                            const double *__restrict__ W1a, const double *__restrict__ W0s,
                                            int p, int q, int r, int s) {
    const int nP = u.n_orb;
    const int pr = p * nP + r;
    double w = 0.0;
    if (W0a) w = W0a[pr];
    if (W1a) {
        const double w1 = W1a[pr];
        if (w1 != 0.0) {
            const double dx = __dsub_rn(u.kp[3 * r], u.kp[3 * p]), dy = __dsub_rn(u.kp[3 * r + 1], u.kp[3 * p + 1]),
                         dz = __dsub_rn(u.kp[3 * r + 2], u.kp[3 * p + 2]);
            const double ex = __dsub_rn(u.kp[3 * r], u.kp[3 * s]), ey = __dsub_rn(u.kp[3 * r + 1], u.kp[3 * s + 1]),
                         ez = __dsub_rn(u.kp[3 * r + 2], u.kp[3 * s + 2]);
            const double dot = __fma_rn(ez, dz, __fma_rn(ey, dy, __dmul_rn(ex, dx)));
            w = __fma_rn(w1, dot, w);
        }
    }
    if (W0s) w = __dadd_rn(w, __dmul_rn(0.5, __dadd_rn(W0s[pr], W0s[q * nP + s])));
    return w;
}

}  // namespace pmb
