// UEG momentum-conserving two-electron integrals (reference pymes/model/ueg.py).
//
// HBM-bound by construction: the dense block write is 8 B per stored element
// with ~10 flops per row, so the kernel is laid out for store bandwidth -- one
// warp owns one (p,q,r) row, resolves the single momentum-conserving s* once,
// and streams the row with fully coalesced 8 B stores (32 lanes x 8 B = two
// 128 B lines per instruction).  The q-dependent heavy sums (u_mat: a
// 226 981-term lattice sum per distinct transfer vector) run as one CTA per q.
#include "common.cuh"
#include "ueg_device.cuh"

namespace pmb {

// Correlator u(k^2).  Every argument the build needs is |integer vector|^2 (2 pi/L)^2
// (twists cancel in differences), so the host tabulates the user's correlator
// once over n2 = |k|^2 and the kernels look it up -- any of the reference's
// correlators (ueg.py:740-956) works, not only `trunc`.
__device__ __forceinline__ double u_of(const double *__restrict__ tab, int len, int x, int y, int z) {
    const int n2 = x * x + y * y + z * z;
    return n2 < len ? tab[n2] : 0.0;
}

__global__ void __launch_bounds__(256)
    umat_kernel(double box_len, int cutoff, double omega, const double *__restrict__ tab, int tab_len,
                const int *__restrict__ qvec, double *__restrict__ out) {
    const int q = blockIdx.x;
    const int qi = qvec[3 * q], qj = qvec[3 * q + 1], qk = qvec[3 * q + 2];
    const double TWO_PI = 2.0 * 3.14159265358979323846;
    const double qx = (TWO_PI * qi) / box_len, qy = (TWO_PI * qj) / box_len, qz = (TWO_PI * qk) / box_len;
    const int side = 2 * cutoff + 1;
    const int total = side * side * side;
    double s[1] = {0.0};
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const int iz = t % side - cutoff;
        const int iy = (t / side) % side - cutoff;
        const int ix = t / (side * side) - cutoff;
        // k1 = 2*pi*kPrime/L evaluated as (2*pi*k)/L like the reference (ueg.py:587)
        const double ax = (TWO_PI * ix) / box_len, ay = (TWO_PI * iy) / box_len, az = (TWO_PI * iz) / box_len;
        const double bx = qx - ax, by = qy - ay, bz = qz - az;
        const double ab = ax * bx + ay * by + az * bz;
        s[0] += ab * u_of(tab, tab_len, ix, iy, iz) * u_of(tab, tab_len, qi - ix, qj - iy, qk - iz);
    }
    const double tot = block_reduce_sum(s[0], 0);
    if (threadIdx.x == 0) out[q] = tot / omega;
}

// sums over the occupied orbitals of the exchange-type single contractions
// (ueg.py:518-573); n_occ = N/2 terms, evaluated per thread.  `o` is the orbital
// whose k vector plays the role of p_vec, (dx,dy,dz)/(di,dj,dk) the transfer.
__device__ double ex3_sum(const pmb_ueg_t &u, int o, double dx, double dy, double dz, double ud) {
    double acc = 0.0;
    for (int i = 0; i < u.n_occ; ++i) {
        const double vx = u.kp[3 * o] - u.kp[3 * i], vy = u.kp[3 * o + 1] - u.kp[3 * i + 1],
                     vz = u.kp[3 * o + 2] - u.kp[3 * i + 2];
        const double uv = u_of(u.u_table, u.u_table_len, u.kvec[3 * o] - u.kvec[3 * i],
                               u.kvec[3 * o + 1] - u.kvec[3 * i + 1], u.kvec[3 * o + 2] - u.kvec[3 * i + 2]);
        acc += (vx * dx + vy * dy + vz * dz) * ud * uv;
    }
    return acc / u.omega;
}

__device__ double pk_sum(const pmb_ueg_t &u, int o, double dx, double dy, double dz, int di, int dj, int dk) {
    double acc = 0.0;
    for (int i = 0; i < u.n_occ; ++i) {
        const double bx = u.kp[3 * o] - u.kp[3 * i], by = u.kp[3 * o + 1] - u.kp[3 * i + 1],
                     bz = u.kp[3 * o + 2] - u.kp[3 * i + 2];
        const double ax = bx - dx, ay = by - dy, az = bz - dz;
        const int bi = u.kvec[3 * o] - u.kvec[3 * i], bj = u.kvec[3 * o + 1] - u.kvec[3 * i + 1],
                  bk = u.kvec[3 * o + 2] - u.kvec[3 * i + 2];
        acc += (ax * bx + ay * by + az * bz) * u_of(u.u_table, u.u_table_len, bi - di, bj - dj, bk - dk) *
               u_of(u.u_table, u.u_table_len, bi, bj, bk);
    }
    return acc / u.omega;
}

__global__ void __launch_bounds__(128)
    pair_tables_kernel(pmb_ueg_t u, int mode, const double *__restrict__ umat_pr, double *__restrict__ W0,
                       double *__restrict__ W1) {
    const int nP = u.n_orb;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nP * nP) return;
    const int p = idx / nP, r = idx % nP;
    const double FOUR_PI = 4.0 * 3.14159265358979323846;
    const double dx = u.kp[3 * r] - u.kp[3 * p], dy = u.kp[3 * r + 1] - u.kp[3 * p + 1],
                 dz = u.kp[3 * r + 2] - u.kp[3 * p + 2];
    const double d2 = dx * dx + dy * dy + dz * dz;
    const bool nz = fabs(d2) > 0.0;
    const int di = u.kvec[3 * r] - u.kvec[3 * p], dj = u.kvec[3 * r + 1] - u.kvec[3 * p + 1],
              dk = u.kvec[3 * r + 2] - u.kvec[3 * p + 2];
    const double ud = u.u_table ? u_of(u.u_table, u.u_table_len, di, dj, dk) : 0.0;
    const double um = umat_pr ? umat_pr[idx] : 0.0;
    double w0 = 0.0, w1 = 0.0;
    switch (mode) {
        case PMB_UEG_COULOMB:  // ueg.py:411-413
            if (nz) w0 = FOUR_PI / d2 / u.omega;
            break;
        case PMB_UEG_RPA:  // ueg.py:416-423
            if (nz) w0 = (-(double)u.n_ele * d2 * (ud * ud) / u.omega) / u.omega;
            break;
        case PMB_UEG_ONLY_2B:  // ueg.py:426-437
            if (nz) {
                w0 = (FOUR_PI / d2 + um + d2 * ud) / u.omega;
                w1 = -ud / u.omega;
            } else {
                w0 = um / u.omega;
            }
            break;
        case PMB_UEG_ONLY_HERMI_2B:  // ueg.py:440-447
            w0 = nz ? (FOUR_PI / d2 + um + d2 * ud) / u.omega : um / u.omega;
            break;
        case PMB_UEG_ONLY_NON_HERMI_2B:  // ueg.py:450-457
            if (nz) {
                w0 = (FOUR_PI / d2) / u.omega;
                w1 = -ud / u.omega;
            }
            break;
        case PMB_UEG_EFFECT_2B:  // ueg.py:461-474
            if (nz)
                w0 = -(double)u.n_ele * d2 * (ud * ud) / u.omega + 2.0 * ex3_sum(u, r, dx, dy, dz, ud) -
                     2.0 * ex3_sum(u, p, dx, dy, dz, ud) + 2.0 * pk_sum(u, r, dx, dy, dz, di, dj, dk);
            else
                w0 = 2.0 * pk_sum(u, r, dx, dy, dz, di, dj, dk);
            w0 /= u.omega;
            break;
        case PMB_UEG_EXCHANGE_1:  // ueg.py:478-484
            if (nz) w0 = 2.0 * ex3_sum(u, r, dx, dy, dz, ud) / u.omega;
            break;
        case PMB_UEG_EXCHANGE_2:  // ueg.py:487-493
            if (nz) w0 = -2.0 * ex3_sum(u, p, dx, dy, dz, ud) / u.omega;
            break;
        case PMB_UEG_EXCHANGE_3:  // ueg.py:496-504
            w0 = 2.0 * pk_sum(u, r, dx, dy, dz, di, dj, dk) / u.omega;
            break;
        default: break;
    }
    W0[idx] = w0;
    if (W1) W1[idx] = w1;
}

struct BlockGeom {
    int lo[4], ext[4];
};

// s* = map[k_q - (k_r - k_p)] with the reference's flattened-index bounds
// check (ueg.py:397-407): only the flattened location is range checked, so a
// component outside [-imax, imax] aliases exactly as it does there.
__device__ __forceinline__ int conserving_s(const pmb_ueg_t &u, int p, int q, int r) {
    const int n = 2 * u.imax + 1;
    const int x = u.kvec[3 * q] - (u.kvec[3 * r] - u.kvec[3 * p]);
    const int y = u.kvec[3 * q + 1] - (u.kvec[3 * r + 1] - u.kvec[3 * p + 1]);
    const int z = u.kvec[3 * q + 2] - (u.kvec[3 * r + 2] - u.kvec[3 * p + 2]);
    const long long loc = (long long)n * n * (x + u.imax) + (long long)n * (y + u.imax) + (z + u.imax);
    if (loc < 0 || loc >= (long long)n * n * n) return -1;
    const int s = u.index_map[loc];
    return (s < 0 || s >= u.n_orb) ? -1 : s;
}

__global__ void __launch_bounds__(256)
    build_block_kernel(pmb_ueg_t u, const double *__restrict__ W0a, const double *__restrict__ W1a,
                       const double *__restrict__ W0s, BlockGeom g, double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long rows = (long long)g.ext[0] * g.ext[1] * g.ext[2];
    for (long long row = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < rows; row += warps) {
        const int r = g.lo[2] + (int)(row % g.ext[2]);
        const int q = g.lo[1] + (int)((row / g.ext[2]) % g.ext[1]);
        const int p = g.lo[0] + (int)(row / ((long long)g.ext[2] * g.ext[1]));
        const int s = conserving_s(u, p, q, r);
        double w = 0.0;
        const int sl = s - g.lo[3];
        if (s >= 0 && sl >= 0 && sl < g.ext[3]) w = ueg_value(u, W0a, W1a, W0s, p, q, r, s);
        double *dst = out + row * (long long)g.ext[3];
        for (int c = lane; c < g.ext[3]; c += 32) dst[c] = (c == sl) ? w : 0.0;
    }
}

// one thread per (p,q,r): the candidate non-zero of that dense row (coalesced 8 B stores)
__global__ void __launch_bounds__(256)
    build_nz_kernel(pmb_ueg_t u, const double *__restrict__ W0a, const double *__restrict__ W1a,
                    const double *__restrict__ W0s, BlockGeom g, double *__restrict__ out) {
    const long long rows = (long long)g.ext[0] * g.ext[1] * g.ext[2];
    for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < rows;
         row += (long long)gridDim.x * blockDim.x) {
        const int r = g.lo[2] + (int)(row % g.ext[2]);
        const int q = g.lo[1] + (int)((row / g.ext[2]) % g.ext[1]);
        const int p = g.lo[0] + (int)(row / ((long long)g.ext[2] * g.ext[1]));
        const int s = conserving_s(u, p, q, r);
        const int sl = s - g.lo[3];
        double w = 0.0;
        if (s >= 0 && sl >= 0 && sl < g.ext[3]) w = ueg_value(u, W0a, W1a, W0s, p, q, r, s);
        out[row] = w;
    }
}

}  // namespace pmb

using namespace pmb;

extern "C" int pmb_ueg_build_nz(const pmb_ueg_t *u, const double *W0a, const double *W1a, const double *W0s,
                                const int32_t lo[4], const int32_t ext[4], double *out, pmb_stream_t stream) {
    if (!u || !u->kvec || !u->kp || !u->index_map || !lo || !ext || !out) return PMB_E_BADARG;
    BlockGeom g;
    long long rows = 1;
    for (int d = 0; d < 4; ++d) {
        if (lo[d] < 0 || ext[d] <= 0 || lo[d] + ext[d] > u->n_orb) return PMB_E_BADARG;
        g.lo[d] = lo[d];
        g.ext[d] = ext[d];
        if (d < 3) rows *= ext[d];
    }
    long long blocks = (rows + 255) / 256;
    const long long cap = (long long)kSmCount * 16;
    if (blocks > cap) blocks = cap;
    build_nz_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*u, W0a, W1a, W0s, g, out);
    count_launch();
    return cuda_status();
}

extern "C" int pmb_ueg_umat(const pmb_ueg_t *u, double box_len, int cutoff, int nq, const int32_t *qvec,
                            double *out, pmb_stream_t stream) {
    if (!u || !u->u_table || cutoff < 0 || cutoff > 200 || nq <= 0 || !qvec || !out) return PMB_E_BADARG;
    umat_kernel<<<nq, 256, 0, (cudaStream_t)stream>>>(box_len, cutoff, u->omega, u->u_table, u->u_table_len,
                                                      qvec, out);
    count_launch();
    return cuda_status();
}

extern "C" int pmb_ueg_pair_tables(const pmb_ueg_t *u, int mode, const double *umat_pr, double *W0,
                                   double *W1, pmb_stream_t stream) {
    if (!u || !u->kp || !u->kvec || !W0 || mode < 0 || mode > PMB_UEG_EXCHANGE_3 || u->n_orb <= 0) return PMB_E_BADARG;
    if ((mode == PMB_UEG_ONLY_2B || mode == PMB_UEG_ONLY_HERMI_2B) && !umat_pr) return PMB_E_BADARG;
    if (mode != PMB_UEG_COULOMB && !u->u_table) return PMB_E_BADARG;
    const int n = u->n_orb * u->n_orb;
    pair_tables_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*u, mode, umat_pr, W0, W1);
    count_launch();
    return cuda_status();
}

extern "C" int pmb_ueg_build_block(const pmb_ueg_t *u, const double *W0a, const double *W1a,
                                   const double *W0s, const int32_t lo[4], const int32_t ext[4], double *out,
                                   pmb_stream_t stream) {
    if (!u || !u->kvec || !u->kp || !u->index_map || !lo || !ext || !out) return PMB_E_BADARG;
    BlockGeom g;
    long long rows = 1;
    for (int d = 0; d < 4; ++d) {
        if (lo[d] < 0 || ext[d] <= 0 || lo[d] + ext[d] > u->n_orb) return PMB_E_BADARG;
        g.lo[d] = lo[d];
        g.ext[d] = ext[d];
        if (d < 3) rows *= ext[d];
    }
    long long blocks = (rows + 7) / 8;  // 8 warps per CTA, one row per warp per step
    const long long cap = (long long)kSmCount * 16;
    if (blocks > cap) blocks = cap;
    build_block_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*u, W0a, W1a, W0s, g, out);
    count_launch();
    return cuda_status();
}
