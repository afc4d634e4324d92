// HBM-bound elementwise / reduction kernels of the coupled-cluster iteration:
// MP2 start amplitudes, fused amplitude update, energy, T-tilde, the
// Ex + Ex^{baji} permutation-add, DIIS dot products and linear combinations.
// All of them are one pass over T2-sized data with coalesced accesses along the
// fastest (ij) index; reductions are deterministic (two fixed-order stages).
#include "common.cuh"

namespace pmb {

long long g_launch_count = 0;

__global__ void finish_reduce_kernel(const double *ws, int nblocks, int ns, double *out,
                                     int accumulate) {
    __shared__ double sh[kMaxScalars][kReduceThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < ns; ++k) {
        double x = 0.0;
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x) x += ws[(size_t)b * ns + k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sh[k][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < ns) {
        double x = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) x += sh[threadIdx.x][w];
        out[threadIdx.x] = accumulate ? out[threadIdx.x] + x : x;
    }
}

int finish_reduce(const double *ws, int nblocks, int ns, double *out, int accumulate,
                  cudaStream_t s) {
    finish_reduce_kernel<<<1, kReduceThreads, 0, s>>>(ws, nblocks, ns, out, accumulate);
    count_launch();
    return cuda_status();
}

static inline int grid_for(size_t n, int threads, int cap) {
    size_t b = (n + threads - 1) / threads;
    if (b > (size_t)cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ---------------------------------------------------------------------------
struct Str4 {
    long long e[4];
    long long si[4];
    long long so[4];
};

// One pass of `kEwCopy` independent elements per thread (bytes in flight, not instruction count, is
// what a strided copy needs); the two slow indices are split off once per CTA pass with 64-bit
// arithmetic, the two fast ones per element in 32 bits.
constexpr int kEwCopy = 4;
__global__ void __launch_bounds__(256) axpby4_kernel(Str4 g, double alpha, const double *__restrict__ in,
                                                     double beta, double *out) {
    const size_t n = (size_t)g.e[0] * g.e[1] * g.e[2] * g.e[3];
    const unsigned e3 = (unsigned)g.e[3], e23 = (unsigned)(g.e[2] * g.e[3]);
    const bool small_inner = g.e[2] * g.e[3] < (1LL << 31);
    const size_t pass = (size_t)blockDim.x * kEwCopy;
    for (size_t base = (size_t)blockIdx.x * pass; base < n; base += (size_t)gridDim.x * pass) {
        double v[kEwCopy];
        long long oo[kEwCopy];
        const size_t r01 = small_inner ? base / e23 : 0;            // CTA-uniform
        const unsigned rem0 = small_inner ? (unsigned)(base - r01 * e23) : 0;
#pragma unroll
        for (int u = 0; u < kEwCopy; ++u) {
            const size_t idx = base + (size_t)u * blockDim.x + threadIdx.x;
            v[u] = 0.0;
            oo[u] = -1;
            if (idx < n) {
                long long i0, i1, i2, i3;
                if (small_inner) {
                    const unsigned local = rem0 + u * blockDim.x + threadIdx.x;      // < e23 + pass
                    const unsigned d01 = local / e23, r = local - d01 * e23;
                    const size_t q = r01 + d01;
                    i0 = (long long)(q / (size_t)g.e[1]);
                    i1 = (long long)(q - (size_t)i0 * g.e[1]);
                    i2 = r / e3;
                    i3 = r - (unsigned)i2 * e3;
                } else {
                    size_t r = idx;
                    i3 = r % g.e[3];
                    r /= g.e[3];
                    i2 = r % g.e[2];
                    r /= g.e[2];
                    i1 = r % g.e[1];
                    i0 = r / g.e[1];
                }
                v[u] = alpha * in[i0 * g.si[0] + i1 * g.si[1] + i2 * g.si[2] + i3 * g.si[3]];
                oo[u] = i0 * g.so[0] + i1 * g.so[1] + i2 * g.so[2] + i3 * g.so[3];
            }
        }
        if (beta != 0.0) {
            double o[kEwCopy];
#pragma unroll
            for (int u = 0; u < kEwCopy; ++u) o[u] = oo[u] >= 0 ? out[oo[u]] : 0.0;
#pragma unroll
            for (int u = 0; u < kEwCopy; ++u) v[u] += beta * o[u];
        }
#pragma unroll
        for (int u = 0; u < kEwCopy; ++u)
            if (oo[u] >= 0) out[oo[u]] = v[u];
    }
}

// Transposing form of the same copy: the input is unit-stride along dimension `dt` (one of 0..2),
// the output along dimension 3.  32 x 32 tiles over (dt, 3) go through shared memory so that both
// the reads (lanes along dt) and the writes (lanes along 3) are full 256 B row segments; the two
// remaining dimensions are the batch.  (The element-per-thread kernel reads one 32 B sector per
// 8 B word here: 0.25 of the HBM rate, tools/bench_hbm.py.)
__global__ void __launch_bounds__(256) axpby4_transpose_kernel(Str4 g, int dt, double alpha,
                                                               const double *__restrict__ in, double beta,
                                                               double *out) {
    __shared__ double tile[32][33];
    int ob[2], nb = 0;                         // the two batch dimensions
    for (int d = 0; d < 3; ++d)
        if (d != dt) ob[nb++] = d;
    const long long tt = (g.e[dt] + 31) / 32, t3 = (g.e[3] + 31) / 32;
    const long long ntile = tt * t3 * g.e[ob[0]] * g.e[ob[1]];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;        // 32 x 8
    for (long long t = blockIdx.x; t < ntile; t += gridDim.x) {
        long long r = t;
        const long long c3 = (r % t3) * 32;
        r /= t3;
        const long long ct = (r % tt) * 32;
        r /= tt;
        const long long b1 = r % g.e[ob[1]], b0 = r / g.e[ob[1]];
        const long long ibase = b0 * g.si[ob[0]] + b1 * g.si[ob[1]], obase = b0 * g.so[ob[0]] + b1 * g.so[ob[1]];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {          // read: lanes along dt, rows along 3
            const long long i3 = c3 + ly + 8 * k, it = ct + lx;
            if (i3 < g.e[3] && it < g.e[dt]) tile[ly + 8 * k][lx] = in[ibase + it * g.si[dt] + i3 * g.si[3]];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {          // write: lanes along 3, rows along dt
            const long long it = ct + ly + 8 * k, i3 = c3 + lx;
            if (i3 < g.e[3] && it < g.e[dt]) {
                double *d = out + obase + it * g.so[dt] + i3 * g.so[3];
                const double v = alpha * tile[lx][ly + 8 * k];
                *d = (beta != 0.0) ? v + beta * (*d) : v;
            }
        }
    }
}

// Index split of one element of an [a,b,i,j] row block, 32-bit arithmetic: the CTA-uniform
// 64-bit division (first element of the pass -> pair ab0, remainder rem0) is done once per
// pass, the per-element part fits in 32 bits.
struct Abij {
    int a, b, i, j;
};
__device__ __forceinline__ Abij split_abij(size_t ab0, unsigned local, unsigned oo, unsigned no, unsigned nv) {
    const unsigned dab = local / oo, ij = local - dab * oo;
    const size_t ab = ab0 + dab;
    Abij r;
    r.a = (int)(ab / nv);
    r.b = (int)(ab - (size_t)r.a * nv);
    r.i = (int)(ij / no);
    r.j = (int)(ij - (unsigned)r.i * no);
    return r;
}

constexpr int kEwUnroll = 4;   // independent 8-byte loads in flight per thread and stream

// T2 = V_abij / (e_i + e_j - e_a - e_b + shift)
__global__ void __launch_bounds__(256)
    mp2_kernel(int no, int nv, int a_lo, int na, const double *__restrict__ ei, const double *__restrict__ ea,
               double shift, const double *__restrict__ V, Str4 g, double *__restrict__ T2) {
    const size_t n = (size_t)na * nv * no * no;
    const unsigned oo = (unsigned)no * no;
    const size_t pass = (size_t)blockDim.x * kEwUnroll;
    for (size_t base = (size_t)blockIdx.x * pass; base < n; base += (size_t)gridDim.x * pass) {
        const size_t ab0 = base / oo;
        const unsigned rem0 = (unsigned)(base - ab0 * oo);
        double v[kEwUnroll], d[kEwUnroll];
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const unsigned off = u * blockDim.x + threadIdx.x;
            v[u] = 0.0;
            d[u] = 1.0;
            if (base + off < n) {
                const Abij x = split_abij(ab0, rem0 + off, oo, no, nv);
                v[u] = V[x.a * g.si[0] + x.b * g.si[1] + x.i * g.si[2] + x.j * g.si[3]];
                d[u] = ei[x.i] + ei[x.j] - ea[a_lo + x.a] - ea[x.b] + shift;
            }
        }
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const size_t idx = base + u * blockDim.x + threadIdx.x;
            if (idx < n) T2[idx] = v[u] / d[u];
        }
    }
}

// dT = R * (1/D); T += delta*dT; partial sum of dT^2
__global__ void __launch_bounds__(kReduceThreads)
    update_doubles_kernel(int no, int nv, int a_lo, int na, const double *__restrict__ ei,
                          const double *__restrict__ ea, double shift, double delta, int denom_mode,
                          const double *__restrict__ R, double *__restrict__ dT, double *__restrict__ T2,
                          double *ws) {
    const size_t n = (size_t)na * nv * no * no;
    const unsigned oo = (unsigned)no * no;
    const size_t pass = (size_t)blockDim.x * kEwUnroll;
    double s[1] = {0.0};
    for (size_t base = (size_t)blockIdx.x * pass; base < n; base += (size_t)gridDim.x * pass) {
        const size_t ab0 = base / oo;
        const unsigned rem0 = (unsigned)(base - ab0 * oo);
        double r[kEwUnroll], t[kEwUnroll];
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const size_t idx = base + u * blockDim.x + threadIdx.x;
            r[u] = idx < n ? R[idx] : 0.0;
            t[u] = idx < n ? T2[idx] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) {
            const unsigned off = u * blockDim.x + threadIdx.x;
            const size_t idx = base + off;
            if (idx < n) {
                const Abij x = split_abij(ab0, rem0 + off, oo, no, nv);
                // denom_mode 1: the PRODUCT e_i e_j (-e_a)(-e_b) that the reference's Brueckner
                // branch builds with einsum('i,j,a,b->abij') (ccd.py:118), in its order
                const double den = denom_mode ? ((ei[x.i] * ei[x.j]) * (-ea[a_lo + x.a])) * (-ea[x.b])
                                              : ei[x.i] + ei[x.j] - ea[a_lo + x.a] - ea[x.b];
                const double dinv = 1.0 / (den + shift);
                const double d = r[u] * dinv;
                dT[idx] = d;
                T2[idx] = t[u] + delta * d;
                s[0] += d * d;
            }
        }
    }
    block_reduce_store<1>(s, ws);
}

__global__ void __launch_bounds__(256)
    update_singles_kernel(int no, int nv, const double *__restrict__ ei, const double *__restrict__ ea,
                          double shift, double delta, const double *__restrict__ R1, double *__restrict__ dT1,
                          double *__restrict__ T1) {
    const int n = nv * no;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int i = idx % no, a = idx / no;
        const double d = R1[idx] * (1.0 / (ei[i] - ea[a] + shift));
        dT1[idx] = d;
        T1[idx] += delta * d;
    }
}

// energy: one CTA handles a set of (a,b) pairs; for each pair the o x o block
// of T2 is contiguous, V[i,j,a,b] and V[i,j,b,a] are gathered.
__global__ void __launch_bounds__(kReduceThreads)
    energy_kernel(int no, int nv, int a_lo, int na, const double *__restrict__ T2,
                  const double *__restrict__ T1, const double *__restrict__ V, Str4 g, int mp2_form,
                  double *ws) {
    const size_t n = (size_t)na * nv * no * no;
    double s[3] = {0.0, 0.0, 0.0};
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx;
        const int j = r % no;
        r /= no;
        const int i = r % no;
        r /= no;
        const int b = r % nv;
        const int a = a_lo + (int)(r / nv);
        const double t = T2[idx];
        double tau = t;
        if (T1) tau += T1[a * no + i] * T1[b * no + j];
        const double vd = V[i * g.si[0] + j * g.si[1] + a * g.si[2] + b * g.si[3]];
        // exchange: ccd.py:261 uses V[i,j,b,a]; mp2.py:20 uses V[j,i,a,b]
        const double vx = mp2_form ? V[j * g.si[0] + i * g.si[1] + a * g.si[2] + b * g.si[3]]
                                   : V[i * g.si[0] + j * g.si[1] + b * g.si[2] + a * g.si[3]];
        s[0] += tau * vd;
        s[1] += tau * vx;
        s[2] += t * t;
    }
    block_reduce_store<3>(s, ws);
}

// Tiled form of the same sums for the ccd.py exchange (V[i,j,b,a]).  A CTA owns one row a,
// 32 columns b and 32 pairs ij per pass: the V[ij,a,b] tile is read with b along the lanes
// (the unit-stride direction of V_ijab), T2[a,b,ij] with ij along the lanes, and the tile is
// turned through shared memory -- every global access of the direct term is a full 256 B row
// segment.  The exchange tile V[ij,b,a] has stride nv between lanes; consecutive tiles in the
// launch order differ in a only, so its 32 B sectors are shared through L2 by the CTAs that
// run side by side (a .. a+3 live in one sector).
constexpr int kEt = 32;
__global__ void __launch_bounds__(kReduceThreads)
    energy_tiled_kernel(int no, int nv, int a_lo, int na, const double *__restrict__ T2,
                        const double *__restrict__ T1, const double *__restrict__ V, Str4 g, double *ws) {
    __shared__ double Vd[kEt][kEt + 1], Vx[kEt][kEt + 1];
    const int oo = no * no;
    const int tb = (nv + kEt - 1) / kEt, tij = (oo + kEt - 1) / kEt;
    const long long tiles = (long long)na * tb * tij;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = kReduceThreads / 32;
    double s[3] = {0.0, 0.0, 0.0};
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int al = (int)(t % na);
        const long long rest = t / na;
        const int b0 = (int)(rest % tb) * kEt, ij0 = (int)(rest / tb) * kEt;
        const int a = a_lo + al;
        const int b = b0 + lane;
        __syncthreads();               // previous pass has consumed the tiles
#pragma unroll
        for (int k = 0; k < kEt / NW; ++k) {
            const int r = warp + k * NW, ij = ij0 + r;
            double vd = 0.0, vx = 0.0;
            if (ij < oo && b < nv) {
                const int i = ij / no, j = ij - i * no;
                const long long o = i * g.si[0] + j * g.si[1];
                vd = V[o + a * g.si[2] + b * g.si[3]];
                vx = V[o + b * g.si[2] + a * g.si[3]];
            }
            Vd[r][lane] = vd;
            Vx[r][lane] = vx;
        }
        __syncthreads();
        const int ij = ij0 + lane;
        const int i = ij / no, j = ij - i * no;
#pragma unroll
        for (int k = 0; k < kEt / NW; ++k) {
            const int r = warp + k * NW, bb = b0 + r;
            if (ij < oo && bb < nv) {
                const double tv = T2[((size_t)al * nv + bb) * oo + ij];
                double tau = tv;
                if (T1) tau += T1[a * no + i] * T1[bb * no + j];
                s[0] += tau * Vd[lane][r];
                s[1] += tau * Vx[lane][r];
                s[2] += tv * tv;
            }
        }
    }
    block_reduce_store<3>(s, ws);
}

// The same sums when V_ijab is STORED as [a,b,i,j] (the solvers keep such a copy of this static
// block: element strides (no, 1, nv o^2, o^2) for the indices (i,j,a,b)): the direct term reads row
// (a,b) of it, the exchange term row (b,a), both contiguous like the row of T2 -- a pure stream,
// 24 B per element.  One warp per row, 4 rows per pass and CTA.
__global__ void __launch_bounds__(kReduceThreads)
    energy_rows_kernel(int no, int nv, int a_lo, int na, const double *__restrict__ T2,
                       const double *__restrict__ T1, const double *__restrict__ Vt, double *ws) {
    const int oo = no * no;
    const long long rows = (long long)na * nv;
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    double s[3] = {0.0, 0.0, 0.0};
    for (long long r = warp0; r < rows; r += nwarps) {
        const int al = (int)(r / nv), b = (int)(r - (long long)al * nv), a = a_lo + al;
        const double *t = T2 + (size_t)r * oo;
        const double *vd = Vt + ((size_t)a * nv + b) * oo;
        const double *vx = Vt + ((size_t)b * nv + a) * oo;
        for (int e = lane; e < oo; e += 32) {
            const double tv = t[e];
            double tau = tv;
            if (T1) {
                const int i = e / no, j = e - i * no;
                tau += T1[a * no + i] * T1[b * no + j];
            }
            s[0] += tau * vd[e];
            s[1] += tau * vx[e];
            s[2] += tv * tv;
        }
    }
    block_reduce_store<3>(s, ws);
}

__global__ void scale3_kernel(double *scal) {
    if (threadIdx.x == 0) {
        scal[0] *= 2.0;
        scal[1] *= -1.0;
    }
}

// Tt[a,b,ij] = 2 T[a,b,ij] - T[b,a,ij].  A CTA owns the PAIR of rows (a,b), (b,a) (each o^2
// contiguous doubles): both are read once and both outputs written, 16 B per element of HBM
// traffic -- the element-per-thread form reads every row twice (24 B).  swap_ij = 1 is
// 2 T[a,b,i,j] - T[a,b,j,i]: a transpose inside one row, turned through shared memory.
constexpr int kRowMax = 4096;      // o^2 <= 4096 (o <= 64) rows live in shared memory
__global__ void __launch_bounds__(256)
    tilde_pair_kernel(int no, int nv, const double *__restrict__ T2, double *__restrict__ Tt, int swap_ij) {
    extern __shared__ double row_sh[];
    const int oo = no * no;
    if (swap_ij) {
        for (long long ab = blockIdx.x; ab < (long long)nv * nv; ab += gridDim.x) {
            const double *src = T2 + (size_t)ab * oo;
            __syncthreads();
            for (int e = threadIdx.x; e < oo; e += blockDim.x) row_sh[e] = src[e];
            __syncthreads();
            for (int e = threadIdx.x; e < oo; e += blockDim.x) {
                const int i = e / no, j = e - i * no;
                Tt[(size_t)ab * oo + e] = 2.0 * row_sh[e] - row_sh[j * no + i];
            }
        }
        return;
    }
    const long long npair = (long long)nv * (nv + 1) / 2;
    for (long long pr = blockIdx.x; pr < npair; pr += gridDim.x) {
        // pair index -> (a <= b): row a of the upper triangle starts at a nv - a (a - 1) / 2
        long long a = (long long)((2.0 * nv + 1.0 - sqrt((2.0 * nv + 1.0) * (2.0 * nv + 1.0) - 8.0 * (double)pr)) * 0.5);
        while (a > 0 && a * nv - a * (a - 1) / 2 > pr) --a;
        while ((a + 1) * nv - (a + 1) * a / 2 <= pr) ++a;
        const long long b = a + (pr - (a * nv - a * (a - 1) / 2));
        const size_t r1 = ((size_t)a * nv + b) * oo, r2 = ((size_t)b * nv + a) * oo;
        for (int e = threadIdx.x; e < oo; e += blockDim.x) {
            const double x = T2[r1 + e], y = T2[r2 + e];
            Tt[r1 + e] = 2.0 * x - y;
            if (a != b) Tt[r2 + e] = 2.0 * y - x;
        }
    }
}

// general-size fallback (o^2 > kRowMax): one element per thread
__global__ void __launch_bounds__(256)
    tilde_kernel(int no, int nv, const double *__restrict__ T2, double *__restrict__ Tt, int swap_ij) {
    const size_t n = (size_t)nv * nv * no * no;
    const size_t oo = (size_t)no * no;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t ij = idx % oo;
        const size_t ab = idx / oo;
        size_t src;
        if (swap_ij) {
            const size_t i = ij / no, j = ij % no;
            src = ab * oo + j * no + i;
        } else {
            const size_t a = ab / nv, b = ab % nv;
            src = (b * nv + a) * oo + ij;
        }
        Tt[idx] = 2.0 * T2[idx] - T2[src];
    }
}

// R[a,b,i,j] (+)= Ex[a,b,i,j] + Ex[b,a,j,i], again by row pairs: X = Ex[a,b,:], Y = Ex[b,a,:] are
// staged in shared memory (their (ji) transposes are read from there), each is read from HBM once:
// 8 B (Ex) + 8/16 B (R written / updated) per element.
__global__ void __launch_bounds__(256)
    sym_baji_pair_kernel(int no, int nv, const double *__restrict__ Ex, double *__restrict__ R, int accumulate) {
    extern __shared__ double row_sh[];
    const int oo = no * no;
    double *X = row_sh, *Y = row_sh + oo;
    const long long npair = (long long)nv * (nv + 1) / 2;
    for (long long pr = blockIdx.x; pr < npair; pr += gridDim.x) {
        long long a = (long long)((2.0 * nv + 1.0 - sqrt((2.0 * nv + 1.0) * (2.0 * nv + 1.0) - 8.0 * (double)pr)) * 0.5);
        while (a > 0 && a * nv - a * (a - 1) / 2 > pr) --a;
        while ((a + 1) * nv - (a + 1) * a / 2 <= pr) ++a;
        const long long b = a + (pr - (a * nv - a * (a - 1) / 2));
        const size_t r1 = ((size_t)a * nv + b) * oo, r2 = ((size_t)b * nv + a) * oo;
        __syncthreads();
        for (int e = threadIdx.x; e < oo; e += blockDim.x) {
            X[e] = Ex[r1 + e];
            Y[e] = Ex[r2 + e];
        }
        __syncthreads();
        for (int e = threadIdx.x; e < oo; e += blockDim.x) {
            const int i = e / no, j = e - i * no, t = j * no + i;
            const double v1 = X[e] + Y[t];
            R[r1 + e] = accumulate ? R[r1 + e] + v1 : v1;
            if (a != b) {
                const double v2 = Y[e] + X[t];
                R[r2 + e] = accumulate ? R[r2 + e] + v2 : v2;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
    sym_baji_kernel(int no, int nv, const double *__restrict__ Ex, double *__restrict__ R, int accumulate) {
    const size_t n = (size_t)nv * nv * no * no;
    const size_t oo = (size_t)no * no;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t ij = idx % oo, ab = idx / oo;
        const size_t i = ij / no, j = ij % no;
        const size_t a = ab / nv, b = ab % nv;
        const double v = Ex[idx] + Ex[(b * nv + a) * oo + j * no + i];
        R[idx] = accumulate ? R[idx] + v : v;
    }
}

constexpr int kMaxVec = 16;
struct VecList {
    const double *p[kMaxVec];
    double c[kMaxVec];
};

template <int NV>
__global__ void __launch_bounds__(kReduceThreads)
    dots_kernel(VecList L, const double *__restrict__ Y, size_t n, double *ws) {
    double s[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) s[k] = 0.0;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (size_t)gridDim.x * blockDim.x) {
        const double y = Y[idx];
#pragma unroll
        for (int k = 0; k < NV; ++k) s[k] += L.p[k][idx] * y;
    }
    block_reduce_store<NV>(s, ws);
}

__global__ void __launch_bounds__(256)
    lincomb_kernel(VecList L, int nvec, size_t n, double beta, double *__restrict__ out) {
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (size_t)gridDim.x * blockDim.x) {
        double v = (beta != 0.0) ? beta * out[idx] : 0.0;
        for (int k = 0; k < nvec; ++k) v += L.c[k] * L.p[k][idx];
        out[idx] = v;
    }
}

// out[I] = beta*out[I] + alpha * sum_R A[I,R]*B[I,R]; one warp per output element when the
// reduction is long (lanes stride over R, fixed-order shuffle tree), one thread otherwise.
struct BdotDev {
    const double *A, *B;
    double *out;
    int ni, nr;
    long long I, R;
    long long i_ext[4], o_istr[4], a_istr[4], b_istr[4], r_ext[4], a_rstr[4], b_rstr[4];
    double alpha, beta;
};

__device__ __forceinline__ void bdot_offsets(const BdotDev &d, long long idx, long long &oo, long long &ao,
                                             long long &bo) {
    oo = ao = bo = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < d.ni) {
            const long long e = d.i_ext[k], q = idx / e, r = idx - q * e;
            oo += r * d.o_istr[k];
            ao += r * d.a_istr[k];
            bo += r * d.b_istr[k];
            idx = q;
        }
    }
}

__device__ __forceinline__ double bdot_term(const BdotDev &d, long long ridx, long long ao, long long bo) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < d.nr) {
            const long long e = d.r_ext[k], q = ridx / e, r = ridx - q * e;
            ao += r * d.a_rstr[k];
            bo += r * d.b_rstr[k];
            ridx = q;
        }
    }
    return d.A[ao] * d.B[bo];
}

__global__ void __launch_bounds__(256) bdot_thread_kernel(const __grid_constant__ BdotDev d) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < d.I;
         idx += (long long)gridDim.x * blockDim.x) {
        long long oo, ao, bo;
        bdot_offsets(d, idx, oo, ao, bo);
        double s = 0.0;
        for (long long r = 0; r < d.R; ++r) s += bdot_term(d, r, ao, bo);
        s *= d.alpha;
        if (d.beta != 0.0) s += d.beta * d.out[oo];
        d.out[oo] = s;
    }
}

__global__ void __launch_bounds__(256) bdot_warp_kernel(const __grid_constant__ BdotDev d) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long idx = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; idx < d.I; idx += warps) {
        long long oo, ao, bo;
        bdot_offsets(d, idx, oo, ao, bo);
        double s = 0.0;
        for (long long r = lane; r < d.R; r += 32) s += bdot_term(d, r, ao, bo);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            s *= d.alpha;
            if (d.beta != 0.0) s += d.beta * d.out[oo];
            d.out[oo] = s;
        }
    }
}

// ---------------------------------------------------------------------------
// out[X] = beta out[X] + alpha sum_K vec[K] B[K,X]   (pmb_gemv)
// ---------------------------------------------------------------------------
struct GemvDev {
    const double *vec, *B;
    double *out;
    int nk, nx;
    long long K, X;
    long long k_ext[4], v_kstr[4], b_kstr[4], x_ext[4], b_xstr[4], o_xstr[4];
    double alpha, beta;
};

__device__ __forceinline__ void gemv_x_offsets(const GemvDev &d, long long x, long long &bo, long long &oo) {
    bo = oo = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < d.nx) {
            const long long e = d.x_ext[k], q = x / e, r = x - q * e;
            bo += r * d.b_xstr[k];
            oo += r * d.o_xstr[k];
            x = q;
        }
    }
}

// B unit-stride along X: one thread per output, consecutive threads read consecutive words
// of every B row; the k loop is a counter nest (no divisions), 4 rows in flight per thread.
__global__ void __launch_bounds__(256) gemv_xfast_kernel(const __grid_constant__ GemvDev d) {
    const long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= d.X) return;
    long long bo, oo;
    gemv_x_offsets(d, x, bo, oo);
    const double *b = d.B + bo;
    double acc = 0.0;
    for (long long k3 = 0; k3 < d.k_ext[3]; ++k3)
        for (long long k2 = 0; k2 < d.k_ext[2]; ++k2)
            for (long long k1 = 0; k1 < d.k_ext[1]; ++k1) {
                const double *bb = b + k3 * d.b_kstr[3] + k2 * d.b_kstr[2] + k1 * d.b_kstr[1];
                const double *vv = d.vec + k3 * d.v_kstr[3] + k2 * d.v_kstr[2] + k1 * d.v_kstr[1];
                const long long n0 = d.k_ext[0];
                // 8 predicated loads issued back to back per pass (bytes in flight, not
                // instruction count, is what an HBM-bound loop needs)
                for (long long base = 0; base < n0; base += 8) {
                    double bv[8], vv8[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const long long k0 = base + u;
                        const bool in = k0 < n0;
                        bv[u] = in ? bb[k0 * d.b_kstr[0]] : 0.0;
                        vv8[u] = in ? __ldg(vv + k0 * d.v_kstr[0]) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc += vv8[u] * bv[u];
                }
            }
    acc *= d.alpha;
    if (d.beta != 0.0) acc += d.beta * d.out[oo];
    d.out[oo] = acc;
}

// B unit-stride along the first K index: one warp per output, lanes walk that index (256 B
// per warp instruction, 4 instructions in flight), fixed-order shuffle tree at the end.
__global__ void __launch_bounds__(256) gemv_kfast_kernel(const __grid_constant__ GemvDev d) {
    const int lane = threadIdx.x & 31;
    const long long x = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (x >= d.X) return;
    long long bo, oo;
    gemv_x_offsets(d, x, bo, oo);
    const double *b = d.B + bo;
    double acc = 0.0;
    const long long n0 = d.k_ext[0];
    for (long long k3 = 0; k3 < d.k_ext[3]; ++k3)
        for (long long k2 = 0; k2 < d.k_ext[2]; ++k2)
            for (long long k1 = 0; k1 < d.k_ext[1]; ++k1) {
                const double *bb = b + k3 * d.b_kstr[3] + k2 * d.b_kstr[2] + k1 * d.b_kstr[1];
                const double *vv = d.vec + k3 * d.v_kstr[3] + k2 * d.v_kstr[2] + k1 * d.v_kstr[1];
                long long k0 = lane;
                for (; k0 + 96 < n0; k0 += 128) {
                    const double b0 = bb[k0 * d.b_kstr[0]], b1 = bb[(k0 + 32) * d.b_kstr[0]],
                                 b2 = bb[(k0 + 64) * d.b_kstr[0]], b3 = bb[(k0 + 96) * d.b_kstr[0]];
                    acc += __ldg(vv + k0 * d.v_kstr[0]) * b0;
                    acc += __ldg(vv + (k0 + 32) * d.v_kstr[0]) * b1;
                    acc += __ldg(vv + (k0 + 64) * d.v_kstr[0]) * b2;
                    acc += __ldg(vv + (k0 + 96) * d.v_kstr[0]) * b3;
                }
                for (; k0 < n0; k0 += 32) acc += __ldg(vv + k0 * d.v_kstr[0]) * bb[k0 * d.b_kstr[0]];
            }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        acc *= d.alpha;
        if (d.beta != 0.0) acc += d.beta * d.out[oo];
        d.out[oo] = acc;
    }
}

__global__ void __launch_bounds__(256)
    cdiv_shifted_kernel(size_t n, const double *__restrict__ diag, double zr, double zi, double shift,
                        const double *xr, const double *xi, double *yr, double *yi) {
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (size_t)gridDim.x * blockDim.x) {
        const double dr = zr - diag[idx] + shift;
        const double inv = 1.0 / (dr * dr + zi * zi);
        const double mr = dr * inv, mi = -zi * inv;      // 1 / (dr + i zi)
        const double a = xr[idx], b = xi[idx];
        yr[idx] = mr * a - mi * b;
        yi[idx] = mr * b + mi * a;
    }
}

static Str4 make_str(const int64_t ext[4], const int64_t in_str[4], const int64_t out_str[4]) {
    Str4 g;
    for (int d = 0; d < 4; ++d) {
        g.e[d] = ext ? ext[d] : 0;
        g.si[d] = in_str ? in_str[d] : 0;
        g.so[d] = out_str ? out_str[d] : 0;
    }
    return g;
}

}  // namespace pmb

using namespace pmb;

extern "C" size_t pmb_reduce_workspace(void) {
    return sizeof(double) * (size_t)kReduceBlocks * kMaxVec;
}

extern "C" int pmb_axpby4(const int64_t ext[4], double alpha, const double *in, const int64_t in_str[4],
                          double beta, double *out, const int64_t out_str[4], pmb_stream_t stream) {
    if (!ext || !in || !out || !in_str || !out_str) return PMB_E_BADARG;
    size_t n = 1;
    for (int d = 0; d < 4; ++d) {
        if (ext[d] <= 0) return PMB_E_BADARG;
        n *= (size_t)ext[d];
    }
    // a transpose (input unit-stride along one of the slow output dimensions): tiled kernel
    int dt = -1;
    if (out_str[3] == 1 && in_str[3] != 1 && ext[3] >= 16)
        for (int d = 0; d < 3; ++d)
            if (in_str[d] == 1 && ext[d] >= 16) dt = d;
    if (dt >= 0) {
        const size_t tiles = n / 1024 + 1;
        axpby4_transpose_kernel<<<grid_for(tiles, 1, 32 * kSmCount), 256, 0, (cudaStream_t)stream>>>(
            make_str(ext, in_str, out_str), dt, alpha, in, beta, out);
    } else {
        axpby4_kernel<<<grid_for(n, 256, 16 * kSmCount), 256, 0, (cudaStream_t)stream>>>(
            make_str(ext, in_str, out_str), alpha, in, beta, out);
    }
    count_launch();
    return cuda_status();
}

static bool bad_rows(int nv, int a_lo, int na) { return a_lo < 0 || na <= 0 || a_lo + na > nv; }

extern "C" int pmb_mp2_amplitudes(int no, int nv, int a_lo, int na, const double *eps_i, const double *eps_a,
                                  double shift, const double *V_abij, const int64_t v_str[4], double *T2,
                                  pmb_stream_t stream) {
    if (no <= 0 || nv <= 0 || bad_rows(nv, a_lo, na) || !eps_i || !eps_a || !V_abij || !v_str || !T2)
        return PMB_E_BADARG;
    const size_t n = (size_t)na * nv * no * no;
    mp2_kernel<<<grid_for(n, 256, 16 * kSmCount), 256, 0, (cudaStream_t)stream>>>(
        no, nv, a_lo, na, eps_i, eps_a, shift, V_abij, make_str(nullptr, v_str, nullptr), T2);
    count_launch();
    return cuda_status();
}

extern "C" int pmb_update_doubles(int no, int nv, int a_lo, int na, const double *eps_i, const double *eps_a,
                                  double shift, double delta, int denom_mode, const double *R, double *dT,
                                  double *T2, double *scal, void *ws, size_t ws_bytes, pmb_stream_t stream) {
    if (no <= 0 || nv <= 0 || bad_rows(nv, a_lo, na) || !eps_i || !eps_a || !R || !dT || !T2 || !scal)
        return PMB_E_BADARG;
    if (denom_mode != 0 && denom_mode != 1) return PMB_E_BADARG;
    if (!ws || ws_bytes < pmb_reduce_workspace()) return PMB_E_WORKSPACE;
    const size_t n = (size_t)na * nv * no * no;
    const int blocks = grid_for(n, kReduceThreads, kReduceBlocks);
    update_doubles_kernel<<<blocks, kReduceThreads, 0, (cudaStream_t)stream>>>(
        no, nv, a_lo, na, eps_i, eps_a, shift, delta, denom_mode, R, dT, T2, (double *)ws);
    count_launch();
    int rc = cuda_status();
    if (rc) return rc;
    return finish_reduce((double *)ws, blocks, 1, scal, 0, (cudaStream_t)stream);
}

extern "C" int pmb_update_singles(int no, int nv, const double *eps_i, const double *eps_a, double shift,
                                  double delta, const double *R1, double *dT1, double *T1,
                                  pmb_stream_t stream) {
    if (no <= 0 || nv <= 0 || !eps_i || !eps_a || !R1 || !dT1 || !T1) return PMB_E_BADARG;
    update_singles_kernel<<<grid_for((size_t)nv * no, 256, kSmCount), 256, 0, (cudaStream_t)stream>>>(
        no, nv, eps_i, eps_a, shift, delta, R1, dT1, T1);
    count_launch();
    return cuda_status();
}

extern "C" int pmb_energy_doubles(int no, int nv, int a_lo, int na, const double *T2, const double *T1,
                                  const double *V_ijab, const int64_t v_str[4], int mp2_form, double *scal,
                                  void *ws, size_t ws_bytes, pmb_stream_t stream) {
    if (no <= 0 || nv <= 0 || bad_rows(nv, a_lo, na) || !T2 || !V_ijab || !v_str || !scal) return PMB_E_BADARG;
    if (!ws || ws_bytes < pmb_reduce_workspace()) return PMB_E_WORKSPACE;
    const size_t n = (size_t)na * nv * no * no;
    const int blocks = grid_for(n, kReduceThreads, kReduceBlocks);
    const bool abij_storage = !mp2_form && v_str[1] == 1 && v_str[0] == no && v_str[3] == (int64_t)no * no &&
                              v_str[2] == (int64_t)nv * no * no;
    if (abij_storage)
        energy_rows_kernel<<<blocks, kReduceThreads, 0, (cudaStream_t)stream>>>(no, nv, a_lo, na, T2, T1, V_ijab,
                                                                                (double *)ws);
    else if (mp2_form)
        energy_kernel<<<blocks, kReduceThreads, 0, (cudaStream_t)stream>>>(
            no, nv, a_lo, na, T2, T1, V_ijab, make_str(nullptr, v_str, nullptr), mp2_form, (double *)ws);
    else
        energy_tiled_kernel<<<blocks, kReduceThreads, 0, (cudaStream_t)stream>>>(
            no, nv, a_lo, na, T2, T1, V_ijab, make_str(nullptr, v_str, nullptr), (double *)ws);
    count_launch();
    int rc = cuda_status();
    if (rc) return rc;
    rc = finish_reduce((double *)ws, blocks, 3, scal, 0, (cudaStream_t)stream);
    if (rc) return rc;
    scale3_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(scal);
    count_launch();
    return cuda_status();
}

extern "C" int pmb_tilde(int no, int nv, const double *T2, double *Tt, int swap_ij, pmb_stream_t stream) {
    if (no <= 0 || nv <= 0 || !T2 || !Tt || T2 == Tt) return PMB_E_BADARG;
    const size_t n = (size_t)nv * nv * no * no;
    if (no * no <= kRowMax) {
        const long long rows = swap_ij ? (long long)nv * nv : (long long)nv * (nv + 1) / 2;
        const unsigned blocks = (unsigned)(rows < 16LL * kSmCount ? rows : 16LL * kSmCount);
        tilde_pair_kernel<<<blocks, 256, swap_ij ? sizeof(double) * no * no : 0, (cudaStream_t)stream>>>(
            no, nv, T2, Tt, swap_ij);
    } else {
        tilde_kernel<<<grid_for(n, 256, 16 * kSmCount), 256, 0, (cudaStream_t)stream>>>(no, nv, T2, Tt, swap_ij);
    }
    count_launch();
    return cuda_status();
}

extern "C" int pmb_sym_baji(int no, int nv, const double *Ex, double *R, int accumulate,
                            pmb_stream_t stream) {
    if (no <= 0 || nv <= 0 || !Ex || !R || Ex == R) return PMB_E_BADARG;
    const size_t n = (size_t)nv * nv * no * no;
    if (2 * sizeof(double) * no * no <= 48 * 1024) {
        const long long rows = (long long)nv * (nv + 1) / 2;
        const unsigned blocks = (unsigned)(rows < 16LL * kSmCount ? rows : 16LL * kSmCount);
        sym_baji_pair_kernel<<<blocks, 256, 2 * sizeof(double) * no * no, (cudaStream_t)stream>>>(no, nv, Ex, R,
                                                                                               accumulate);
    } else {
        sym_baji_kernel<<<grid_for(n, 256, 16 * kSmCount), 256, 0, (cudaStream_t)stream>>>(no, nv, Ex, R, accumulate);
    }
    count_launch();
    return cuda_status();
}

extern "C" int pmb_dots(int nvec, const double *const *X, const double *Y, int64_t n, double *out,
                        void *ws, size_t ws_bytes, pmb_stream_t stream) {
    if (nvec < 1 || nvec > kMaxVec || !X || !Y || n <= 0 || !out) return PMB_E_BADARG;
    if (!ws || ws_bytes < pmb_reduce_workspace()) return PMB_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = grid_for((size_t)n, kReduceThreads, kReduceBlocks);
    int done = 0;
    while (done < nvec) {  // chunks of up to 4 vectors share one pass over Y
        VecList L;
        const int c = (nvec - done >= 4) ? 4 : (nvec - done);
        for (int k = 0; k < c; ++k) L.p[k] = X[done + k];
        switch (c) {
            case 4: dots_kernel<4><<<blocks, kReduceThreads, 0, s>>>(L, Y, (size_t)n, (double *)ws); break;
            case 3: dots_kernel<3><<<blocks, kReduceThreads, 0, s>>>(L, Y, (size_t)n, (double *)ws); break;
            case 2: dots_kernel<2><<<blocks, kReduceThreads, 0, s>>>(L, Y, (size_t)n, (double *)ws); break;
            default: dots_kernel<1><<<blocks, kReduceThreads, 0, s>>>(L, Y, (size_t)n, (double *)ws); break;
        }
        count_launch();
        int rc = cuda_status();
        if (rc) return rc;
        rc = finish_reduce((double *)ws, blocks, c, out + done, 0, s);
        if (rc) return rc;
        done += c;
    }
    return 0;
}

extern "C" int pmb_lincomb(int nvec, const double *c_host, const double *const *X, int64_t n, double beta,
                           double *out, pmb_stream_t stream) {
    if (nvec < 1 || nvec > kMaxVec || !c_host || !X || n <= 0 || !out) return PMB_E_BADARG;
    VecList L;
    for (int k = 0; k < nvec; ++k) {
        L.p[k] = X[k];
        L.c[k] = c_host[k];
    }
    lincomb_kernel<<<grid_for((size_t)n, 256, 16 * kSmCount), 256, 0, (cudaStream_t)stream>>>(
        L, nvec, (size_t)n, beta, out);
    count_launch();
    return cuda_status();
}

extern "C" int pmb_bdot(const pmb_bdot_t *d, pmb_stream_t stream) {
    if (!d || !d->A || !d->B || !d->out || d->ni < 0 || d->ni > 4 || d->nr < 0 || d->nr > 4) return PMB_E_BADARG;
    BdotDev k;
    k.A = d->A;
    k.B = d->B;
    k.out = d->out;
    k.ni = d->ni;
    k.nr = d->nr;
    k.alpha = d->alpha;
    k.beta = d->beta;
    k.I = k.R = 1;
    for (int i = 0; i < 4; ++i) {
        const bool ui = i < d->ni, ur = i < d->nr;
        if ((ui && d->i_ext[i] <= 0) || (ur && d->r_ext[i] <= 0)) return PMB_E_BADARG;
        k.i_ext[i] = ui ? d->i_ext[i] : 1;
        k.o_istr[i] = ui ? d->o_istr[i] : 0;
        k.a_istr[i] = ui ? d->a_istr[i] : 0;
        k.b_istr[i] = ui ? d->b_istr[i] : 0;
        k.r_ext[i] = ur ? d->r_ext[i] : 1;
        k.a_rstr[i] = ur ? d->a_rstr[i] : 0;
        k.b_rstr[i] = ur ? d->b_rstr[i] : 0;
        k.I *= k.i_ext[i];
        k.R *= k.r_ext[i];
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (k.R >= 64) {
        long long blocks = (k.I + 7) / 8;
        if (blocks > 32LL * kSmCount) blocks = 32LL * kSmCount;
        bdot_warp_kernel<<<(unsigned)blocks, 256, 0, s>>>(k);
    } else {
        bdot_thread_kernel<<<grid_for((size_t)k.I, 256, 16 * kSmCount), 256, 0, s>>>(k);
    }
    count_launch();
    return cuda_status();
}

extern "C" int pmb_gemv(const pmb_gemv_t *d, pmb_stream_t stream) {
    if (!d || !d->vec || !d->B || !d->out || d->nk < 1 || d->nk > 4 || d->nx < 0 || d->nx > 4) return PMB_E_BADARG;
    GemvDev k;
    k.vec = d->vec;
    k.B = d->B;
    k.out = d->out;
    k.nk = d->nk;
    k.nx = d->nx;
    k.alpha = d->alpha;
    k.beta = d->beta;
    k.K = k.X = 1;
    for (int i = 0; i < 4; ++i) {
        const bool uk = i < d->nk, ux = i < d->nx;
        if ((uk && d->k_ext[i] <= 0) || (ux && d->x_ext[i] <= 0)) return PMB_E_BADARG;
        k.k_ext[i] = uk ? d->k_ext[i] : 1;
        k.v_kstr[i] = uk ? d->v_kstr[i] : 0;
        k.b_kstr[i] = uk ? d->b_kstr[i] : 0;
        k.x_ext[i] = ux ? d->x_ext[i] : 1;
        k.b_xstr[i] = ux ? d->b_xstr[i] : 0;
        k.o_xstr[i] = ux ? d->o_xstr[i] : 0;
        k.K *= k.k_ext[i];
        k.X *= k.x_ext[i];
    }
    if (k.X >= (1LL << 31) * 8) return PMB_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    const long long bk0 = k.b_kstr[0] < 0 ? -k.b_kstr[0] : k.b_kstr[0];
    const long long bx0 = k.nx ? (k.b_xstr[0] < 0 ? -k.b_xstr[0] : k.b_xstr[0]) : (1LL << 62);
    if (bk0 < bx0 && k.k_ext[0] >= 16)
        gemv_kfast_kernel<<<(unsigned)((k.X + 7) / 8), 256, 0, s>>>(k);
    else
        gemv_xfast_kernel<<<(unsigned)((k.X + 255) / 256), 256, 0, s>>>(k);
    count_launch();
    return cuda_status();
}

extern "C" int pmb_cdiv_shifted(int64_t n, const double *diag, double zr, double zi, double shift,
                                const double *xr, const double *xi, double *yr, double *yi,
                                pmb_stream_t stream) {
    if (n <= 0 || !diag || !xr || !xi || !yr || !yi) return PMB_E_BADARG;
    cdiv_shifted_kernel<<<grid_for((size_t)n, 256, 16 * kSmCount), 256, 0, (cudaStream_t)stream>>>(
        (size_t)n, diag, zr, zi, shift, xr, xi, yr, yi);
    count_launch();
    return cuda_status();
}
