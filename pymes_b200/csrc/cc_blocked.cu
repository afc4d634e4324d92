// Block-diagonal contraction on the FP64 tensor cores: pmb_blocked_contract.
//
// A momentum-conserving integral block V[p,q,r,s] (pymes/model/ueg.py:411-513: non-zero only
// where k_p + k_q = k_r + k_s) is, as the matrix [(p,q), (r,s)] that a contraction such as the
// particle-particle ladder "abcd,cdij->abij" (ccd.py:187) multiplies with, block diagonal once
// its rows and columns are grouped by total momentum Q: rows (p,q) with k_p + k_q = Q meet only
// the entries (r,s) with k_r + k_s = Q.  At 54 electrons / 515 plane waves the 5.7e10 elements of
// V_abcd hold 3.3e7 non-zeros in 3809 such groups, so the ladder is 4.8e10 flop
// instead of 8.3e13.  The host sorts the row and entry lists by group and cuts each group's rows
// into tiles of <= 64; one CTA owns (row tile, 128 columns), walks the group's entries 16 at a
// time through a two-stage cp.async pipeline and multiplies with DMMA.m8n8k4.  Every address is
// list-driven (row / entry offset tables), so the same kernel serves the compressed integrals
// (pmb_ueg_build_nz: A[a_moff[m] + a_koff[k]] with a_koff = r) and a dense strided block.
//
// Each C element is written by exactly one CTA in a fixed summation order: deterministic, no
// atomics, no workspace.
#include "common.cuh"

namespace pmb {
namespace {

constexpr int BM = 64, BN = 128, BK = 16, NT = 256, SPAD = 4;
constexpr int LDA = BM + SPAD, LDB = BN + SPAD;   // (LD mod 16) == 4: conflict-free fragment loads
constexpr int WARPS_M = 2, WARPS_N = 4;
constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
constexpr int MT = WM / 8, NTL = WN / 8;
constexpr int KCH = 512;                           // entry offsets staged in shared memory at a time
constexpr int PER_A = BM * BK / NT, PER_B = BN * BK / NT;
static_assert(NT % BK == 0 && NT % BN == 0, "thread -> element mappings");
static_assert(WARPS_M * WARPS_N * 32 == NT, "warp grid");

constexpr size_t kSmemBytes = (size_t)2 * BK * (LDA + LDB) * 8 + (size_t)2 * KCH * 8 + (size_t)2 * BM * 8;

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c[0]), "+d"(c[1])
        : "d"(a), "d"(b));
}
// src_bytes = 0 zero-fills the destination (ragged edges need no branches)
__device__ __forceinline__ void cp_async8z(unsigned smem_addr, const double *gsrc, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_addr), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(NT, 2) blocked_kernel(const __grid_constant__ pmb_blocked_t p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *As = reinterpret_cast<double *>(smem_raw);          // [2][BK][LDA]
    double *Bs = As + 2 * BK * LDA;                              // [2][BK][LDB]
    long long *s_ak = reinterpret_cast<long long *>(Bs + 2 * BK * LDB);   // [KCH]
    long long *s_bk = s_ak + KCH;                                // [KCH]
    long long *s_am = s_bk + KCH;                                // [BM]
    long long *s_cm = s_am + BM;                                 // [BM]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_m = warp % WARPS_M, warp_n = warp / WARPS_M;
    const int4 tile = *reinterpret_cast<const int4 *>(p.tiles + 4 * (size_t)blockIdx.x);
    const int m0 = tile.x, mn = tile.y, k0 = tile.z, kn = tile.w;
    const int n_base = blockIdx.y * BN;
    const int nrem = p.n0_ext * p.n1_ext - n_base;

    for (int i = tid; i < BM; i += NT) {
        const bool ok = i < mn;
        s_am[i] = ok ? p.a_moff[m0 + i] : 0;
        s_cm[i] = ok ? p.c_moff[m0 + i] : 0;
    }
    __syncthreads();
    // B tile: a thread keeps its column and steps the entry
    const int b_n = tid % BN, b_k0 = tid / BN;
    constexpr int B_KSTEP = NT / BN;
    const bool b_nok = b_n < nrem;
    long long b_noff = 0;
    if (b_nok) {
        const int n = n_base + b_n, q = n / p.n0_ext;
        b_noff = (long long)q * p.b_n1str + (n - q * p.n0_ext);
    }
    // A tile: a thread keeps its entry and steps the row
    const int a_k = tid % BK, a_m0 = tid / BK;
    constexpr int A_MSTEP = NT / BK;
    const unsigned as_base = (unsigned)__cvta_generic_to_shared(As);
    const unsigned bs_base = (unsigned)__cvta_generic_to_shared(Bs);

    double acc[MT][NTL][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTL; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kc = 0; kc < kn; kc += KCH) {
        const int kcn = min(KCH, kn - kc);
        __syncthreads();           // the previous chunk's tables are no longer read
        for (int i = tid; i < kcn; i += NT) {
            s_ak[i] = p.a_koff[k0 + kc + i];
            s_bk[i] = p.b_koff[k0 + kc + i];
        }
        __syncthreads();
        const int nkt = (kcn + BK - 1) / BK;
        auto issue = [&](int kt, int st) {
            const int kb = kt * BK;
            {
                const int k = kb + a_k;
                const bool kok = k < kcn;
                const long long ko = kok ? s_ak[k] : 0;
                const unsigned dst = as_base + (unsigned)(((st * BK + a_k) * LDA + a_m0) * 8);
#pragma unroll
                for (int it = 0; it < PER_A; ++it) {
                    const int m = a_m0 + it * A_MSTEP;
                    const bool ok = kok && m < mn;
                    cp_async8z(dst + (unsigned)(it * A_MSTEP * 8), p.A + (ok ? s_am[m] + ko : 0), ok ? 8 : 0);
                }
            }
            {
                const unsigned dst = bs_base + (unsigned)(((st * BK + b_k0) * LDB + b_n) * 8);
#pragma unroll
                for (int it = 0; it < PER_B; ++it) {
                    const int k = kb + b_k0 + it * B_KSTEP;
                    const bool ok = b_nok && k < kcn;
                    cp_async8z(dst + (unsigned)(it * B_KSTEP * LDB * 8), p.B + (ok ? s_bk[k] + b_noff : 0),
                               ok ? 8 : 0);
                }
            }
            cp_async_commit();
        };
        issue(0, 0);
        for (int kt = 0; kt < nkt; ++kt) {
            const int st = kt & 1;
            if (kt + 1 < nkt) {
                issue(kt + 1, st ^ 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();       // everyone's copies of tile kt have landed
            const double *a = As + st * BK * LDA + warp_m * WM + (lane >> 2) + (lane & 3) * LDA;
            const double *b = Bs + st * BK * LDB + warp_n * WN + (lane >> 2) + (lane & 3) * LDB;
#pragma unroll
            for (int ks = 0; ks < BK / 4; ++ks) {
                double af[MT], bf[NTL];
#pragma unroll
                for (int i = 0; i < MT; ++i) af[i] = a[ks * 4 * LDA + i * 8];
#pragma unroll
                for (int j = 0; j < NTL; ++j) bf[j] = b[ks * 4 * LDB + j * 8];
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NTL; ++j) dmma(acc[i][j], af[i], bf[j]);
            }
            __syncthreads();       // stage st may be overwritten by the next iteration's copies
        }
    }

    // ---- epilogue: C = alpha * acc + beta * C, one owner per element ----------------------
    const int row0 = warp_m * WM + (lane >> 2), col0 = warp_n * WN + (lane & 3) * 2;
    long long c_noff[NTL][2];
#pragma unroll
    for (int j = 0; j < NTL; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int nl = col0 + j * 8 + c;
            long long off = 0;
            if (nl < nrem) {
                const int n = n_base + nl, q = n / p.n0_ext;
                off = (long long)q * p.c_n1str + (n - q * p.n0_ext);
            }
            c_noff[j][c] = off;
        }
    const bool rd = p.beta != 0.0;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int ml = row0 + i * 8;
        if (ml >= mn) continue;
        double *crow = p.C + s_cm[ml];
        double old[NTL][2];
#pragma unroll
        for (int j = 0; j < NTL; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c)
                old[j][c] = (rd && col0 + j * 8 + c < nrem) ? crow[c_noff[j][c]] : 0.0;
#pragma unroll
        for (int j = 0; j < NTL; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c)
                if (col0 + j * 8 + c < nrem) crow[c_noff[j][c]] = p.alpha * acc[i][j][c] + p.beta * old[j][c];
    }
}

}  // namespace
}  // namespace pmb

using namespace pmb;

extern "C" int pmb_blocked_contract(const pmb_blocked_t *d, pmb_stream_t stream) {
    if (!d || d->n_tiles < 0 || d->n0_ext < 1 || d->n1_ext < 1) return PMB_E_BADARG;
    if (d->n_tiles == 0) return 0;
    if (!d->A || !d->B || !d->C || !d->a_moff || !d->c_moff || !d->a_koff || !d->b_koff || !d->tiles)
        return PMB_E_BADARG;
    const long long n = (long long)d->n0_ext * d->n1_ext;
    if (n > 0x7fffffffLL) return PMB_E_UNSUPPORTED;
    const long long tiles_n = (n + BN - 1) / BN;
    if (tiles_n > 65535) return PMB_E_UNSUPPORTED;
    // per device, so not cached: a host-side call of a microsecond
    cudaError_t e = cudaFuncSetAttribute(blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)d->n_tiles, (unsigned)tiles_n);
    blocked_kernel<<<grid, NT, kSmemBytes, (cudaStream_t)stream>>>(*d);
    count_launch();
    return cuda_status();
}

// ---------------------------------------------------------------------------
// pmb_gather_expand: a momentum-conserving integral block times a ONE-index contraction.
//
// out[x0,x1,x2,j] (+)= alpha * sum_y V[x0,x1,x2 | y] * D[y, j] has, for a UEG block, exactly one
// candidate y per (x0,x1,x2) -- the orbital the reference's index lookup finds (ueg.py:395-404).  With
// the values val[x] = V[x, y*(x)] and partners idx[x] = y*(x) (or -1) tabulated once per block and
// summed axis, the product is an HBM-bound pass over the OUTPUT: 8 B written (16 with beta != 0)
// per element, 12 B of table per (x0,x1,x2), instead of a pass over the o.v^3 block (25 GB at
// v = 488).  These are the T1 dressing products "abid,dj->abij", "abcj,ci->abij", "iabc,cj->iabj",
// "iacb,cj->iajb" of ccsd.py:322-419.  One thread per output element in output memory order (the
// output is C-contiguous); `role` says which output dimension is x0 / x1 / x2 / j.
// ---------------------------------------------------------------------------
namespace pmb {
namespace {

// I = unsigned when the output has fewer than 2^31 elements (32-bit index arithmetic), else long long
template <typename I>
__global__ void __launch_bounds__(256) gather_expand_kernel(const __grid_constant__ pmb_gather_t p) {
    const I e3 = (I)p.ext[3], e2 = (I)p.ext[2], e1 = (I)p.ext[1];
    const long long total = (long long)p.ext[0] * p.ext[1] * p.ext[2] * p.ext[3];
    const bool rd = p.beta != 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        I c[4];
        I q = (I)e;
        c[3] = q % e3; q /= e3;
        c[2] = q % e2; q /= e2;
        c[1] = q % e1;
        c[0] = q / e1;
        long long xoff = 0, j = 0;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            if (p.role[d] == 3) j = (long long)c[d];
            else xoff += (long long)c[d] * p.x_str[p.role[d]];
        }
        const int t = __ldg(p.idx + xoff);
        double v = 0.0;
        if (t >= 0) v = p.alpha * __ldg(p.val + xoff) * __ldg(p.D + (long long)t * p.d_ystr + j * p.d_jstr);
        if (rd) v += p.beta * p.out[e];
        p.out[e] = v;
    }
}

}  // namespace
}  // namespace pmb

extern "C" int pmb_gather_expand(const pmb_gather_t *d, pmb_stream_t stream) {
    if (!d || !d->val || !d->idx || !d->D || !d->out) return PMB_E_BADARG;
    long long total = 1;
    int seen = 0;
    for (int k = 0; k < 4; ++k) {
        if (d->ext[k] < 1 || d->role[k] < 0 || d->role[k] > 3) return PMB_E_BADARG;
        seen |= 1 << d->role[k];
        total *= d->ext[k];
    }
    if (seen != 15) return PMB_E_BADARG;
    long long blocks = (total + 255) / 256;
    if (blocks > 32LL * kSmCount) blocks = 32LL * kSmCount;
    if (total < (1LL << 31))
        gather_expand_kernel<unsigned><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*d);
    else
        gather_expand_kernel<long long><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*d);
    count_launch();
    return cuda_status();
}
