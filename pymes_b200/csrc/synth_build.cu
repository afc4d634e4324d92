// Synthetic non-hermitian two-body integrals (BASELINE.json configs[2]: "synthetic FCIDUMP, o = 50,
// v = 500, random permutation-symmetric real integrals"), generated block-wise on the device.
//
//   V[p,q,r,s] = eps * Z[ h(seed, c) >> 48 ],  c = min( ((p n + q) n + r) n + s, ((q n + p) n + s) n + r )
//
// c is the canonical representative of the only symmetry a transcorrelated Hamiltonian keeps,
// (p,q,r,s) ~ (q,p,s,r) (pymes/util/fcidump.py:147-149); h is two rounds of the splitmix64
// finaliser; Z is a 65536-entry table of standard-normal quantiles computed once on the host.
// Integer arithmetic, one table look-up and one multiply: the host generator
// (pymes_b200/util/synthetic.py) and this kernel produce the same doubles bit for bit, so any
// rank can generate its own row block of the 500 GB V_abcd and tests can regenerate any
// sub-block on the host.  HBM-bound: 8 B per stored element, coalesced along s.
#include "common.cuh"

namespace pmb {

__device__ __forceinline__ unsigned long long synth_mix(unsigned long long x) {
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

struct SynthArgs {
    unsigned long long n, seed_term;
    double eps;
    const double *table;
    int lo[4], ext[4];
    double *out;
};

// one thread per (element of s) with the (p,q,r) row index carried by blockIdx / a grid-stride
// loop over rows: consecutive threads write consecutive s
__global__ void __launch_bounds__(256) synth_block_kernel(const __grid_constant__ SynthArgs a) {
    const unsigned long long rows = (unsigned long long)a.ext[0] * a.ext[1] * a.ext[2];
    const int es = a.ext[3];
    const int per_row = (es + blockDim.x - 1) / blockDim.x;          // thread chunks along s
    const unsigned long long work = rows * per_row;
    for (unsigned long long w = blockIdx.x; w < work; w += gridDim.x) {
        const unsigned long long row = w / per_row;
        const int sl = (int)(w - row * per_row) * blockDim.x + threadIdx.x;
        if (sl >= es) continue;
        unsigned long long t = row;
        const unsigned long long r = a.lo[2] + t % a.ext[2];
        t /= a.ext[2];
        const unsigned long long q = a.lo[1] + t % a.ext[1];
        const unsigned long long p = a.lo[0] + t / a.ext[1];
        const unsigned long long s = a.lo[3] + sl;
        const unsigned long long i1 = ((p * a.n + q) * a.n + r) * a.n + s;
        const unsigned long long i2 = ((q * a.n + p) * a.n + s) * a.n + r;
        const unsigned long long c = i1 < i2 ? i1 : i2;
        const unsigned long long h = synth_mix(synth_mix(c * 0x9E3779B97F4A7C15ULL + a.seed_term));
        a.out[row * es + sl] = a.eps * __ldg(a.table + (h >> 48));
    }
}

}  // namespace pmb

using namespace pmb;

extern "C" int pmb_synth_block(int n_orb, unsigned long long seed, double eps, const double *table,
                               const int32_t lo[4], const int32_t ext[4], double *out, pmb_stream_t stream) {
    if (n_orb <= 0 || n_orb > 65535 || !table || !lo || !ext || !out) return PMB_E_BADARG;
    SynthArgs a;
    a.n = (unsigned long long)n_orb;
    a.seed_term = seed * 0xBF58476D1CE4E5B9ULL + 1ULL;
    a.eps = eps;
    a.table = table;
    unsigned long long rows = 1;
    for (int d = 0; d < 4; ++d) {
        if (lo[d] < 0 || ext[d] <= 0 || lo[d] + ext[d] > n_orb) return PMB_E_BADARG;
        a.lo[d] = lo[d];
        a.ext[d] = ext[d];
        if (d < 3) rows *= (unsigned long long)ext[d];
    }
    a.out = out;
    const int threads = ext[3] >= 256 ? 256 : (ext[3] >= 128 ? 128 : (ext[3] >= 64 ? 64 : 32));
    const unsigned long long work = rows * (unsigned long long)((ext[3] + threads - 1) / threads);
    unsigned long long blocks = work < (unsigned long long)(32 * kSmCount) ? work : (unsigned long long)(32 * kSmCount);
    synth_block_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    return cuda_status();
}
