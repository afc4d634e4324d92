// Shared helpers for the pymes_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pymes_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "pymes_b200 kernels are written for sm_100a (Blackwell B200) only"
#endif

namespace pmb {

extern long long g_launch_count;  // defined in c_api.cu

inline void count_launch(int n = 1) { g_launch_count += n; }

inline int cuda_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

constexpr int kSmCount = 148;          // B200: 2 dies x 74 SMs
constexpr int kReduceBlocks = 8 * kSmCount;   // 2048 resident threads per SM
constexpr int kReduceThreads = 256;
constexpr int kMaxScalars = 8;

// Deterministic two-stage reduction: every block writes its partial sums to
// ws[block][k]; finish_reduce adds them in a fixed order.
// Sum of x over the CTA; the result is valid in thread 0.  `slot` selects one
// of kMaxScalars shared scratch rows so several sums can be in flight.
__device__ inline double block_reduce_sum(double x, int slot) {
    __shared__ double sh[kMaxScalars][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sh[slot][warp] = x;
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        x = lane < nw ? sh[slot][lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    }
    return x;
}

template <int NS>
__device__ inline void block_reduce_store(double (&v)[NS], double *ws) {
    static_assert(NS <= kMaxScalars, "too many scalars");
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const double x = block_reduce_sum(v[k], k);
        if (threadIdx.x == 0) ws[(size_t)blockIdx.x * NS + k] = x;
    }
}

// out[k] (+)= scale[k] * sum_b ws[b][k]; single block, fixed summation order.
__global__ void finish_reduce_kernel(const double *ws, int nblocks, int ns,
                                     double *out, int accumulate);

int finish_reduce(const double *ws, int nblocks, int ns, double *out,
                  int accumulate, cudaStream_t s);

}  // namespace pmb
