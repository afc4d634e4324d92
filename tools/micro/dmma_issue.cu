// Microbenchmark: what does it take to keep the FP64 tensor pipe (DMMA.8x8x4) of one B200
// SM busy?  Sweeps warps per SM and independent accumulators per warp, with the operand
// fragments either register-resident or re-read from shared memory per k sub-step exactly as
// contract_kernel does (MT + NT 8-byte LDS per MT*NT DMMAs).  Not a bench value; it sets the
// ceiling the contraction kernel's consumer loop can be held to.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_issue dmma_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

template <int MT, int NT, bool LDS, bool SYNC>
__global__ void k(double *out, int iters) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 16 * 264; i += blockDim.x) sm[i] = 1e-3 * (i % 7);
    __syncthreads();
    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double af[MT], bf[NT];
#pragma unroll
    for (int i = 0; i < MT; ++i) af[i] = 1.0 + lane + i;
#pragma unroll
    for (int j = 0; j < NT; ++j) bf[j] = 0.5 + lane * j;
    const double *a = sm + (warp & 3) * 32 + (lane >> 2) + (lane & 3) * 132;
    const double *b = sm + 8 * 264 + (warp >> 2) * 64 % 128 + (lane >> 2) + (lane & 3) * 132;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            if (LDS) {
#pragma unroll
                for (int i = 0; i < MT; ++i) af[i] = ((volatile const double *)a)[(ks & 1) * 4 * 132 + i * 8];
#pragma unroll
                for (int j = 0; j < NT; ++j) bf[j] = ((volatile const double *)b)[(ks & 1) * 4 * 132 + (j * 8) % 64];
            }
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma(acc[i][j], af[i], bf[j]);
        }
        if (SYNC) __syncthreads();
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 123.456) out[threadIdx.x] = s;
}

template <int MT, int NT, bool LDS, bool SYNC>
void run(int warps, int ctas_per_sm, const char *tag) {
    int dev_sms;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    double *out;
    cudaMalloc(&out, 4096);
    const int iters = 4000;
    auto kern = k<MT, NT, LDS, SYNC>;
    const int smem = 16 * 264 * 8;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<<<dev_sms * ctas_per_sm, warps * 32, smem>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<<<dev_sms * ctas_per_sm, warps * 32, smem>>>(out, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 256 * MT * NT * 4.0 * iters * warps * ctas_per_sm * dev_sms;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, smem);
    printf("%-28s MT=%d NT=%2d warps/CTA=%2d CTAs/SM=%d (occ %d)  %7.2f TFLOP/s  err=%s\n", tag, MT, NT, warps,
           ctas_per_sm, occ, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    // register-resident fragments: pure issue ceiling
    run<1, 1, false, false>(4, 1, "reg 1 acc");
    run<1, 4, false, false>(4, 1, "reg 4 acc");
    run<2, 4, false, false>(4, 1, "reg 8 acc");
    run<4, 4, false, false>(4, 1, "reg 16 acc");
    run<4, 8, false, false>(4, 1, "reg 32 acc");
    run<4, 8, false, false>(8, 1, "reg 32 acc");
    run<4, 4, false, false>(8, 1, "reg 16 acc");
    run<4, 4, false, false>(16, 1, "reg 16 acc");
    run<2, 4, false, false>(16, 1, "reg 8 acc");
    run<1, 1, false, false>(16, 1, "reg 1 acc");
    run<1, 1, false, false>(32, 1, "reg 1 acc");
    // fragments from shared memory per sub-step, as in contract_kernel
    run<4, 8, true, false>(4, 1, "lds 32 acc");
    run<4, 8, true, false>(8, 1, "lds 32 acc");
    run<4, 8, true, true>(8, 1, "lds 32 acc + barrier/ktile");
    run<4, 4, true, false>(8, 1, "lds 16 acc");
    run<4, 4, true, false>(16, 1, "lds 16 acc");
    run<4, 4, true, true>(16, 1, "lds 16 acc + barrier/ktile");
    run<4, 4, true, true>(4, 3, "lds 16 acc + barrier, 3 CTA");
    run<4, 4, true, true>(4, 4, "lds 16 acc + barrier, 4 CTA");
    run<8, 4, true, false>(8, 1, "lds 32 acc (64x32)");
    return 0;
}
