// fp64_peak.cu — DFMA microbenchmark: the measured FP64 vector peak of this GPU (the
// denominator of the FP64-pipe fraction reported for the narrow-phase and coupling kernels; B200's FP64 rate
// is not in MEASURED_PEAKS.json).  Prints one JSON line.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o fp64_peak fp64_peak.cu (done by __graft_entry__.build()).
#include <cuda_runtime.h>
#include <stdio.h>

#define CHAINS 8
template <bool FMA>
__global__ void __launch_bounds__(256) k_fp64(double *out, int iters, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = (double)(threadIdx.x + c) * 1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (FMA) x[c] = fma(x[c], a, b);
            else x[c] = __dadd_rn(__dmul_rn(x[c], a), b);  // what -fmad=false code issues: DMUL + DADD
        }
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 12345.6789) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true: keeps the chains alive
}

template <bool FMA>
static double run(int sms, int iters) {
    double *out;
    cudaMalloc(&out, sizeof(double) * 256 * sms * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sms * 8;
    k_fp64<FMA><<<blocks, 256>>>(out, iters / 10, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_fp64<FMA><<<blocks, 256>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(out);
    double flops = 2.0 * CHAINS * (double)iters * 256.0 * blocks;
    return flops / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) {
        printf("{\"error\": \"no CUDA device\"}\n");
        return 1;
    }
    double fma = run<true>(p.multiProcessorCount, 20000), nofma = run<false>(p.multiProcessorCount, 20000);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.3f, \"dmul_dadd_tflops\": %.3f, "
           "\"how\": \"8 independent chains per thread, 256 threads x 8 blocks per SM, best of 5, CUDA events\"}\n",
           p.name, p.multiProcessorCount, fma, nofma);
    return 0;
}
