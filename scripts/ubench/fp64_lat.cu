// Microbenchmark: fp64 FMA dependent-issue latency and throughput on sm_100a.
// chains = independent DFMA chains per thread, warps = warps per CTA (one CTA per SM).
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, int iters, double a, double b) {
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = threadIdx.x + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int CH>
void run(int warps) {
    double* d; cudaMalloc(&d, 148 * 1024 * 8);
    const int iters = 4096;
    k<CH><<<148, warps * 32>>>(d, iters, 1.0000001, 1e-9);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<CH><<<148, warps * 32>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    double per_iter = cyc / iters;
    double dfma_per_clk_sm = (double)warps * 32 * CH * iters / cyc;
    printf("chains=%2d warps/SM=%2d  cycles/iter=%7.2f  cycles per DFMA per warp=%6.2f  DFMA lanes/clk/SM=%6.1f  (%.3f ms)\n", CH, warps,
           per_iter, per_iter / CH, dfma_per_clk_sm, ms);
    cudaFree(d);
}
int main() {
    for (int w : {1, 4, 8, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); run<16>(w); }
    return 0;
}
