// Measures DFMA throughput (per SM per clock) and dependent-issue latency on this GPU for a few
// warps/SM and ILP settings.  Development aid for sizing the fused-pass kernel's FP64 budget.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = double(t1 - t0);
}

template <int ILP>
void run(int warps_per_sm, int sms) {
    double* d;
    int threads = warps_per_sm * 32;
    cudaMalloc(&d, sizeof(double) * (size_t)(sms * threads + 1));
    int iters = 20000;
    dfma_kernel<ILP><<<sms, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    dfma_kernel<ILP><<<sms, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc; cudaMemcpy(&cyc, d + sms * threads, sizeof(double), cudaMemcpyDeviceToHost);
    double fma_per_sm = (double)threads * ILP * iters;
    printf("warps/SM=%2d ILP=%2d : %.1f DFMA/clk/SM  (%.2f cycles per warp-DFMA per SMSP)  %.1f TFLOP/s  cycles/iter=%.1f\n",
           warps_per_sm, ILP, fma_per_sm / cyc, cyc / (iters * ILP * (warps_per_sm / 4.0)), 2.0 * fma_per_sm * sms / (ms * 1e-3) / 1e12, cyc / iters);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    run<1>(4, sms); run<4>(4, sms); run<16>(4, sms);
    run<1>(8, sms); run<4>(8, sms); run<16>(8, sms);
    run<4>(16, sms); run<16>(16, sms); run<8>(32, sms); run<4>(64, sms);
    return 0;
}
