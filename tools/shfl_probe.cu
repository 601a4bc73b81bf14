// Measures the cost of exchanging a thread's 16 doubles with its lane partner three ways (development aid):
// (a) 32 SHFL.BFLY, (b) half exchange: 16 SHFL + selects, (c) through shared memory (STS.128 + LDS.128),
// each combined with the 32 FP64 ops of a real 2x2 update, 16 warps per SM like the fused-pass kernel.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(double* out, int iters, double c0, double c1, int lm) {
    extern __shared__ double2 sm[];
    double xr[8], xi[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { xr[k] = threadIdx.x + k; xi[k] = threadIdx.x - k; }
    const bool b = threadIdx.x & lm;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double pr = __shfl_xor_sync(0xffffffffu, xr[k], lm), pi = __shfl_xor_sync(0xffffffffu, xi[k], lm);
                xr[k] = c0 * xr[k] + c1 * pr; xi[k] = c0 * xi[k] + c1 * pi;
            }
        } else if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double sr = b ? xr[k] : xr[k + 4], si = b ? xi[k] : xi[k + 4];
                const double rr = __shfl_xor_sync(0xffffffffu, sr, lm), ri = __shfl_xor_sync(0xffffffffu, si, lm);
                const double ar = b ? rr : xr[k], ai = b ? ri : xi[k];
                const double br = b ? xr[k + 4] : rr, bi = b ? xi[k + 4] : ri;
                xr[k] = c0 * ar + c1 * br; xi[k] = c0 * ai + c1 * bi;
                xr[k + 4] = c1 * ar - c0 * br; xi[k + 4] = c1 * ai - c0 * bi;
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int k = 0; k < 8; ++k) sm[k * 512 + threadIdx.x] = make_double2(xr[k], xi[k]);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double2 p = sm[k * 512 + (threadIdx.x ^ lm)];
                xr[k] = c0 * xr[k] + c1 * p.x; xi[k] = c0 * xi[k] + c1 * p.y;
            }
            __syncwarp();
        } else {   // FP64 only
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double ar = xr[k], ai = xi[k], br = xr[k + 4], bi = xi[k + 4];
                xr[k] = c0 * ar + c1 * br; xi[k] = c0 * ai + c1 * bi;
                xr[k + 4] = c1 * ar - c0 * br; xi[k + 4] = c1 * ai - c0 * bi;
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += xr[k] + xi[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = double(t1 - t0);
}

template <int MODE>
void run(const char* name, int sms) {
    double* d;
    cudaMalloc(&d, sizeof(double) * (size_t)(sms * 512 + 1));
    const int iters = 4000;
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int rep = 0; rep < 2; ++rep) probe<MODE><<<sms, 512, 65536>>>(d, iters, 0.70710678, 0.70710677, 4);
    cudaDeviceSynchronize();
    double cyc; cudaMemcpy(&cyc, d + sms * 512, sizeof(double), cudaMemcpyDeviceToHost);
    printf("%-28s %.0f cycles per op-tile (16 warps x 8 slots)\n", name, cyc / iters);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s\n", p.name);
    run<3>("fp64 only (32 per thread)", p.multiProcessorCount);
    run<0>("32 SHFL + fp64", p.multiProcessorCount);
    run<1>("16 SHFL + selects + fp64", p.multiProcessorCount);
    run<2>("STS.128/LDS.128 + fp64", p.multiProcessorCount);
    return 0;
}
