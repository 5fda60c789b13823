// Probe: DFMA dependent-issue latency and pipe interval on this GPU.
// One block per SM; W warps per SM; each thread runs CH independent DFMA chains of length ITER.
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double *out, int iters, double a, double b, long long *cyc)
{
    double x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < CH; ++i) x[i] = fma(x[i], a, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
void run(int warps)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<CH><<<148, warps * 32>>>(out, iters, 0.999999, 1e-9, cyc);
    k<CH><<<148, warps * 32>>>(out, iters, 0.999999, 1e-9, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (iters * 8.0 * CH); // cycles per DFMA per warp
    double wps = warps / 4.0;
    printf("chains %d warps/SM %2d (%.2g per scheduler): %.2f cycles per DFMA per warp, %.2f cycles per DFMA per scheduler\n", CH, warps, wps, per, per / (wps < 1 ? 1 : wps));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int w : {1, 4, 8, 16, 24, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
