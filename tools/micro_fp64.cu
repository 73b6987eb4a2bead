// Microbenchmark: latency and issue rate of dependent / independent DFMA chains on one SM, against warps per scheduler and
// independent chains per thread (ILP), plus the same with an LDS.128 per 8 DFMAs (the mix of the by-point PCG sweep).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_fp64 tools/micro_fp64.cu
// Prints cycles per DFMA warp-instruction per scheduler (1 / throughput) and cycles per dependent step of one warp (latency seen).
#include <cuda_runtime.h>
#include <cstdio>
template <int ILP, bool WITH_LDS>
__global__ void k_chain(double* out, long long* cyc, int n, double a, double b)
{
    __shared__ double2 tab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = make_double2(1.0 + 1e-9 * i, 1e-9 * i);
    __syncthreads();
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = 1.0 + threadIdx.x * 1e-6 + k;
    unsigned idx = threadIdx.x * 7u;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
        }
        if (WITH_LDS) { const double2 t = tab[idx & 1023u]; idx += 33u; a = t.x; b = t.y * 1e-30 + b; }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP, bool WITH_LDS>
void run(int warps, double* out, long long* cyc)
{
    const int n = 2000;
    k_chain<ILP, WITH_LDS><<<148, 32 * warps>>>(out, cyc, n, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    k_chain<ILP, WITH_LDS><<<148, 32 * warps>>>(out, cyc, n, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += (double)h[i]; c /= 148;
    const double steps = 8.0 * n;                                  // dependent steps per chain
    const double instr_per_sched = steps * ILP * warps / 4.0;      // DFMA warp-instructions per scheduler
    printf("lds=%d warps/SM=%2d (%.2f per scheduler) ILP=%d : %7.2f cycles per dependent step, %5.2f cycles per DFMA per scheduler\n",
           (int)WITH_LDS, warps, warps / 4.0, ILP, c / steps, c / instr_per_sched);
}
int main()
{
    double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    const int ws[] = {1, 4, 8, 12, 16, 20, 24, 32};
    for (int w : ws) { run<1, false>(w, out, cyc); run<2, false>(w, out, cyc); run<4, false>(w, out, cyc); run<8, false>(w, out, cyc); }
    for (int w : ws) { run<1, true>(w, out, cyc); run<2, true>(w, out, cyc); run<4, true>(w, out, cyc); }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
    return 0;
}
