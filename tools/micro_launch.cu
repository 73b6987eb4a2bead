// Microbenchmark: cost of a small kernel between two others, launched (a) normally, (b) cooperatively with two grid.sync(), (c) normally
// with a hand-rolled grid barrier - in a stream and replayed from a CUDA graph.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=false
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
namespace cg = cooperative_groups;
__global__ void k_plain(double* p) { if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.0; }
__global__ void k_coop(double* p) { cg::grid_group g = cg::this_grid(); g.sync(); if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.0; g.sync(); }
__device__ __forceinline__ void bar(unsigned* c, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(c, 1u); unsigned v; do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory"); } while (v < target); }
    __syncthreads();
}
__global__ void k_hand(double* p, unsigned* c, unsigned* epoch) { const unsigned e = *epoch, nb = gridDim.x; bar(c, (2 * e + 1) * nb); if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.0; bar(c, (2 * e + 2) * nb); if (threadIdx.x == 0 && blockIdx.x == 0) *epoch = e + 1; }
__global__ void k_big(double* p, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = p[i] * 1.0000001 + 1.0; }
int main()
{
    double* d; cudaMalloc(&d, 8 << 20); cudaMemset(d, 0, 8 << 20);
    unsigned* c; cudaMalloc(&c, 64); cudaMemset(c, 0, 64);
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int NB = 56, NT = 192, REP = 200;
    for (int mode = 0; mode < 4; ++mode) {
        for (int graph = 0; graph < 2; ++graph) {
            auto seq = [&]() {
                for (int r = 0; r < REP; ++r) {
                    k_big<<<148 * 4, 256, 0, s>>>(d + 1024, 148 * 1024);
                    if (mode == 1) k_plain<<<NB, NT, 0, s>>>(d);
                    if (mode == 2) { void* a[] = {(void*)&d}; cudaLaunchCooperativeKernel((const void*)k_coop, dim3(NB), dim3(NT), a, 0, s); }
                    if (mode == 3) { unsigned* ep = c + 8; k_hand<<<NB, NT, 0, s>>>(d, c, ep); }
                }
            };
            float ms = 0;
            if (!graph) { seq(); cudaStreamSynchronize(s); cudaEventRecord(e0, s); seq(); cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
            else {
                cudaGraph_t g; cudaGraphExec_t ge;
                cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal); seq(); cudaStreamEndCapture(s, &g);
                if (cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) { printf("graph instantiate failed mode %d\n", mode); continue; }
                cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
                cudaEventRecord(e0, s); cudaGraphLaunch(ge, s); cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            }
            const char* names[] = {"big kernel alone", "+ plain small kernel", "+ cooperative kernel, 2 grid.sync", "+ plain kernel, 2 hand-rolled barriers"};
            printf("%-42s %s: %.2f us per pair (%s)\n", names[mode], graph ? "graph " : "stream", ms * 1e3 / REP, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
