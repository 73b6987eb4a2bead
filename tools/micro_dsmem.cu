// Microbenchmark: random 16-byte reads of a table held in (distributed) shared memory, as the by-point sweeps do them.
//   mode 0: table replicated in every CTA's own shared memory (ld.shared)                      -- today's Venice plan
//   mode 1: table sharded over the CTAs of a cluster, read with ld.shared::cluster (mapa)      -- candidate for big camera tables
// Every lane reads FIELDS consecutive 16-byte units of a random record (record = 7 units = 112 bytes), like tab_row_fwd.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/micro/micro_dsmem tools/micro_dsmem.cu ; run: micro_dsmem <cluster>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;
constexpr int UNITS = 7, THREADS = 640, ITERS = 2000;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int CL> __global__ void __launch_bounds__(THREADS, 1) k(const int* __restrict__ idx, int n_idx, int recs_per_cta, double* out, int fields)
{
    extern __shared__ __align__(128) unsigned char smem[];
    double2* tab = reinterpret_cast<double2*>(smem);
    for (int i = threadIdx.x; i < recs_per_cta * UNITS; i += THREADS) tab[i] = make_double2(i * 1e-6, 1.0);
    cg::cluster_group cl = cg::this_cluster();
    if (CL > 1) cl.sync(); else __syncthreads();
    const uint32_t base = smem_u32(smem);
    double acc = 0;
    int k0 = (blockIdx.x * THREADS + threadIdx.x) % n_idx;
    for (int it = 0; it < ITERS; ++it) {
        const int c = __ldg(idx + k0); k0 += 32 * 7; if (k0 >= n_idx) k0 -= n_idx;
        uint32_t addr;
        if (CL > 1) {
            const uint32_t rank = (uint32_t)c / (uint32_t)recs_per_cta, local = (uint32_t)c - rank * recs_per_cta;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(base + local * (UNITS * 16)), "r"(rank));
        } else addr = base + (uint32_t)(c % recs_per_cta) * (UNITS * 16);
#pragma unroll 7
        for (int f = 0; f < fields; ++f) {
            double a, b;
            if (CL > 1) asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr + 16u * f));
            else asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr + 16u * f));
            acc += a * b;
        }
    }
    if (CL > 1) cl.sync();
    if (acc == 123.456) out[0] = acc;
}
template <int CL> float run(const int* d_idx, int n_idx, int n_rec, int fields, double* d_out, int sms)
{
    const int recs_per_cta = CL > 1 ? (n_rec + CL - 1) / CL : n_rec;
    const size_t smem = (size_t)recs_per_cta * UNITS * 16;
    cudaFuncSetAttribute(k<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (CL > 8) cudaFuncSetAttribute(k<CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute at[1];
    cfg.gridDim = dim3((sms / CL) * CL); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem;
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) cudaLaunchKernelEx(&cfg, k<CL>, d_idx, n_idx, recs_per_cta, d_out, fields);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) cudaLaunchKernelEx(&cfg, k<CL>, d_idx, n_idx, recs_per_cta, d_out, fields);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("  cluster %d: %s\n", CL, cudaGetErrorString(e)); return -1; }
    const double reads = (double)cfg.gridDim.x * THREADS * ITERS * 5;      // record reads (each `fields` x 16 bytes)
    printf("  cluster %2d, %5d records/CTA (%6.1f KB), %d fields: %.3f ms for 5 launches -> %.2f G record reads/s, %.1f cycles per warp-LDS per SM @1.965GHz, grid %d\n",
           CL, recs_per_cta, smem / 1024.0, fields, ms, reads / (ms * 1e-3) / 1e9,
           (ms * 1e-3 / 5) * 1.965e9 / ((double)THREADS / 32 * ITERS * fields), cfg.gridDim.x);
    return ms;
}
int main()
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int n_idx = 1 << 22;
    std::vector<int> h(n_idx);
    double* d_out; cudaMalloc(&d_out, 64);
    int* d_idx; cudaMalloc(&d_idx, sizeof(int) * n_idx);
    for (int n_rec : {1778, 13682}) {
        srand(1); for (auto& v : h) v = rand() % n_rec;
        cudaMemcpy(d_idx, h.data(), sizeof(int) * n_idx, cudaMemcpyHostToDevice);
        printf("table of %d records (%.0f KB)\n", n_rec, n_rec * 112 / 1024.0);
        for (int fields : {7}) {
            if (n_rec * 112 <= 227 * 1024) run<1>(d_idx, n_idx, n_rec, fields, d_out, sms);
            if (n_rec * 112 / 2 <= 227 * 1024) run<2>(d_idx, n_idx, n_rec, fields, d_out, sms);
            if (n_rec * 112 / 4 <= 227 * 1024) run<4>(d_idx, n_idx, n_rec, fields, d_out, sms);
            run<8>(d_idx, n_idx, n_rec, fields, d_out, sms);
            run<16>(d_idx, n_idx, n_rec, fields, d_out, sms);
        }
    }
    return 0;
}
