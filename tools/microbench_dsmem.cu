// Micro-benchmark: throughput of red.shared::cluster.add.u32 to random (rank, word) of a cluster-distributed table
// (what a 4^9-bin histogram spread over the shared memory of an 8-CTA cluster would do), against local red.shared.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}

template <int CL, bool REMOTE>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(1024) bench(uint32_t* gout, int iters, long long* cycles, uint32_t* info)
{
    extern __shared__ uint32_t sh[];
    constexpr uint32_t WORDS = 32768;
    cg::cluster_group cluster = cg::this_cluster();
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) sh[i] = 0;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sh);
    uint32_t b[CL];
#pragma unroll
    for (int r = 0; r < CL; ++r) b[r] = mapa(base, r);
    if (blockIdx.x == 0 && threadIdx.x == 0) { for (int r = 0; r < CL; ++r) info[r] = b[r]; info[15] = base; }
    const uint32_t stride = CL > 1 ? b[1] - b[0] : 0;
    cluster.sync();
    uint32_t s = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u;
    s ^= s >> 13; s *= 0x9E3779B1u; s ^= s >> 16; s |= 1;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            s ^= s << 13; s ^= s >> 17; s ^= s << 5;
            const uint32_t off = (s >> 7) & ((WORDS - 1) << 2);
            if (REMOTE) {
                const uint32_t rank = (s >> 27) & (CL - 1);
                const uint32_t a = b[0] + rank * stride + off;
                asm volatile("red.shared::cluster.add.u32 [%0], 1;" ::"r"(a) : "memory");
            } else {
                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + off) : "memory");
            }
        }
    }
    long long t1 = clock64();
    cluster.sync();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    uint32_t sum = 0;
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) sum += sh[i];
    atomicAdd(gout, sum);
}

template <int CL, bool REMOTE>
void run(const char* name, int iters, uint32_t* gout, long long* dcyc, uint32_t* info)
{
    size_t smem = 131072;
    CK(cudaFuncSetAttribute(bench<CL, REMOTE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = 144 / CL * CL;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    bench<CL, REMOTE><<<grid, 1024, smem>>>(gout, 2, dcyc, info);
    CK(cudaDeviceSynchronize());
    float best = 1e30f; long long cyc = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(gout, 0, 4));
        CK(cudaEventRecord(e0));
        bench<CL, REMOTE><<<grid, 1024, smem>>>(gout, iters, dcyc, info);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) { best = ms; long long hh[148]; CK(cudaMemcpy(hh, dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost)); cyc = 0; for (int i = 0; i < grid; ++i) cyc = hh[i] > cyc ? hh[i] : cyc; }
    }
    uint32_t total = 0, hinfo[16];
    CK(cudaMemcpy(&total, gout, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hinfo, info, 64, cudaMemcpyDeviceToHost));
    double ops = 1024.0 * iters * 16.0;
    printf("%-34s grid %3d: %8.3f ms  %8.1f Gops/s  %6.3f ops/clk/SM  sum %s  mapa: base %08x r0 %08x r1 %08x stride %08x\n", name, grid, best,
           ops * grid / best * 1e-6, ops / cyc, total == (uint32_t)(ops * grid) ? "ok" : "MISMATCH", hinfo[15], hinfo[0], hinfo[1], hinfo[1] - hinfo[0]);
    fflush(stdout);
}

int main()
{
    uint32_t *gout, *info; long long* dcyc;
    CK(cudaMalloc(&gout, 4096)); CK(cudaMalloc(&info, 4096)); CK(cudaMalloc(&dcyc, sizeof(long long) * 148));
    const int IT = 500;
    run<1, false>("local red.shared", IT, gout, dcyc, info);
    run<2, true>("cluster 2, red.shared::cluster", IT, gout, dcyc, info);
    run<4, true>("cluster 4, red.shared::cluster", IT, gout, dcyc, info);
    run<8, true>("cluster 8, red.shared::cluster", IT, gout, dcyc, info);
    return 0;
}
