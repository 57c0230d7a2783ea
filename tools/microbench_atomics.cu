// Micro-benchmark: shared-memory / L2 atomic increment throughput on B200 (sm_100a).
// Design input for the k-mer counting kernel (DESIGN.md "atomic rate" section).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench_atomics tools/microbench_atomics.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t xs(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

enum Mode { RANDOM = 0, BANKFREE = 1, SAMEADDR = 2, RET = 3, U64 = 4, MANUAL = 5, MATCH = 6, GLOBAL = 7, NOATOM = 8, POLY = 9, PACK16 = 10, MANUAL_U8 = 11 };

template <int MODE>
__global__ void __launch_bounds__(1024) bench(uint32_t* __restrict__ gout, uint32_t* __restrict__ ghist, int iters, uint32_t nbins_mask, long long* cycles, uint32_t gmask)
{
    extern __shared__ uint32_t sh[];
    const int nb = nbins_mask + 1;
    for (int i = threadIdx.x; i < nb * (MODE == U64 ? 2 : 1); i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint32_t s = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u;
    uint32_t acc = 0;
    const uint32_t lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint32_t r = xs(s);
            uint32_t b = r & nbins_mask;
            if (MODE == RANDOM) { atomicAdd(&sh[b], 1u); }
            else if (MODE == BANKFREE) { b = ((b & ~31u) | lane); atomicAdd(&sh[b], 1u); }
            else if (MODE == SAMEADDR) { b = (r >> 20) & nbins_mask & ~0u; b = __shfl_sync(0xffffffffu, b, 0); atomicAdd(&sh[b], 1u); }
            else if (MODE == RET) { acc += atomicAdd(&sh[b], 1u); }
            else if (MODE == U64) { atomicAdd(reinterpret_cast<unsigned long long*>(sh) + b, 1ull); }
            else if (MODE == MANUAL) { uint32_t v = sh[b]; sh[b] = v + 1; }
            else if (MODE == MANUAL_U8) { uint8_t* p = reinterpret_cast<uint8_t*>(sh) + (r & (nbins_mask * 4 + 3)); *p = *p + 1; }
            else if (MODE == MATCH) { uint32_t m = __match_any_sync(0xffffffffu, b); if ((m & ((1u << lane) - 1)) == 0) atomicAdd(&sh[b], __popc(m)); }
            else if (MODE == GLOBAL) { atomicAdd(&ghist[r & gmask], 1u); }
            else if (MODE == NOATOM) { acc += b; }
            else if (MODE == POLY) { if (lane < 4) b = 0x3fff & nbins_mask; atomicAdd(&sh[b], 1u); }
            else if (MODE == PACK16) { atomicAdd(&sh[b >> 1], (b & 1) ? 0x10000u : 1u); }
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    uint32_t sum = acc;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) sum += sh[i];
    if (sum == 0xdeadbeef) gout[0] = sum;
}

template <int MODE>
void run(const char* name, int threads, int ctas_per_sm, uint32_t nbins, int iters, uint32_t* gout, uint32_t* ghist, long long* dcyc, uint32_t gbins = 1)
{
    int nsm = 148;
    size_t smem = (size_t)nbins * 4 * (MODE == U64 ? 2 : 1);
    CK(cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bench<MODE>, threads, smem));
    if (occ < ctas_per_sm) { printf("%s,threads=%d,ctas=%d,bins=%u: occupancy only %d, skipped\n", name, threads, ctas_per_sm, nbins, occ); return; }
    int grid = nsm * ctas_per_sm;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    bench<MODE><<<grid, threads, smem>>>(gout, ghist, 4, nbins - 1, dcyc, gbins - 1);
    CK(cudaDeviceSynchronize());
    float best = 1e30f; long long cyc = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        bench<MODE><<<grid, threads, smem>>>(gout, ghist, iters, nbins - 1, dcyc, gbins - 1);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) { best = ms; long long h[148 * 4]; CK(cudaMemcpy(h, dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost)); cyc = 0; for (int i = 0; i < grid; ++i) cyc = h[i] > cyc ? h[i] : cyc; }
    }
    double ops = (double)grid * threads * iters * 8.0;
    double per_clk_sm = (double)threads * ctas_per_sm * iters * 8.0 / (double)cyc;
    printf("%-10s threads=%4d ctas/sm=%d bins=%6u gbins=%7u : %8.3f ms  %8.1f Gops/s  %6.3f ops/clk/SM (max-CTA cycles %lld, eff clk %.0f MHz)\n",
           name, threads, ctas_per_sm, nbins, gbins, best, ops / best * 1e-6, per_clk_sm, cyc, cyc / (best * 1e3));
    fflush(stdout);
}

int main()
{
    uint32_t *gout, *ghist; long long* dcyc;
    CK(cudaMalloc(&gout, 4096));
    CK(cudaMalloc(&ghist, 4u << 20));
    CK(cudaMemset(ghist, 0, 4u << 20));
    CK(cudaMalloc(&dcyc, sizeof(long long) * 148 * 4));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s SMs %d smem/SM %zu smem/block optin %zu clock %d kHz\n", p.name, p.multiProcessorCount, p.sharedMemPerMultiprocessor, p.sharedMemPerBlockOptin, p.clockRate);
    const int IT = 2000;
    run<NOATOM>("noatom", 1024, 1, 16384, IT, gout, ghist, dcyc);
    for (int th : {256, 512, 1024}) run<RANDOM>("random", th, 1, 16384, IT, gout, ghist, dcyc);
    run<RANDOM>("random", 1024, 2, 16384, IT, gout, ghist, dcyc);
    run<RANDOM>("random", 640, 3, 16384, IT, gout, ghist, dcyc);
    run<RANDOM>("random", 1024, 2, 4096, IT, gout, ghist, dcyc);
    run<RANDOM>("random", 1024, 2, 1024, IT, gout, ghist, dcyc);
    run<BANKFREE>("bankfree", 1024, 1, 16384, IT, gout, ghist, dcyc);
    run<BANKFREE>("bankfree", 1024, 2, 16384, IT, gout, ghist, dcyc);
    run<SAMEADDR>("sameaddr", 1024, 2, 16384, IT / 4, gout, ghist, dcyc);
    run<POLY>("poly4", 1024, 2, 16384, IT, gout, ghist, dcyc);
    run<RET>("ret", 1024, 2, 16384, IT, gout, ghist, dcyc);
    run<U64>("u64", 1024, 1, 8192, IT, gout, ghist, dcyc);
    run<U64>("u64", 1024, 1, 16384, IT, gout, ghist, dcyc);
    run<PACK16>("pack16", 1024, 2, 16384, IT, gout, ghist, dcyc);
    run<MANUAL>("manual", 1024, 2, 16384, IT, gout, ghist, dcyc);
    run<MANUAL_U8>("manual_u8", 1024, 2, 16384, IT, gout, ghist, dcyc);
    run<MATCH>("match", 1024, 2, 16384, IT / 4, gout, ghist, dcyc);
    run<MATCH>("match", 1024, 2, 1024, IT / 4, gout, ghist, dcyc);
    for (uint32_t nb : {16384u, 65536u, 262144u, 1048576u}) run<GLOBAL>("global", 1024, 2, 1024, IT / 8, gout, ghist, dcyc, nb);
    return 0;
}
