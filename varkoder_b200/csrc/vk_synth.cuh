// vk_synth.cuh -- deterministic synthetic FASTQ written straight into device memory (bench / tests only).
//
// Shape from SURVEY.md section 8d: record = "@S%010d\n" + L bases + "\n+\n" + L qualities + "\n" (2L + 17 bytes);
// bases i.i.d. A,T = 0.30 / C,G = 0.20, each replaced by N with p ~ 0.001; 0.5 % of reads end in a poly-G tail
// of 20..60; qualities uniform over '#'..'I' (so quality lines may start with '@' or '+').
// Every byte is a pure function of (seed, read index, offset): varkoder_b200/synth.py produces the identical
// bytes with numpy, which is what lets the CPU oracle check the GPU path on the same reads.
#pragma once
#include "vk_common.cuh"

namespace vk {

__host__ __device__ __forceinline__ uint64_t synth_hash(uint64_t seed, uint64_t r, uint64_t slot)
{
    uint64_t z = seed + r * 0x9E3779B97F4A7C15ull + slot * 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ uint8_t synth_base(uint64_t seed, uint64_t r, uint32_t i, uint32_t len)
{
    const uint64_t hr = synth_hash(seed, r, 0);
    uint32_t tail = 0;
    if ((hr & 0xFFFFu) < 328u) tail = 20u + (uint32_t)((hr >> 16) % 41u);      // 0.5 % of reads
    if (i + tail >= len) return 'G';
    const uint64_t h = synth_hash(seed, r, (uint64_t)i + 1);
    if (((h >> 16) & 0xFFFFFu) < 1049u) return 'N';                               // ~0.001
    const uint32_t u = (uint32_t)(h & 0xFFFFu);
    return u < 19661u ? 'A' : (u < 32768u ? 'C' : (u < 45875u ? 'G' : 'T'));
}

__host__ __device__ __forceinline__ uint8_t synth_qual(uint64_t seed, uint64_t r, uint32_t i)
{
    const uint64_t h = synth_hash(seed, r, (uint64_t)i + 1);
    return (uint8_t)(35u + (uint32_t)((h >> 40) % 39u));
}

// fixed read length L; the last read is truncated so that exactly n_bases bases are written
__global__ void __launch_bounds__(256)
synth_fixed_kernel(uint8_t* __restrict__ out, uint64_t n_out, uint64_t n_reads, uint32_t L, uint32_t last_len,
                   uint64_t seed, uint64_t first_read)
{
    pdl_wait();
    const uint64_t rs = 2ull * L + 17ull;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_out; g += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t rl = g / rs;
        uint32_t o = (uint32_t)(g - rl * rs);
        if (rl >= n_reads) { rl = n_reads - 1; o = (uint32_t)(g - rl * rs); }
        const uint32_t len = (rl == n_reads - 1) ? last_len : L;
        const uint64_t r = first_read + rl;
        uint8_t c;
        if (o == 0) c = '@';
        else if (o == 1) c = 'S';
        else if (o < 12) {
            uint64_t v = r % 10000000000ull;
            for (uint32_t d = 11; d > o; --d) v /= 10;
            c = (uint8_t)('0' + v % 10);
        } else if (o == 12) c = '\n';
        else if (o < 13 + len) c = synth_base(seed, r, o - 13, len);
        else if (o == 13 + len) c = '\n';
        else if (o == 14 + len) c = '+';
        else if (o == 15 + len) c = '\n';
        else if (o < 16 + 2 * len) c = synth_qual(seed, r, o - 16 - len);
        else c = '\n';
        out[g] = c;
    }
}

// ---- variable read lengths ("Bembidion-shaped"): lengths uniform min_len..max_len, a
// fraction short_per_10000 / 10000 of the reads shorter than k (0..k-1 bases, empty ones included).  Same bytes as
// varkoder_b200/synth.py variable().
__host__ __device__ __forceinline__ uint32_t synth_var_len(uint64_t seed, uint64_t r, uint32_t min_len, uint32_t max_len,
                                                           uint32_t short_per_10000, uint32_t k)
{
    const uint64_t hl = synth_hash(seed ^ 0x5EEDull, r, 0);
    if ((uint32_t)((hl >> 32) % 10000u) < short_per_10000) return (uint32_t)((hl >> 48) % k);
    return min_len + (uint32_t)(hl % (uint64_t)(max_len - min_len + 1));
}

// record sizes (2 len + 17) of reads first_read .. first_read + n_reads - 1
__global__ void __launch_bounds__(256)
synth_var_sizes_kernel(uint64_t* __restrict__ rec, uint64_t n_reads, uint64_t seed, uint64_t first_read, uint32_t min_len,
                       uint32_t max_len, uint32_t short_per_10000, uint32_t k)
{
    pdl_wait();
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_reads; g += (uint64_t)gridDim.x * blockDim.x)
        rec[g] = 2ull * synth_var_len(seed, first_read + g, min_len, max_len, short_per_10000, k) + 17ull;
}

// in-place exclusive prefix sum of n 64-bit values, one CTA (generator only: a sample has < 10^6 reads); v[n] = total
__global__ void __launch_bounds__(1024)
synth_scan_kernel(uint64_t* __restrict__ v, uint64_t n)
{
    pdl_wait();
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t c0 = 0; c0 < n; c0 += 1024) {
        const uint64_t i = c0 + tid;
        const uint64_t x = i < n ? v[i] : 0;
        uint64_t incl = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint64_t wv = s_warp[lane];
            uint64_t wi = wv;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t t = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= (uint32_t)d) wi += t;
            }
            s_warp[lane] = wi - wv;
        }
        __syncthreads();
        const uint64_t carry = s_carry;
        if (i < n) v[i] = carry + s_warp[warp] + incl - x;
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_warp[warp] + incl;
        __syncthreads();
    }
    if (tid == 0) v[n] = s_carry;
}

// one warp per record
__global__ void __launch_bounds__(256)
synth_var_kernel(uint8_t* __restrict__ out, const uint64_t* __restrict__ off, uint64_t n_reads, uint64_t seed,
                 uint64_t first_read)
{
    pdl_wait();
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t rl = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rl < n_reads; rl += n_warps) {
        const uint64_t o0 = off[rl];
        const uint32_t rs = (uint32_t)(off[rl + 1] - o0);
        const uint32_t len = (rs - 17u) >> 1;
        const uint64_t r = first_read + rl;
        for (uint32_t o = lane; o < rs; o += 32) {
            uint8_t c;
            if (o == 0) c = '@';
            else if (o == 1) c = 'S';
            else if (o < 12) {
                uint64_t v = r % 10000000000ull;
                for (uint32_t d = 11; d > o; --d) v /= 10;
                c = (uint8_t)('0' + v % 10);
            } else if (o == 12) c = '\n';
            else if (o < 13 + len) c = synth_base(seed, r, o - 13, len);
            else if (o == 13 + len) c = '\n';
            else if (o == 14 + len) c = '+';
            else if (o == 15 + len) c = '\n';
            else if (o < 16 + 2 * len) c = synth_qual(seed, r, o - 16 - len);
            else c = '\n';
            out[o0 + o] = c;
        }
    }
}

}  // namespace vk
