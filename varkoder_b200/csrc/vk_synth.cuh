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

}  // namespace vk
