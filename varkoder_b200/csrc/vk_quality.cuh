// vk_quality.cuh -- K5: per-position base content of the framed reads, the input of the reference's quality flag.
//
// Stands in for the fastp report that get_basefrequency_sd reads (varKoder/commands/image.py:49-88): it takes fastp's
// "content_curves" (per read position: share of A, T, C, G among all bases at that position), keeps positions 5..39 and
// returns mean over the four bases of the standard deviation along the positions.  fastp classifies a base by
// (byte & 7): A 1, C 3, T 4, N 6, G 7, and divides by the number of reads that reach the position.  This kernel produces
// the integer numerators and denominators from the reads the framing pass has already found; the float64 tail (divide,
// np.std, mean) stays on the host and is the reference's own expression (varkoder_b200/quality.py).
//
// One read per lane; per position one byte load, five warp votes; lanes 0..4 add the vote counts to the warp's private
// shared-memory row (no atomics), rows are summed per CTA and added to the global table once.
#pragma once
#include "vk_common.cuh"

namespace vk {

constexpr int kContentMaxPos = 64;      // positions per call
constexpr int kContentThreads = 256;

__global__ void __launch_bounds__(kContentThreads)
base_content_kernel(const uint8_t* __restrict__ text, const uint64_t* __restrict__ starts, const uint64_t* __restrict__ ends,
                    const Plan* __restrict__ plan, uint32_t pos_begin, uint32_t n_pos, unsigned long long* __restrict__ counts)
{
    pdl_wait();
    constexpr uint32_t FULL = 0xffffffffu;
    __shared__ uint32_t s_cnt[kContentThreads / 32][kContentMaxPos * 5];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < (kContentThreads / 32) * kContentMaxPos * 5; i += blockDim.x) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const uint64_t n_reads = plan->table_overflow ? 0 : plan->n_reads;      // an overflowed table is never dereferenced
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    // at most 2^26 reads per warp row: 32-bit rows cannot overflow
    for (uint64_t r0 = (uint64_t)blockIdx.x * blockDim.x + (tid & ~31u); r0 < n_reads; r0 += stride) {
        const uint64_t r = r0 + lane;
        const uint64_t st = r < n_reads ? starts[r] : 0;
        const uint64_t len = r < n_reads ? ends[r] - st : 0;
        const uint8_t* const p0 = text + st + pos_begin;
        for (uint32_t p = 0; p < n_pos; ++p) {
            const bool has = (uint64_t)pos_begin + p < len;
            const uint32_t all = __ballot_sync(FULL, has);
            if (all == 0) break;                                      // warp-uniform: nobody reaches this position
            const uint32_t b = has ? (uint32_t)__ldg(p0 + p) & 7u : 0u;
            const uint32_t mA = __ballot_sync(FULL, b == 1u), mT = __ballot_sync(FULL, b == 4u);
            const uint32_t mC = __ballot_sync(FULL, b == 3u), mG = __ballot_sync(FULL, b == 7u);
            const uint32_t mine = lane == 0 ? mA : lane == 1 ? mT : lane == 2 ? mC : lane == 3 ? mG : all;
            if (lane < 5) s_cnt[warp][p * 5 + lane] += (uint32_t)__popc(mine);
        }
        __syncwarp();
    }
    __syncthreads();
    for (uint32_t i = tid; i < n_pos * 5; i += blockDim.x) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < kContentThreads / 32; ++w) t += s_cnt[w][i];
        if (t) atomicAdd(counts + i, t);
    }
}

}  // namespace vk
