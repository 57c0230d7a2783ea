// vk_bucket.cuh -- K1b: seeded nested sub-sampling as a counting sort of the read table by ladder segment.
//
// Stands in for the L independent `reformat.sh samplebasestarget=... sampleseed=seed+i` runs of
// run_parallel_reformats (varKoder/commands/image.py:577-627).  The Java RNG stream cannot be reproduced, so
// the rule is this project's own (DESIGN.md "Sub-sampling"): read r belongs to level l iff
// prio64(seed, r) < floor(level_bp[l] * 2^64 / nsites).  Levels are nested, so every read has one
// *segment* s = (number of levels it is in) - 1, and the histogram of level l is the sum of the
// histograms of segments >= l: one pass over the reads serves every level.
//
// The read table (16 B per read, ~1/20 of the text) is sorted by segment so that a CTA of the count
// kernel only ever sees reads of one segment and can keep a single shared-memory histogram.
#pragma once
#include "vk_common.cuh"

namespace vk {

constexpr int kBucketThreads = 256;
constexpr int kBucketCountThreads = 1024;

// pass A: reads and bases per segment.  One CTA per SM; a warp tallies its 32 reads per segment with a ballot and
// a redux, lane (s mod 32) keeps the running totals of segment s in registers, so the only atomics are a few
// per warp at the very end (64-bit shared-memory atomics are CAS loops on sm_100 and must stay off the hot loop).
__global__ void __launch_bounds__(kBucketCountThreads)
bucket_count_kernel(const uint64_t* __restrict__ starts, const uint64_t* __restrict__ ends, int k, uint64_t seed,
                    uint64_t read_index_base, Plan* __restrict__ plan)
{
    __shared__ unsigned long long s_reads[kMaxLevels], s_bases[kMaxLevels];
    __shared__ uint32_t s_long;
    const int nl = plan->n_levels;
    const uint64_t n_reads = plan->n_reads;
    const uint32_t lane = threadIdx.x & 31;
    if (threadIdx.x < kMaxLevels) { s_reads[threadIdx.x] = 0; s_bases[threadIdx.x] = 0; }
    if (threadIdx.x == 0) s_long = 0;
    __syncthreads();
    unsigned long long my_reads[2] = {0, 0}, my_bases[2] = {0, 0};
    uint32_t my_long = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_iter = (n_reads + stride - 1) / stride;
    for (uint64_t it = 0; it < n_iter; ++it) {
        const uint64_t r = it * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        int seg = -1;
        uint32_t len32 = 0;
        if (r < n_reads) {
            const uint64_t len = ends[r] - starts[r];
            if (len > kEntryLenMask) ++my_long;
            else if (len >= (uint64_t)k) {
                seg = levels_of(plan, nl, prio64(seed, read_index_base + r)) - 1;
                len32 = (uint32_t)len;
            }
        }
        if (__ballot_sync(0xffffffffu, seg >= 0) == 0) continue;
        for (int s = 0; s < nl; ++s) {
            const uint32_t m = __ballot_sync(0xffffffffu, seg == s);
            if (m == 0) continue;
            const uint32_t sum = __reduce_add_sync(0xffffffffu, seg == s ? len32 : 0u);    // 32 x 2^24 fits
            if (lane == (uint32_t)(s & 31)) {
                my_reads[s >> 5] += __popc(m);
                my_bases[s >> 5] += sum;
            }
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (my_reads[h]) {
            atomicAdd(&s_reads[lane + 32 * h], my_reads[h]);
            atomicAdd(&s_bases[lane + 32 * h], my_bases[h]);
        }
    }
    if (my_long) atomicAdd(&s_long, my_long);
    __syncthreads();
    if (threadIdx.x < kMaxLevels && s_reads[threadIdx.x]) {
        atomicAdd(&plan->seg_reads[threadIdx.x], s_reads[threadIdx.x]);
        atomicAdd(&plan->seg_bases[threadIdx.x], s_bases[threadIdx.x]);
    }
    if (threadIdx.x == 0 && s_long) atomicAdd(&plan->long_reads, s_long);
}

// one warp: segment offsets (padded to whole units) and the CTA allocation of the count kernel.
// lane l owns segments l and l + 32.
__global__ void __launch_bounds__(32)
bucket_layout_kernel(Plan* __restrict__ plan, uint32_t n_count_ctas)
{
    const uint32_t lane = threadIdx.x;
    const int nl = plan->n_levels;
    uint64_t reads[2], bases[2], padded[2];
    uint32_t n_cta[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int s = lane + 32 * h;
        reads[h] = s < nl ? plan->seg_reads[s] : 0;
        bases[h] = s < nl ? plan->seg_bases[s] : 0;
        padded[h] = (reads[h] + kUnitReads - 1) / kUnitReads * kUnitReads;
        n_cta[h] = reads[h] ? 1u : 0u;
    }
    // exclusive scan of the padded sizes over the 64 segments
    uint64_t run = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint64_t incl = padded[h];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        plan->seg_begin[lane + 32 * h] = run + incl - padded[h];
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) plan->seg_begin[kMaxLevels] = run;

    // CTAs: one per non-empty segment, the rest in proportion to the bases, leftovers one by one to the segment
    // with the most bases per CTA (minimises the slowest segment)
    uint64_t total_bases = bases[0] + bases[1];
    uint32_t used = n_cta[0] + n_cta[1];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        total_bases += __shfl_xor_sync(0xffffffffu, total_bases, d);
        used += __shfl_xor_sync(0xffffffffu, used, d);
    }
    if (used > 0 && used < n_count_ctas && total_bases > 0) {
        const uint64_t spare = n_count_ctas - used;
        uint32_t extra = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t e = (uint32_t)(spare * bases[h] / total_bases);      // exact floor: spare < 2^16, bases < 2^47
            n_cta[h] += reads[h] ? e : 0;
            extra += reads[h] ? e : 0;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) extra += __shfl_xor_sync(0xffffffffu, extra, d);
        uint32_t left = (uint32_t)spare - extra;               // fewer than the number of non-empty segments
        while (left > 0) {
            double best = -1.0;
            int best_s = -1;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (n_cta[h]) {
                    const double load = (double)bases[h] / (double)n_cta[h];
                    if (load > best) { best = load; best_s = lane + 32 * h; }
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, d);
                const int os = __shfl_xor_sync(0xffffffffu, best_s, d);
                if (ob > best || (ob == best && os >= 0 && (best_s < 0 || os < best_s))) { best = ob; best_s = os; }
            }
            if (best_s < 0) break;
            if ((uint32_t)(best_s & 31) == lane) ++n_cta[best_s >> 5];
            --left;
        }
    }
    uint32_t crun = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t incl = n_cta[h];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        plan->seg_cta_begin[lane + 32 * h] = crun + incl - n_cta[h];
        crun += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) plan->seg_cta_begin[kMaxLevels] = crun;
}

// pass B: scatter (start, len) entries into their segment's range
__global__ void __launch_bounds__(kBucketThreads)
bucket_scatter_kernel(const uint64_t* __restrict__ starts, const uint64_t* __restrict__ ends, int k, uint64_t seed,
                      uint64_t read_index_base, uint64_t text_base, uint64_t* __restrict__ sorted,
                      Plan* __restrict__ plan)
{
    __shared__ uint32_t s_cnt[kMaxLevels];
    __shared__ unsigned long long s_base[kMaxLevels];
    const int nl = plan->n_levels;
    const uint64_t n_reads = plan->n_reads;
    const uint64_t per_iter = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r0 = (uint64_t)blockIdx.x * blockDim.x; r0 < n_reads; r0 += per_iter) {
        if (threadIdx.x < kMaxLevels) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        const uint64_t r = r0 + threadIdx.x;
        int seg = -1;
        uint32_t rank = 0;
        uint64_t entry = 0;
        if (r < n_reads) {
            const uint64_t st = starts[r];
            const uint64_t len = ends[r] - st;
            if (len >= (uint64_t)k && len <= kEntryLenMask) {
                seg = levels_of(plan, nl, prio64(seed, read_index_base + r)) - 1;
                if (seg >= 0) {
                    rank = atomicAdd(&s_cnt[seg], 1u);
                    entry = ((st - text_base) << kEntryLenBits) | len;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < kMaxLevels && s_cnt[threadIdx.x])
            s_base[threadIdx.x] = atomicAdd(&plan->seg_cursor[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
        __syncthreads();
        if (seg >= 0) sorted[plan->seg_begin[seg] + s_base[seg] + rank] = entry;
        __syncthreads();
    }
}

}  // namespace vk
