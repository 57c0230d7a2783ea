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

// pass A: reads and bases per segment
__global__ void __launch_bounds__(kBucketThreads)
bucket_count_kernel(const uint64_t* __restrict__ starts, const uint64_t* __restrict__ ends, int k, uint64_t seed,
                    uint64_t read_index_base, Plan* __restrict__ plan)
{
    __shared__ unsigned long long s_reads[kMaxLevels], s_bases[kMaxLevels];
    __shared__ uint32_t s_long;
    const int nl = plan->n_levels;
    const uint64_t n_reads = plan->n_reads;
    if (threadIdx.x < kMaxLevels) { s_reads[threadIdx.x] = 0; s_bases[threadIdx.x] = 0; }
    if (threadIdx.x == 0) s_long = 0;
    __syncthreads();
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t len = ends[r] - starts[r];
        if (len < (uint64_t)k) continue;
        if (len > kEntryLenMask) { atomicAdd(&s_long, 1u); continue; }
        const int c = levels_of(plan, nl, prio64(seed, read_index_base + r));
        if (c == 0) continue;
        atomicAdd(&s_reads[c - 1], 1ull);
        atomicAdd(&s_bases[c - 1], (unsigned long long)len);
    }
    __syncthreads();
    if (threadIdx.x < kMaxLevels && s_reads[threadIdx.x]) {
        atomicAdd(&plan->seg_reads[threadIdx.x], s_reads[threadIdx.x]);
        atomicAdd(&plan->seg_bases[threadIdx.x], s_bases[threadIdx.x]);
    }
    if (threadIdx.x == 0 && s_long) atomicAdd(&plan->long_reads, s_long);
}

// single thread: segment offsets (padded to whole units) and the CTA allocation of the count kernel
__global__ void bucket_layout_kernel(Plan* __restrict__ plan, uint32_t n_count_ctas)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int nl = plan->n_levels;
    uint64_t off = 0;
    int nonempty = 0;
    for (int s = 0; s < kMaxLevels; ++s) {
        plan->seg_begin[s] = off;
        if (s < nl) {
            off += (plan->seg_reads[s] + kUnitReads - 1) / kUnitReads * kUnitReads;
            nonempty += plan->seg_reads[s] != 0;
        }
    }
    plan->seg_begin[kMaxLevels] = off;
    // CTAs per segment: start with one per non-empty segment, then hand out the rest one at a time to the
    // segment with the most bases per CTA (minimises the maximum load; <= 296 x 64 steps)
    uint32_t n_cta[kMaxLevels];
    uint32_t used = 0;
    for (int s = 0; s < kMaxLevels; ++s) { n_cta[s] = (s < nl && plan->seg_reads[s]) ? 1u : 0u; used += n_cta[s]; }
    while (used < n_count_ctas && nonempty > 0) {
        int best = -1;
        double best_load = -1.0;
        for (int s = 0; s < nl; ++s) {
            if (!n_cta[s]) continue;
            // a segment cannot use more CTAs than it has units
            const uint64_t units = (plan->seg_reads[s] + kUnitReads - 1) / kUnitReads;
            if (n_cta[s] >= units) continue;
            const double load = (double)plan->seg_bases[s] / (double)n_cta[s];
            if (load > best_load) { best_load = load; best = s; }
        }
        if (best < 0) break;
        ++n_cta[best];
        ++used;
    }
    uint32_t c = 0;
    for (int s = 0; s < kMaxLevels; ++s) { plan->seg_cta_begin[s] = c; c += n_cta[s]; }
    plan->seg_cta_begin[kMaxLevels] = c;
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
