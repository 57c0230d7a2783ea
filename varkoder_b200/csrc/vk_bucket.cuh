// vk_bucket.cuh -- K1c: seeded nested sub-sampling as a one-pass scatter of the read table into ladder segments.
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
constexpr int kBucketItems = 4;      // reads per thread and iteration of the scatter kernel (read-table mode)
constexpr int kBucketItemsChunk = 2; // chunk mode: half as many, so that the staging area stays small enough for 8 blocks per SM
constexpr uint32_t kStageChunks = 3072;      // chunk descriptors a block stages in (dynamic) shared memory, 24 KiB: 512 reads of up to ~170 bases

// A read of 2^24 bases or more (a chromosome) does not fit one entry of the read table: it becomes several.  With a break
// length the pieces end at multiples of it (no k-mer spans a cut point anyway, image.py:586); without one they overlap by
// k - 1 bases, so that every k-mer starts in exactly one piece.  The scatter kernel only lists such reads (they are rare
// and the arithmetic -- 64-bit divisions by the break length -- would cost its loop a third of its occupancy); this kernel,
// one thread per listed read, cuts them up and takes the slots straight from the segments' counters.
constexpr uint32_t kLongListCap = 4096;
__global__ void __launch_bounds__(256)
long_reads_kernel(const uint64_t* __restrict__ starts, const uint64_t* __restrict__ ends, const StepArgs* __restrict__ sa,
                  uint64_t text_base, uint64_t* __restrict__ sorted, Plan* __restrict__ plan, const uint64_t* __restrict__ long_list)
{
    pdl_wait();
    const uint32_t n = plan->n_long;
    if (n == 0 || plan->table_overflow || plan->bucket_overflow) return;
    if (n > kLongListCap) { if (threadIdx.x == 0 && blockIdx.x == 0) plan->long_reads = n; return; }
    const uint32_t k = (uint32_t)sa->pa.p.k, breaklen = (uint32_t)sa->pa.p.breaklength;
    const int nl = plan->n_levels;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint64_t r = long_list[i];
        const uint64_t start = starts[r] - text_base, len = ends[r] - starts[r];
        const int sg = levels_of(plan, nl, prio64(sa->pa.p.seed, plan->read_index_base + r)) - 1;
        if (sg < 0) continue;
        const uint64_t step = breaklen ? (kEntryLenMask / breaklen) * breaklen : kEntryLenMask - (uint64_t)(k - 1);
        const uint64_t ov = breaklen ? 0 : (uint64_t)(k - 1);
        const uint64_t n_pieces = (len + step - 1) / step;
        const unsigned long long slot0 = atomicAdd(&plan->seg_reads[sg], (unsigned long long)n_pieces);
        atomicAdd(&plan->seg_bases[sg], (unsigned long long)len);
        atomicAdd(&plan->seg_extra[sg], (unsigned long long)(n_pieces - 1));
        const uint64_t begin = plan->seg_begin[sg], cap = plan->seg_cap[sg];
        for (uint64_t j = 0; j < n_pieces; ++j) {
            const uint64_t off = j * step;
            const uint64_t pl = len - off < step + ov ? len - off : step + ov;
            if (slot0 + j < cap) sorted[begin + slot0 + j] = ((start + off) << kEntryLenBits) | pl;
            else plan->bucket_overflow = 1u;
        }
    }
}

// One pass: scatter (start, len) entries into their segment's region, count reads and bases per segment.
// Regions were sized from the expected segment shares by plan_kernel (vk_parse.cuh); a read that does not fit
// raises plan->bucket_overflow and the host repeats the step with regions that hold every read.
template <int kBucketItems>
// (forcing 6 or 8 blocks per SM through the launch bounds spills and is slower: 720 / 728 against 746 Gbases/s)
__global__ void __launch_bounds__(kBucketThreads)
bucket_scatter_kernel(const uint64_t* __restrict__ starts, const uint64_t* __restrict__ ends, const StepArgs* __restrict__ sa,
                      uint64_t text_base, uint64_t* __restrict__ sorted, uint64_t* __restrict__ chunks, Plan* __restrict__ plan,
                      uint64_t* __restrict__ long_list)
{
    pdl_wait();
    const int k = sa->pa.p.k;
    const uint64_t seed = sa->pa.p.seed, read_index_base = plan->read_index_base;
    constexpr uint32_t FULL = 0xffffffffu;
    __shared__ uint32_t s_cnt[kMaxLevels], s_all[kMaxLevels];
    __shared__ unsigned long long s_len[kMaxLevels];
    __shared__ unsigned long long s_base[kMaxLevels];
    __shared__ uint64_t s_thr[kMaxLevels], s_begin[kMaxLevels], s_cap[kMaxLevels];
    __shared__ uint32_t s_long, s_lmin, s_lmax;
    // chunk mode (k <= 7, chunks != nullptr): instead of one entry per read, one descriptor per 32-byte chunk of the read
    // goes into the segment's region of the chunk table (vk_common.cuh make_chunk_desc); the count kernel then walks
    // descriptors, coalesced, with nothing to work out per chunk
    __shared__ uint32_t s_ccnt[kMaxLevels];
    __shared__ unsigned long long s_cbase[kMaxLevels];
    __shared__ uint64_t s_cbegin[kMaxLevels], s_ccap[kMaxLevels];
    extern __shared__ uint64_t s_stage[];      // kStageChunks descriptors (chunk mode)
    const bool chunk_mode = chunks != nullptr;      // (as a compile-time constant ptxas takes 80 registers instead of 48: 22.7 -> 29.7 us)
    const uint32_t breaklen = (uint32_t)sa->pa.p.breaklength;
    // the tables do not fit (plan_kernel): the step is repeated with larger ones, nothing here may be dereferenced.
    // (bucket_overflow can also be raised by this kernel itself, below; a CTA that starts late and sees it only skips
    // work whose result is thrown away.)
    __shared__ uint32_t s_abort;
    if (threadIdx.x == 0) s_abort = plan->table_overflow | plan->bucket_overflow | plan->chunk_table_small;      // one reader: the branch is block-uniform
    __syncthreads();
    if (s_abort) return;
    const int nl = plan->n_levels;
    const uint64_t n_reads = plan->n_reads;
    const uint64_t per_iter = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31;
    if (threadIdx.x < kMaxLevels) {
        s_thr[threadIdx.x] = plan->level_thr[threadIdx.x];
        s_all[threadIdx.x] = plan->level_all[threadIdx.x];
        s_begin[threadIdx.x] = plan->seg_begin[threadIdx.x];
        s_cap[threadIdx.x] = plan->seg_cap[threadIdx.x];
        s_cbegin[threadIdx.x] = plan->seg_cbegin[threadIdx.x];
        s_ccap[threadIdx.x] = plan->seg_ccap[threadIdx.x];
    }
    if (threadIdx.x == 0) { s_long = 0; s_lmin = 0xFFFFFFFFu; s_lmax = 0; }
    uint32_t my_min = 0xFFFFFFFFu, my_max = 0;
    // kBucketItems reads per thread and iteration: their table loads are in flight together, and one round trip of the
    // global cursors serves 1024 reads (the kernel is latency-bound: ~1 iteration per CTA at 200 Mbp)
    for (uint64_t r0 = (uint64_t)blockIdx.x * (blockDim.x * kBucketItems); r0 < n_reads; r0 += per_iter * kBucketItems) {
        if (threadIdx.x < kMaxLevels) { s_cnt[threadIdx.x] = 0; s_len[threadIdx.x] = 0; s_ccnt[threadIdx.x] = 0; }
        __syncthreads();
        uint64_t st[kBucketItems], en[kBucketItems];
#pragma unroll
        for (int i = 0; i < kBucketItems; ++i) {
            const uint64_t r = r0 + (uint64_t)i * blockDim.x + threadIdx.x;
            st[i] = r < n_reads ? starts[r] : 0;
            en[i] = r < n_reads ? ends[r] : 0;
        }
        int seg[kBucketItems];
        uint32_t rank[kBucketItems];
        uint64_t entry[kBucketItems];
        uint32_t nchunk[kBucketItems], crank[kBucketItems];
#pragma unroll
        for (int i = 0; i < kBucketItems; ++i) {
            const uint64_t r = r0 + (uint64_t)i * blockDim.x + threadIdx.x;
            seg[i] = -1;
            entry[i] = 0;
            uint32_t len32 = 0;
            if (r < n_reads) {
                const uint64_t len = en[i] - st[i];
                if (len > kEntryLenMask) {
                    // several table entries: listed here, cut up by long_reads_kernel (keeps this loop at 48 registers)
                    if (chunk_mode) atomicAdd(&s_long, 1u);
                    else {
                        const uint32_t slot = atomicAdd(&plan->n_long, 1u);
                        if (slot < kLongListCap) long_list[slot] = r;
                        my_max = (uint32_t)kEntryLenMask;
                        my_min = my_min < (uint32_t)kEntryLenMask ? my_min : (uint32_t)kEntryLenMask;
                    }
                } else if (len >= (uint64_t)k) {
                    const uint64_t h = prio64(seed, read_index_base + r);
                    // number of levels the read is in: membership is monotone in the level (levels are nested: the
                    // "all reads" levels come first, thresholds never increase), so a binary search with a
                    // warp-uniform trip count replaces the divergent walk down the ladder
                    int lo = 0, hi = nl;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (s_all[mid] || h < s_thr[mid]) lo = mid + 1; else hi = mid;
                    }
                    seg[i] = lo - 1;
                    if (seg[i] >= 0) {
                        len32 = (uint32_t)len;
                        my_min = len32 < my_min ? len32 : my_min;
                        my_max = len32 > my_max ? len32 : my_max;
                        entry[i] = ((st[i] - text_base) << kEntryLenBits) | len;
                    }
                }
            }
            // warp-aggregated: one shared-memory atomic per (warp, segment) instead of one per read
            const uint32_t peers = __match_any_sync(FULL, seg[i]);
            const uint32_t lsum = __reduce_add_sync(peers, len32);               // 32 x 2^24 fits 32 bits
            const int leader = __ffs(peers) - 1;
            uint32_t wbase = 0;
            if ((int)lane == leader && seg[i] >= 0) {
                wbase = atomicAdd(&s_cnt[seg[i]], (uint32_t)__popc(peers));
                atomicAdd(&s_len[seg[i]], (unsigned long long)lsum);
            }
            wbase = __shfl_sync(peers, wbase, leader);
            rank[i] = wbase + __popc(peers & ((1u << lane) - 1u));
            nchunk[i] = 0;
            crank[i] = 0;
            if (chunk_mode) {
                // chunks of this read, and its offset among the chunks of the warp's reads of the same segment: one plain
                // warp scan per segment present in the warp (a handful: the segments halve in size down the ladder)
                const uint32_t rlo = (uint32_t)(entry[i] >> kEntryLenBits) & 15u;
                const uint32_t n = seg[i] >= 0 ? (rlo + len32 + 31u) >> 5 : 0u;
                nchunk[i] = n;
                uint32_t todo = __ballot_sync(FULL, seg[i] >= 0);
                while (todo) {
                    const int cur = __shfl_sync(FULL, seg[i], __ffs(todo) - 1);
                    const bool mine = seg[i] == cur;
                    uint32_t incl = mine ? n : 0u;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t t = __shfl_up_sync(FULL, incl, d);
                        if (lane >= (uint32_t)d) incl += t;
                    }
                    const uint32_t total = __shfl_sync(FULL, incl, 31);
                    uint32_t cb = 0;
                    if (mine && (int)lane == leader) cb = atomicAdd(&s_ccnt[cur], total);      // leader = lowest lane of the segment
                    cb = __shfl_sync(FULL, cb, __ffs(todo) - 1);
                    if (mine) crank[i] = cb + incl - n;
                    todo &= ~__ballot_sync(FULL, mine);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < kMaxLevels && s_cnt[threadIdx.x]) {
            s_base[threadIdx.x] = atomicAdd(&plan->seg_reads[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
            atomicAdd(&plan->seg_bases[threadIdx.x], s_len[threadIdx.x]);
            if (chunk_mode) s_cbase[threadIdx.x] = atomicAdd(&plan->seg_chunks[threadIdx.x], (unsigned long long)s_ccnt[threadIdx.x]);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kBucketItems; ++i) {
            if (seg[i] >= 0 && !chunk_mode) {
                const uint64_t slot = s_base[seg[i]] + rank[i];
                VK_ASSERT(seg[i] < nl && s_begin[seg[i]] + s_cap[seg[i]] <= plan->seg_begin[kMaxLevels]);
                if (slot < s_cap[seg[i]]) sorted[s_begin[seg[i]] + slot] = entry[i];
                else plan->bucket_overflow = 1u;
            }
        }
        if (chunk_mode) {
            // The descriptors of the block's reads go through shared memory: per segment they form ONE contiguous range
            // of the segment's region (s_cbase .. + s_ccnt), so the block can write them with coalesced 8-byte stores
            // (a thread writing its own read's five descriptors touches 32 different sectors per store instruction:
            // 61 us instead of 23 us for the whole kernel).  A block whose reads have more chunks than the staging
            // area holds (reads of thousands of bases) writes directly.
            __shared__ uint32_t s_coff[kMaxLevels + 1];
            if (threadIdx.x == 0) {
                uint32_t run = 0;
                for (int sg = 0; sg < kMaxLevels; ++sg) { s_coff[sg] = run; run += s_ccnt[sg]; }
                s_coff[kMaxLevels] = run;
            }
            __syncthreads();
            const bool staged = s_coff[kMaxLevels] <= kStageChunks;
#pragma unroll
            for (int i = 0; i < kBucketItems; ++i) {
                if (seg[i] < 0) continue;
                const uint64_t first = s_cbase[seg[i]] + crank[i];
                const uint32_t n = nchunk[i];
                if (first + n > s_ccap[seg[i]]) { plan->bucket_overflow = 1u; continue; }
                const uint64_t rstart = entry[i] >> kEntryLenBits;
                const uint32_t rlen = (uint32_t)(entry[i] & kEntryLenMask);
                const uint32_t rlo = (uint32_t)rstart & 15u;
                const bool is_long = breaklen != 0 && rlen > breaklen;
                uint64_t* const dst = staged ? s_stage + s_coff[seg[i]] + crank[i] : chunks + s_cbegin[seg[i]] + first;
                for (uint32_t j = 0; j < n; ++j) {
                    const uint32_t endrel = rlo + rlen - 32u * j;               // > 0
                    dst[j] = make_chunk_desc((rstart >> 4) + 2ull * j, rlo, endrel < 32u ? endrel : 32u, j, j + 1 == n, is_long);
                }
            }
            __syncthreads();
            if (staged) {
                for (int sg = 0; sg < nl; ++sg) {
                    const uint32_t cnt = s_ccnt[sg];
                    if (cnt == 0 || s_cbase[sg] + cnt > s_ccap[sg]) continue;      // (overflow: flagged above, nothing written)
                    uint64_t* const dst = chunks + s_cbegin[sg] + s_cbase[sg];
                    const uint64_t* const src = s_stage + s_coff[sg];
                    for (uint32_t t = threadIdx.x; t < cnt; t += blockDim.x) dst[t] = src[t];
                }
            }
        }
        __syncthreads();
    }
    // shortest / longest counted read of the sample (which count kernel pays off depends on it, vk_countu.cuh)
    my_min = __reduce_min_sync(FULL, my_min);
    my_max = __reduce_max_sync(FULL, my_max);
    __syncthreads();
    if (lane == 0 && my_max != 0u) { atomicMin(&s_lmin, my_min); atomicMax(&s_lmax, my_max); }
    __syncthreads();
    if (threadIdx.x == 0 && s_lmax != 0u) { atomicMin(&plan->len_min, s_lmin); atomicMax(&plan->len_max, s_lmax); }
    if (threadIdx.x == 0 && s_long) atomicAdd(&plan->long_reads, s_long);
}

}  // namespace vk
