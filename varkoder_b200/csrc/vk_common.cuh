// vk_common.cuh -- shared definitions of the sm_100a kernels of the varKoder image hot path.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/varkoder_b200.h"

// make -C varkoder_b200/csrc DEBUG=1 builds with device-side assertions on every data-dependent address (the pool's
// compute-sanitizer is closed): a failed one aborts the kernel and the C ABI call returns VK_ECUDA.
#ifdef VK_DEBUG
#include <cassert>
#define VK_ASSERT(c) assert(c)
#else
#define VK_ASSERT(c) ((void)0)
#endif

namespace vk {

constexpr int kMaxLevels = VK_MAX_LEVELS;
constexpr int kNumSMs = 148;              // B200
constexpr uint32_t kUnitReads = 256;      // granularity of the segment regions of the sorted read table
constexpr int kEntryLenBits = 24;         // sorted read entry = start << 24 | len
constexpr uint64_t kEntryLenMask = (1ull << kEntryLenBits) - 1;
constexpr uint64_t kThrAll = ~0ull;

// Device-resident plan: written by the parse / plan / bucket kernels, read by the count / render kernels,
// copied to the host once at the end.  One per context.
struct Plan {
    // ---- parse (K1)
    uint64_t n_bytes;
    uint64_t n_newlines;      // running total over all chunks parsed so far (look-back carry)
    uint64_t sum_starts;      // sum over header-line newlines of (pos + 1), mod 2^64
    uint64_t sum_ends;        // sum over sequence-line newlines of pos, mod 2^64
    uint32_t parse_ticket;    // dynamic tile counter of the current parse launch
    uint32_t table_overflow;  // 1 if a read index exceeded the table capacity
    // ---- plan (K1L)
    uint64_t n_lines, n_reads, nsites_ref, nsites_true, nsites_ladder;
    int32_t status;
    int32_t n_levels;
    uint64_t level_bp[kMaxLevels];
    uint64_t level_thr[kMaxLevels];   // read is in level l iff prio < thr[l]; kThrAll + level_all => every read
    uint32_t level_all[kMaxLevels];
    // ---- bucket (K1c): segment s = reads in levels 0..s and not in s+1
    unsigned long long seg_reads[kMaxLevels];
    unsigned long long seg_bases[kMaxLevels];
    unsigned long long seg_next[kMaxLevels];     // dynamic unit counters of the count kernel
    uint64_t seg_begin[kMaxLevels + 1];          // offsets into the sorted entry array
    uint32_t seg_cta_begin[kMaxLevels + 1];      // CTA ranges of the count kernel
    uint32_t long_reads;                         // reads of 2^24 bases or more that could not be cut into entries (chunk-table mode; more than 4096 of them)
    uint32_t bucket_overflow;                    // a segment received more reads than its region holds (retry, exact layout)
    uint64_t seg_cap[kMaxLevels];                // entries each segment's region of the sorted array can hold
    unsigned long long seg_next2[kMaxLevels];    // second set of unit counters (k = 9: the other half of every CTA pair)
    // chunk table (k <= 7): the reads of a segment cut into 32-byte text chunks, one 8-byte descriptor each (vk_bucket.cuh)
    unsigned long long seg_chunks[kMaxLevels];   // chunks scattered into each segment's region
    uint64_t seg_cbegin[kMaxLevels + 1];         // offsets of the regions in the chunk table
    uint64_t seg_ccap[kMaxLevels];               // descriptors each region holds
    uint64_t chunks_needed;                      // table size the layout asked for (host grows the table and repeats)
    uint32_t chunk_table_small;                  // the layout does not fit the table
    uint32_t count_overflow;                     // a 16-bit bin of a FAST count kernel wrapped (vk_count.cuh): exact recount
    uint64_t read_index_base;                    // global index of this buffer's first record (plan_kernel -> scatter)
    uint64_t total_reads;                        // records of the whole sample (= n_reads unless read-sharded)
    uint32_t hist_ticket;                        // CTAs of thr_calibrate_kernel that are done (the last one fits the thresholds)
    unsigned long long seg_extra[kMaxLevels];    // entries beyond one per read: a read of 2^24 bases or more is several entries (vk_bucket.cuh)
    uint32_t n_long;                             // reads of 2^24 bases or more listed by the scatter kernel (long_reads_kernel cuts them up)
    uint32_t len_min, len_max;                   // shortest / longest read that is counted (scatter kernel): one read per lane pays only for reads of one length
    uint32_t lanes_verdict;                      // k = 7 count kernels: bit 0: the sample is one for countt_kernel (says the flat-lane kernel, which counted it);
                                                 // bit 1: it is not (says countt_kernel, which did not count it: the host repeats the step with the other kernel)
};

// What plan_kernel needs beside the framing: the sample's options and the capacities of the tables.
struct PlanArgs {
    vk_params p;
    uint64_t n_bytes;       // total bytes of the buffer
    uint64_t cap_reads;
    uint64_t cap_sorted;    // entries the sorted array holds
    uint32_t n_count_ctas;  // grid of the count kernel
    uint32_t reads_per_cta; // > 0: use at most ceil(n_reads / reads_per_cta) CTAs (small samples leave SMs to other samples)
    uint32_t exact_layout;  // 1: every segment region holds all reads (retry after a bucket overflow)
    uint32_t test_tight;    // tests only (VK_TEST_TIGHT_BUCKETS=1): regions of half the expected size, to force the retry
    // read-sharded sample (vk_sharded_reads_to_images): the shards' (records, bases) pairs as the all-gather left them on
    // the device, [shard_world][2]; nullptr for an unsharded sample.  The sample-wide base count and this shard's global
    // read index then come from the table instead of p.nsites_override / p.read_index_base.
    uint64_t cap_chunks;    // descriptors the chunk table holds (0: the count kernel reads the sorted read table instead)
    const unsigned long long* shard_table;
    uint32_t shard_rank, shard_world;
};

// Everything about ONE step (one sample through the path) that varies from step to step lives in this device-resident
// block, not in kernel arguments: the kernels of a step are then launched with arguments that never change, so the
// whole step can be captured once as a CUDA graph and replayed for every sample (vk_capi.cu).  The host writes the
// block into pinned memory; the first node of a step copies it to the device.
struct StepArgs {
    const uint8_t* text;    // device pointer of the FASTQ bytes (16-byte aligned, readable up to the next 64-byte boundary)
    uint64_t n_bytes;
    uint32_t n_tiles;       // 32 KiB framing tiles
    uint32_t reserved;
    PlanArgs pa;
};

// splitmix64 finaliser over (seed, global read index): the seeded per-read priority (DESIGN.md "Sub-sampling").
__host__ __device__ __forceinline__ uint64_t prio64(uint64_t seed, uint64_t read_index)
{
    uint64_t z = seed + (read_index + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// number of (nested) levels a read with priority h belongs to; 0 = in no level
__device__ __forceinline__ int levels_of(const Plan* __restrict__ p, int n_levels, uint64_t h)
{
    int c = 0;
    while (c < n_levels && (p->level_all[c] || h < p->level_thr[c])) ++c;
    return c;
}

// ---- index conventions ------------------------------------------------------------------------------
// internal index: base j of the k-mer at bits [2j, 2j+2), 2-bit code (ascii >> 1) & 3  (A0 C1 T2 G3)
// lex index:      base j of the k-mer at bits [2(k-1-j), ...), code A0 C1 G2 T3
__host__ __device__ __forceinline__ uint32_t lex_to_internal(uint32_t x, int k)
{
    uint32_t r = 0;
    for (int j = 0; j < k; ++j) {
        uint32_t d = (x >> (2 * (k - 1 - j))) & 3u;
        r |= (d ^ (d >> 1)) << (2 * j);      // 0,1,2,3 -> 0,1,3,2
    }
    return r;
}
__host__ __device__ __forceinline__ uint32_t internal_revcomp(uint32_t i, int k)
{
    uint32_t r = 0;
    for (int j = 0; j < k; ++j) {
        uint32_t c = (i >> (2 * (k - 1 - j))) & 3u;
        r |= (c ^ 2u) << (2 * j);            // complement under the dsk code is XOR 2
    }
    return r;
}

// ---- chunk descriptor (k <= 7 count path): one 32-byte text chunk of one read --------------------------------------
//   bits  0..32  index of the 16-byte text word the chunk starts at (texts below 2^37 bytes = 137 GB; larger ones count
//                from the read table)
//   bits 33..36  rlo: offset of the read's first base inside its first 16-byte word (every chunk of the read carries it)
//   bits 37..41  hi - 1: the chunk's bytes [lo, hi) belong to the read, lo = rlo in the read's first chunk and 0 after it
//   bits 42..60  j: number of the chunk inside its read (reads of up to 2^24 - 1 bases have fewer than 2^19 chunks)
//   bit  61      this is the read's last chunk
//   bit  62      the read is longer than the break length (reformat.sh breaklength; cut points have to be masked)
//   bit  63      set in every descriptor (0 = empty slot)
constexpr uint64_t kChunkValid = 1ull << 63;
constexpr uint64_t kChunkMaxText = 1ull << 37;
__host__ __device__ __forceinline__ uint64_t make_chunk_desc(uint64_t word16, uint32_t rlo, uint32_t hi, uint32_t j, bool is_last,
                                                             bool is_long)
{
    return kChunkValid | word16 | ((uint64_t)rlo << 33) | ((uint64_t)(hi - 1u) << 37) | ((uint64_t)j << 42) |
           ((uint64_t)(is_last ? 1u : 0u) << 61) | ((uint64_t)(is_long ? 1u : 0u) << 62);
}
__device__ __forceinline__ uint64_t chunk_word16(uint64_t d) { return d & ((1ull << 33) - 1); }
__device__ __forceinline__ uint32_t chunk_rlo(uint64_t d) { return (uint32_t)(d >> 33) & 15u; }
__device__ __forceinline__ uint32_t chunk_hi(uint64_t d) { return ((uint32_t)(d >> 37) & 31u) + 1u; }
__device__ __forceinline__ uint32_t chunk_j(uint64_t d) { return (uint32_t)(d >> 42) & 0x7FFFFu; }
__device__ __forceinline__ bool chunk_last(uint64_t d) { return (d >> 61) & 1u; }
__device__ __forceinline__ bool chunk_long(uint64_t d) { return (d >> 62) & 1u; }

// programmatic dependent launch: block until the preceding kernel of the stream has completed and its writes are
// visible (no-op when the kernel was not launched with programmatic stream serialisation)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t* p, uint64_t v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

}  // namespace vk
