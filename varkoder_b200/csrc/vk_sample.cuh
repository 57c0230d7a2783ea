// vk_sample.cuh -- K1h / K1t: the level thresholds fitted to the sample's base targets.
//
// Stands in for what `reformat.sh samplebasestarget=<bp>` does beyond drawing reads at random (run_parallel_reformats,
// varKoder/commands/image.py:582-596): it keeps writing reads until the base target is met, so a sub-sample holds its
// target to within a read.  A fixed threshold thr = bp * 2^64 / nsites on the reads' priorities (round 1) meets the target
// only in expectation: the 500 Kbp level of a 200 Mbp sample came out 3 % short.  Here the threshold of every level is
// fitted to the reads actually there:
//   K1h prio_hist_kernel      hist[b] = bases of the reads (all records, also those too short to hold a k-mer) whose priority
//                             falls into bucket b = prio >> 48, 2^16 buckets;
//   K1t thr_calibrate_kernel  prefix over the buckets; for a level with target T the bucket b* in which the cumulative base
//                             count reaches T, and inside it a linear interpolation:
//                                 thr = (b* << 48) + floor((T - C) * 2^48 / hist[b*]),  C = bases in the buckets before b*.
// Reads are still "in level l iff prio < thr[l]" (nested levels, one segment per read, vk_bucket.cuh); what changes is that
// the realised bases are T up to the reads of ONE bucket times the interpolation error: n_reads / 65536 reads, +- 2 reads for
// a 200 Mbp sample.  A sample with fewer than 65536 reads has one read per bucket and the level ends at a read of the
// sorted-by-priority order with probability proportional to the part of it the target covers: within one read, unbiased.
// The oracle restates the same integer arithmetic (oracle/dsk.py calibrated_thresholds).
//
// Read-sharded samples: every shard builds the histogram of its own reads, the shards' histograms are summed (one more
// all-reduce of 512 KiB, issued by the library between K1h and K1t, or by the caller: vk_prio_hist / vk_params.prio_hist).
#pragma once
#include "vk_common.cuh"
#include "vk_parse.cuh"

namespace vk {

constexpr uint32_t kPrioBuckets = VK_PRIO_BUCKETS;
constexpr int kPrioShift = 48;
static_assert(kPrioBuckets == (1u << (64 - kPrioShift)), "bucket = priority >> 48");
constexpr int kPrioHistThreads = 256;       // a light kernel: it waits for table entries and atomics, other samples' kernels fit beside it

// K1h.  hist[65536], zeroed by the first kernel of the step (or by the caller: vk_prio_hist).  One global atomic per read.
// force: vk_prio_hist -- the caller's buffer, whatever the sample's options say.
__global__ void __launch_bounds__(kPrioHistThreads)
prio_hist_kernel(const uint64_t* __restrict__ starts, const uint64_t* __restrict__ ends, const StepArgs* __restrict__ sa,
                 const Plan* __restrict__ plan, unsigned long long* __restrict__ hist, int force)
{
    pdl_wait();
    // nothing to do: expected-value thresholds, or the caller brings the histogram of the whole sample
    if (!force && (sa->pa.p.sampling != VK_SAMPLING_CALIBRATED || sa->pa.p.prio_hist != 0)) return;
    if (plan->table_overflow) return;                      // the read table does not hold every read: the step is repeated
    const uint64_t n_reads = plan->n_reads;
    const uint64_t seed = sa->pa.p.seed, base = force ? sa->pa.p.read_index_base : plan->read_index_base;
    // eight reads per thread and round, their table entries requested together (one read per round left the kernel waiting
    // for DRAM nine times in a row)
    constexpr int U = 8;
    for (uint64_t r0 = (uint64_t)blockIdx.x * (blockDim.x * U); r0 < n_reads; r0 += (uint64_t)gridDim.x * (blockDim.x * U)) {
        uint64_t st[U], en[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint64_t r = r0 + (uint64_t)j * blockDim.x + threadIdx.x;
            st[j] = r < n_reads ? starts[r] : 0;
            en[j] = r < n_reads ? ends[r] : 0;
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint64_t r = r0 + (uint64_t)j * blockDim.x + threadIdx.x;
            const uint64_t len = en[j] - st[j];
            if (r < n_reads && len != 0) atomicAdd(hist + (prio64(seed, base + r) >> kPrioShift), (unsigned long long)len);
        }
    }
}

// K1t: 64 CTAs of 1024 threads.  CTA c sums its 1024 buckets in blocks of 64 (block j = buckets [64 j, 64 j + 64)) into
// coarse[16 c ..]; the CTA that finishes last scans the 1024 block sums and fits the levels, warp l level l.
constexpr uint32_t kCalibCtas = kPrioBuckets / 1024;
__global__ void __launch_bounds__(1024)
thr_calibrate_kernel(const StepArgs* __restrict__ sa, const unsigned long long* __restrict__ hist_own,
                     unsigned long long* __restrict__ coarse, Plan* __restrict__ plan)
{
    pdl_wait();
    if (sa->pa.p.sampling != VK_SAMPLING_CALIBRATED) return;
    const unsigned long long* __restrict__ hist =
        sa->pa.p.prio_hist ? reinterpret_cast<const unsigned long long*>(sa->pa.p.prio_hist) : hist_own;
    const int nl = plan->n_levels;
    if (nl <= 0 || plan->table_overflow) return;
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr uint32_t PER = 64;                            // buckets per block: two per lane in the second step
    constexpr uint32_t NB = kPrioBuckets / PER;             // 1024 blocks
    static_assert(NB == 1024, "one block per thread in the scan");
    __shared__ unsigned long long s_excl[NB];               // bases in the blocks before each block
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_total;                  // bases of all reads
    __shared__ uint32_t s_last;
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5;
    {
        unsigned long long x = hist[(size_t)blockIdx.x * 1024 + t];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(FULL, x, d);
        if (lane == 0) s_warp[w] = x;                       // half-block sums: warps 2m, 2m + 1 = block 16 c + m
        __syncthreads();
        if (t < 16) coarse[blockIdx.x * 16 + t] = s_warp[2 * t] + s_warp[2 * t + 1];
        __threadfence();
        __syncthreads();
        if (t == 0) s_last = atomicAdd(&plan->hist_ticket, 1u) == gridDim.x - 1u ? 1u : 0u;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
    }
    const unsigned long long mine = reinterpret_cast<const volatile unsigned long long*>(coarse)[t];
    unsigned long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(FULL, incl, d);
        if (lane >= (uint32_t)d) incl += o;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        unsigned long long x = s_warp[lane], xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(FULL, xi, d);
            if (lane >= (uint32_t)d) xi += o;
        }
        s_warp[lane] = xi - x;                              // exclusive over warps
    }
    __syncthreads();
    s_excl[t] = s_warp[w] + incl - mine;
    if (t == 1023) s_total = s_warp[w] + incl;
    __syncthreads();
    // ---- warp l: level l
    for (int l = (int)w; l < nl; l += 32) {
        if (plan->level_all[l]) continue;
        const unsigned long long T = plan->level_bp[l];
        if (T >= s_total) {                                 // the target is not met before the last read: every read
            if (lane == 0) { plan->level_all[l] = 1u; plan->level_thr[l] = kThrAll; }
            continue;
        }
        // the block in which the cumulative count reaches T: last j with excl[j] < T
        uint32_t lo = 0, hi = NB - 1;                       // excl[0] = 0 < T (level targets are positive)
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (s_excl[mid] < T) lo = mid; else hi = mid - 1;
        }
        // inside it: lane j holds buckets 2j, 2j + 1
        const ulonglong2 v = reinterpret_cast<const ulonglong2*>(hist + (size_t)lo * PER)[lane];
        unsigned long long pin = v.x + v.y;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(FULL, pin, d);
            if (lane >= (uint32_t)d) pin += o;
        }
        const unsigned long long before = s_excl[lo] + pin - (v.x + v.y);      // bases before bucket 2j of this block
        // first bucket whose inclusive cumulative count is >= T
        const bool hit0 = before + v.x >= T, hit1 = before + v.x + v.y >= T;
        const uint32_t m = __ballot_sync(FULL, hit1);
        const int src = __ffs(m) - 1;                       // m != 0: the block's inclusive total is >= T by construction
        if ((int)lane == src) {
            const uint32_t b = lo * PER + 2 * lane + (hit0 ? 0u : 1u);
            const unsigned long long C = hit0 ? before : before + v.x;
            const unsigned long long W = hit0 ? v.x : v.y;                      // > 0
            const unsigned long long rem = T - C;                               // 1 .. W
            unsigned long long thr;
            if (rem >= W) thr = b + 1u < kPrioBuckets ? (unsigned long long)(b + 1u) << kPrioShift : kThrAll;
            else thr = ((unsigned long long)b << kPrioShift) + (div_2p64(rem, W) >> (64 - kPrioShift));
            plan->level_thr[l] = thr;
            if (thr == kThrAll) plan->level_all[l] = 1u;
        }
    }
}

}  // namespace vk
