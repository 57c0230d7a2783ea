// vk_count.cuh -- K2: k-mer counting of every ladder segment in one pass over the sequence lines.
//
// Stands in for `dsk -kmer-size k -abundance-min 1` run once per sub-sample file
// (varKoder/commands/image.py:771-796), restated per SURVEY.md section 8c:
//   D1 k-mers never span records (nor the <=500-base pieces reformat.sh breaklength=500 cuts, image.py:586)
//   D2 2-bit code (ascii >> 1) & 3;  D3 a window holding a non-ACGT byte is dropped;
//   D4/D5 canonical abundance = forward count of K + forward count of rc(K): the fold happens after
//   counting (vk_image.cuh), so the hot loop only builds a forward-strand histogram.
//
// Design (DESIGN.md "K2"): one thread walks one read, one aligned 16-byte word per step; the 16 bases are
// packed to 2 bits with SIMD-in-register arithmetic and joined to the previous k-1 codes in a 64-bit window,
// so each k-mer is one funnel shift + mask.  Validity (ACGT, read bounds, break points) is a bit mask whose
// k-long runs give the 16-bit "emit" mask.  Every lane issues its shared-memory increment on every step
// (invalid positions go to a per-lane trash word) because a fully populated ATOMS.POPC.INC costs the same as
// a sparse one (tools/microbench_atomics2.cu: 12.2 increments/clk/SM full, 4.0 at 52 % lanes with branches).
// Reads arrive sorted by segment (vk_bucket.cuh), a CTA serves one segment and keeps one 4^k x u32 histogram
// in shared memory, written once at the end to its private slab (plain stores; no global atomics).
#pragma once
#include "vk_common.cuh"

namespace vk {

constexpr int kCountThreads = 1024;

// 2-bit codes of the 4 bytes of x, densely packed: byte i -> bits [2i, 2i+2)
__device__ __forceinline__ uint32_t pack4(uint32_t x)
{
    const uint32_t y = (x >> 1) & 0x03030303u;
    return (y * 0x01041040u) >> 24;      // byte i (at bit 8i) lands at bit 24 + 2i; partial products never overlap
}
// 4-bit mask of the bytes of x that are one of ACGTacgt
__device__ __forceinline__ uint32_t acgt4(uint32_t x)
{
    // bits (b2 b1) are the code; for a valid letter the other bits are forced: b7=0 b6=1 b3=0,
    // b4 = (code == T) = b2 & ~b1, b0 = ~b4; b5 is the case bit
    const uint32_t t = (x >> 2) & ~(x >> 1);                 // bit0 of each byte: b2 & ~b1
    const uint32_t ok = ~((x >> 4) ^ t) & (x ^ t) & ~(x >> 3) & (x >> 6) & ~(x >> 7) & 0x01010101u;
    return (ok * 0x01020408u) >> 24;
}

template <int K>
__device__ __forceinline__ uint32_t runs_of_k(uint32_t m)
{
    // bit j of the result = bits j .. j+K-1 of m are all set
    uint32_t r = m;
    int len = 1;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        if (len * 2 <= K) { r &= r >> len; len *= 2; }
    }
    if (len < K) r &= r >> (K - len);
    return r;
}

// SMEM = true : histogram of the CTA's segment in shared memory (k <= 7), flushed to slabs[blockIdx.x]
// SMEM = false: increments go straight to the global (L2-resident) segment histogram (k = 8, 9)
template <int K, bool SMEM>
__global__ void __launch_bounds__(kCountThreads, 1)
count_kernel(const uint4* __restrict__ text16, const uint64_t* __restrict__ sorted, Plan* __restrict__ plan,
             uint32_t* __restrict__ slabs, unsigned long long* __restrict__ seg_hist, int breaklen)
{
    constexpr uint32_t NK = 1u << (2 * K);
    constexpr uint32_t KMASK = NK - 1;
    constexpr int KM1 = K - 1;
    extern __shared__ uint32_t s_hist[];          // SMEM: NK bins + 32 trash words
    const uint32_t tid = threadIdx.x, lane = tid & 31;

    // ---- which segment does this CTA serve?
    int seg = -1;
    for (int s = 0; s < kMaxLevels; ++s)
        if (blockIdx.x >= plan->seg_cta_begin[s] && blockIdx.x < plan->seg_cta_begin[s + 1]) seg = s;
    if (seg < 0) return;
    const uint64_t seg_b = plan->seg_begin[seg];
    const uint64_t seg_e = seg_b + plan->seg_reads[seg];
    unsigned long long* const gh = SMEM ? nullptr : seg_hist + (size_t)seg * NK;

    if (SMEM) {
        for (uint32_t i = tid; i < NK + 32; i += kCountThreads) s_hist[i] = 0;
        __syncthreads();
    }
    const uint32_t trash = NK + lane;

    // ---- read distribution: the warps serving a segment take its sorted reads in units of 32 (one per lane,
    // neighbouring lanes = neighbouring records), unit u going to warp (u mod number-of-warps); inside a warp a
    // lane that finishes its read takes the next one of the warp's current unit, so lanes stay busy whatever
    // the read lengths are and no global atomic is needed.
    const uint32_t seg_ctas = plan->seg_cta_begin[seg + 1] - plan->seg_cta_begin[seg];
    const uint64_t unit_stride = (uint64_t)seg_ctas * (kCountThreads / 32) * 32;                 // in reads
    uint64_t unit_at = seg_b + ((uint64_t)(blockIdx.x - plan->seg_cta_begin[seg]) * (kCountThreads / 32) + (tid >> 5)) * 32;
    uint64_t wnext = 0, wend = 0;     // warp-uniform: range of sorted entries the warp still owns
    bool exhausted = false;
    // per-lane read state
    bool active = false;
    uint64_t cur = 0;                 // byte offset of the current 16-byte word
    uint64_t end = 0;                 // byte offset one past the read
    uint64_t nbrk = ~0ull;            // next reformat.sh break position
    uint32_t lo = 0;                  // first valid byte of the current word (non-zero only in a read's first word)
    uint32_t Cc = 0, Vc = 0;          // carried codes / validity of the previous K-1 bases
    uint4 w = make_uint4(0, 0, 0, 0);

    for (;;) {
        const uint32_t need = __ballot_sync(0xffffffffu, !active);
        if (need) {
            if (wnext >= wend && !exhausted) {
                wnext = unit_at;
                wend = wnext + 32 < seg_e ? wnext + 32 : seg_e;
                unit_at += unit_stride;
                if (wnext >= seg_e) { exhausted = true; wnext = wend = seg_e; }
            }
            if (!active) {
                const uint64_t idx = wnext + __popc(need & ((1u << lane) - 1));
                if (idx < wend) {
                    const uint64_t ent = sorted[idx];
                    const uint64_t start = ent >> kEntryLenBits;
                    const uint64_t len = ent & kEntryLenMask;
                    end = start + len;
                    cur = start & ~15ull;
                    lo = (uint32_t)(start & 15ull);
                    nbrk = (breaklen > 0 && len > (uint64_t)breaklen) ? start + (uint64_t)breaklen : ~0ull;
                    Cc = 0;
                    Vc = 0;
                    w = __ldg(text16 + (cur >> 4));
                    active = true;
                }
            }
            const uint64_t adv = wnext + __popc(need);
            wnext = adv < wend ? adv : wend;
            if (exhausted && __ballot_sync(0xffffffffu, active) == 0) break;
        }
        if (!active) continue;

        // ---- one 16-byte word of this lane's read
        const uint64_t rem = end - cur;                       // > lo by construction
        const bool last = rem <= 16;
        uint4 wn = w;
        if (!last) wn = __ldg(text16 + (cur >> 4) + 1);       // next word, issued before the arithmetic below
        const uint32_t hi = last ? (uint32_t)rem : 16u;
        const uint32_t range = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
        const uint32_t V = (acgt4(w.x) | (acgt4(w.y) << 4) | (acgt4(w.z) << 8) | (acgt4(w.w) << 12)) & range;
        const uint32_t P = pack4(w.x) | (pack4(w.y) << 8) | (pack4(w.z) << 16) | (pack4(w.w) << 24);
        const uint64_t W = (uint64_t)Cc | ((uint64_t)P << (2 * KM1));
        uint32_t VW = Vc | (V << KM1);
        uint32_t E;
        if (nbrk < cur + 16) {
            // a break point falls in this word: no window may span it (windows wholly before or wholly after)
            const uint32_t b = (uint32_t)(nbrk - cur);         // 0..15: first base of the new piece
            const uint32_t below = (1u << (b + KM1)) - 1u;
            E = runs_of_k<K>(VW & below);
            VW &= ~below;                                      // bases before the cut are dead for later windows too
            E |= runs_of_k<K>(VW);
            nbrk += (uint64_t)breaklen;
            if (nbrk >= end) nbrk = ~0ull;
        } else {
            E = runs_of_k<K>(VW);
        }
        const uint32_t Wl = (uint32_t)W, Wh = (uint32_t)(W >> 32);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t idx = __funnelshift_r(Wl, Wh, 2 * j) & KMASK;
            if (SMEM) {
                atomicAdd(&s_hist[(E >> j) & 1u ? idx : trash], 1u);
            } else {
                if ((E >> j) & 1u) atomicAdd(gh + idx, 1ull);
            }
        }
        Cc = Wh;
        Vc = VW >> 16;
        lo = 0;
        if (last) {
            active = false;
        } else {
            cur += 16;
            w = wn;
        }
    }

    if (SMEM) {
        __syncthreads();
        uint32_t* slab = slabs + (size_t)blockIdx.x * NK;
        for (uint32_t i = tid; i < NK; i += kCountThreads) slab[i] = s_hist[i];
    }
}

// K3: per-segment histograms (uint64) = sum of the slabs of the CTAs that served the segment.
// grid covers kMaxLevels * 4^k bins; segments beyond the ladder are written as zero.
__global__ void __launch_bounds__(256)
reduce_slabs_kernel(const uint32_t* __restrict__ slabs, const Plan* __restrict__ plan, uint32_t nk,
                    unsigned long long* __restrict__ seg_hist)
{
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (uint64_t)kMaxLevels * nk) return;
    const uint32_t s = (uint32_t)(g / nk), i = (uint32_t)(g % nk);
    unsigned long long sum = 0;
    const uint32_t c0 = plan->seg_cta_begin[s], c1 = plan->seg_cta_begin[s + 1];
    for (uint32_t c = c0; c < c1; ++c) sum += slabs[(size_t)c * nk + i];
    seg_hist[g] = sum;
}

__global__ void __launch_bounds__(256)
zero_u64_kernel(unsigned long long* __restrict__ p, uint64_t n)
{
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) p[g] = 0;
}

}  // namespace vk
