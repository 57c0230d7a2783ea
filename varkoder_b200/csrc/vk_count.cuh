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

// SIMD-in-register classification of 4 text bytes (one 32-bit word x).
//   y     = per byte the 2-bit code (ascii >> 1) & 3                      (A0 C1 T2 G3)
//   pack  : (y * 0x01041040) puts the four codes, densely packed, into byte 3 (partial products never overlap)
//   valid : a byte is one of ACGTacgt iff it equals the letter rebuilt from its own code:
//           letter = 'A' + 2*code + 15*[code == T]   ->  A 0x41, C 0x43, G 0x47, T 0x54;  bit 5 (case) is ignored.
//           The rebuild runs on the FMA pipe (IMAD), which the rest of the loop leaves idle.
struct Cls4 {
    uint32_t packed_hi;   // byte 3 = 4 packed codes
    uint32_t valid_hi;    // bits 28..31 = validity of bytes 0..3
};
__device__ __forceinline__ Cls4 classify4(uint32_t x)
{
    const uint32_t y = (x >> 1) & 0x03030303u;
    const uint32_t t = (y >> 1) & ~y & 0x01010101u;                  // 1 where the code is T
    const uint32_t e = t * 15u + (y * 2u + 0x41414141u);            // expected upper-case letter per byte
    const uint32_t d = (x & 0xDFDFDFDFu) ^ e;                        // 0 in a byte <=> valid
    const uint32_t z = ~(((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;   // 0x80 in every zero byte of d
    Cls4 c;
    c.packed_hi = y * 0x01041040u;
    c.valid_hi = z * 0x00204081u;                                    // flags gathered into bits 28..31
    return c;
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

// (a & b) | c in one LOP3 (c must not share set bits with the field selected by b)
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ void smem_inc(uint32_t shared_addr)
{
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(shared_addr) : "memory");      // SASS: ATOMS.POPC.INC
}

// One 16-byte word of a read: classify, extend the code / validity windows, emit up to 16 increments.
//   lo, hi : valid byte range [lo, hi) of this word inside the read;  brk : bytes from the start of this word to
//   the next reformat.sh break (>= 16: none here).  Cc / Vc carry the last K-1 codes / validity bits.
//   hist_addr : SHARED-WINDOW address of bin 0, a multiple of 64 KiB, so that "mask the offset" and "add the base"
//   are a single LOP3; trash_addr : this lane's trash word (absolute shared address).
template <int K, bool SMEM>
__device__ __forceinline__ void step16(const uint4 w, const uint32_t lo, const uint32_t hi, int32_t& brk,
                                       const int32_t rem, const int breaklen, uint32_t& Cc, uint32_t& Vc,
                                       const uint32_t hist_addr, const uint32_t trash_addr, unsigned long long* const gh)
{
    constexpr uint32_t NK = 1u << (2 * K);
    constexpr uint32_t KMASK = NK - 1;
    constexpr int KM1 = K - 1;
    const uint32_t range = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
    const Cls4 c0 = classify4(w.x), c1 = classify4(w.y), c2 = classify4(w.z), c3 = classify4(w.w);
    const uint32_t V = ((c0.valid_hi >> 28) | ((c1.valid_hi >> 24) & 0xF0u) | ((c2.valid_hi >> 20) & 0xF00u) |
                        ((c3.valid_hi >> 16) & 0xF000u)) & range;
    const uint32_t P = __byte_perm(__byte_perm(c0.packed_hi, c1.packed_hi, 0x0073),
                                   __byte_perm(c2.packed_hi, c3.packed_hi, 0x0073), 0x5410);
    // window of codes, pre-multiplied by 4 (byte offsets into the histogram): carried codes at bits [2, 2+2(K-1))
    const uint64_t W4 = ((uint64_t)Cc | ((uint64_t)P << (2 * KM1))) << 2;
    uint32_t VW = Vc | (V << KM1);
    uint32_t E;
    if (brk < 16) {
        // a break point falls in this word: no window may span it (windows wholly before or wholly after)
        const uint32_t below = (1u << ((uint32_t)brk + KM1)) - 1u;       // brk = first base of the new piece
        E = runs_of_k<K>(VW & below);
        VW &= ~below;                                      // bases before the cut are dead for later windows too
        E |= runs_of_k<K>(VW);
        brk += breaklen;
        if (brk >= rem) brk = 0x7fffffff;
    } else {
        E = runs_of_k<K>(VW);
    }
    const uint32_t Wl = (uint32_t)W4, Wh = (uint32_t)(W4 >> 32);
    const uint32_t fmask = KMASK << 2;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint32_t sh = __funnelshift_r(Wl, Wh, 2 * j);
        if (SMEM) {
            smem_inc((E >> j) & 1u ? and_or(sh, fmask, hist_addr) : trash_addr);
        } else {
            if ((E >> j) & 1u) atomicAdd(gh + ((sh & fmask) >> 2), 1ull);
        }
    }
    Cc = (uint32_t)(W4 >> 34);                            // the last K-1 codes of this word
    Vc = VW >> 16;
}

// SMEM = true : histogram of the CTA's segment in shared memory (k <= 7), flushed to slabs[blockIdx.x]
// SMEM = false: increments go straight to the global (L2-resident) segment histogram (k = 8, 9)
template <int K, bool SMEM>
__global__ void __launch_bounds__(kCountThreads)
count_kernel(const uint4* __restrict__ text16, const uint64_t* __restrict__ sorted, Plan* __restrict__ plan,
             uint32_t* __restrict__ slabs, unsigned long long* __restrict__ seg_hist, int breaklen)
{
    constexpr uint32_t NK = 1u << (2 * K);
    extern __shared__ uint32_t s_raw[];           // SMEM: [pad to a 64 KiB shared address][NK bins][32 trash words]
    const uint32_t tid = threadIdx.x, lane = tid & 31;

    // ---- which segment does this CTA serve?
    int seg = -1;
    for (int s = 0; s < kMaxLevels; ++s)
        if (blockIdx.x >= plan->seg_cta_begin[s] && blockIdx.x < plan->seg_cta_begin[s + 1]) seg = s;
    if (seg < 0) return;
    const uint64_t* const seg_sorted = sorted + plan->seg_begin[seg];
    const uint32_t seg_len = (uint32_t)plan->seg_reads[seg];            // reads of this segment (< 2^32)
    unsigned long long* const gh = SMEM ? nullptr : seg_hist + (size_t)seg * NK;

    const uint32_t raw_addr = (uint32_t)__cvta_generic_to_shared(s_raw);
    const uint32_t hist_addr = SMEM ? (raw_addr + 0xFFFFu) & ~0xFFFFu : 0u;
    uint32_t* const s_hist = s_raw + ((hist_addr - raw_addr) >> 2);
    const uint32_t trash_addr = hist_addr + (NK + lane) * 4u;
    if (SMEM) {
        for (uint32_t i = tid; i < NK + 32; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
    }

    // ---- read distribution.  A warp takes the segment's sorted reads in units of 32 from the segment's global
    // counter and always holds one more unit in reserve (entries already in registers), so neither the atomic nor
    // the entry load is ever waited for; inside a warp a lane that is on the last chunk of its read claims the
    // unit's next read and requests that read's first chunk in the slot a continuing read uses for its next chunk:
    // every global load is issued one loop iteration before its data is needed.  Warps and CTAs that run faster
    // simply take more units (a static split left 37 % of warp time at the final barrier; claiming on demand left
    // 29 % of the stall samples on the first-chunk load: profiles/r01_notes.md).
    unsigned long long* const seg_counter = &plan->seg_next[seg];
    uint32_t ubase = 0, unext = 0, uend = 0;          // current unit: entries [ubase, uend), next unclaimed = unext
    uint64_t uent = 0;                                // lane i holds entry ubase + i
    uint32_t rbase;                                   // reserve unit
    {
        unsigned long long r0 = 0;
        if (lane == 0) r0 = atomicAdd(seg_counter, 32ull);
        rbase = (uint32_t)__shfl_sync(0xffffffffu, r0, 0);
    }
    uint64_t rent = (rbase < seg_len && rbase + lane < seg_len) ? seg_sorted[rbase + lane] : 0ull;
    bool exhausted = false;

    // per-lane read state (32-bit, relative to the current 32-byte chunk: the text itself may be up to 2^40 bytes).
    // A lane consumes its read in aligned 32-byte chunks = one DRAM/L2 sector per step.
    bool active = false;
    const uint4* ptr = text16;        // current chunk (two 16-byte words)
    int32_t rem = 0;                  // bytes from the start of the current chunk to the end of the read
    int32_t brk = 0x7fffffff;         // bytes from the start of the current chunk to the next reformat.sh break
    uint32_t lo = 0;                  // first valid byte of the current chunk (non-zero only in a read's first chunk)
    uint32_t Cc = 0, Vc = 0;          // carried codes / validity of the previous K-1 bases
    uint4 wa = make_uint4(0, 0, 0, 0), wb = wa;

    for (;;) {
        // ---- who needs a (next) read?  idle lanes, and lanes whose current chunk is their read's last
        const bool last = active && rem <= 32;
        const bool want = !active || last;
        const uint32_t wmask = __ballot_sync(0xffffffffu, want);
        if (unext >= uend && !exhausted && wmask) {                   // warp-uniform: bring in the reserve unit
            ubase = unext = rbase;
            uend = rbase + 32 < seg_len ? rbase + 32 : seg_len;
            uent = rent;
            if (rbase >= seg_len) { exhausted = true; ubase = unext = uend = seg_len; }
            else {
                unsigned long long r1 = 0;
                if (lane == 0) r1 = atomicAdd(seg_counter, 32ull);
                rbase = (uint32_t)__shfl_sync(0xffffffffu, r1, 0);
                rent = (rbase < seg_len && rbase + lane < seg_len) ? seg_sorted[rbase + lane] : 0ull;
            }
        }
        const uint32_t my = unext + __popc(wmask & ((1u << lane) - 1));
        const bool got = want && my < uend;
        const uint64_t ent = __shfl_sync(0xffffffffu, uent, got ? (int)(my - ubase) : 0);
        {
            const uint32_t adv = unext + __popc(wmask);
            unext = adv < uend ? adv : uend;
        }
        // ---- issue this iteration's loads: next chunk of the same read, or first chunk of the claimed read
        uint4 na = wa, nb = wb;
        const uint4* nptr = ptr + 2;
        uint32_t nlo = 0, nlen = 0;
        if (got) {
            const uint64_t start = ent >> kEntryLenBits;
            nlen = (uint32_t)(ent & kEntryLenMask);
            nlo = (uint32_t)start & 31u;
            nptr = text16 + ((start >> 5) << 1);
            na = __ldg(nptr);
            if (nlo + nlen > 16) nb = __ldg(nptr + 1);
        } else if (active && !last) {
            na = __ldg(nptr);
            if (rem > 48) nb = __ldg(nptr + 1);
        }
        if (exhausted && __ballot_sync(0xffffffffu, active || got) == 0) break;

        // ---- one 32-byte chunk of this lane's read
        if (active) {
            if (lo < 16) {
                const uint32_t hi0 = rem < 16 ? (uint32_t)rem : 16u;
                step16<K, SMEM>(wa, lo, hi0, brk, rem, breaklen, Cc, Vc, hist_addr, trash_addr, gh);
            }
            if (rem > 16) {
                const uint32_t lo1 = lo > 16 ? lo - 16 : 0u;
                const uint32_t hi1 = rem < 32 ? (uint32_t)rem - 16u : 16u;
                int32_t brk1 = brk == 0x7fffffff ? brk : brk - 16;
                step16<K, SMEM>(wb, lo1, hi1, brk1, rem - 16, breaklen, Cc, Vc, hist_addr, trash_addr, gh);
                brk = brk1 == 0x7fffffff ? brk1 : brk1 + 16;
            }
        }
        // ---- advance
        if (got) {                                            // hand over to the claimed read
            ptr = nptr;
            lo = nlo;
            rem = (int32_t)(nlo + nlen);
            brk = (breaklen > 0 && nlen > (uint32_t)breaklen) ? (int32_t)(nlo + breaklen) : 0x7fffffff;
            Cc = 0;
            Vc = 0;
            active = true;
        } else if (active && !last) {
            ptr = nptr;
            lo = 0;
            rem -= 32;
            if (brk != 0x7fffffff) brk -= 32;
        } else {
            active = false;
        }
        wa = na;
        wb = nb;
    }

    if (SMEM) {
        __syncthreads();
        uint32_t* slab = slabs + (size_t)blockIdx.x * NK;
        for (uint32_t i = tid; i < NK; i += blockDim.x) slab[i] = s_hist[i];
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
