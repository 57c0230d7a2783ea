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

// 8 validity flags of two classified words in byte 3: bits 24..27 = word a (bytes 0..3), bits 28..31 = word b.
// za / zb carry 0x80 in every valid byte; one multiply gathers both nibbles (no two partial products share a bit).
__device__ __forceinline__ uint32_t gather8(uint32_t za, uint32_t zb) { return (zb | (za >> 4)) * 0x00204081u; }

struct Cls4z {
    uint32_t packed_hi;   // byte 3 = 4 packed codes
    uint32_t z;           // 0x80 in every byte that is one of ACGTacgt
};
__device__ __forceinline__ Cls4z classify4z(uint32_t x)
{
    const uint32_t y = (x >> 1) & 0x03030303u;
    const uint32_t t = (x >> 2) & ~(x >> 1) & 0x01010101u;          // 1 where the code is T (10b)
    uint32_t e0;                                                      // 'A' + 2*code: bits 1,2 of 'A' are clear, so OR
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(e0) : "r"(x), "r"(0x06060606u), "r"(0x41414141u));
    const uint32_t e = t * 15u + e0;                                  // expected upper-case letter per byte
    const uint32_t d = (x & 0xDFDFDFDFu) ^ e;                         // 0 in a byte <=> valid
    Cls4z c;
    c.z = ~(((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
    c.packed_hi = y * 0x01041040u;
    return c;
}

template <int K>
__device__ __forceinline__ uint64_t runs_of_k64(uint64_t m)
{
    uint64_t r = m;
    int len = 1;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        if (len * 2 <= K) { r &= r >> len; len *= 2; }
    }
    if (len < K) r &= r >> (K - len);
    return r;
}

// 16 increments: window W4 holds K-1 carried codes then 16 new ones, pre-multiplied by 4 (byte offsets);
// bit j of E = the k-mer ending at new base j is to be counted.
#ifndef VK_EMIT_TRASH
#define VK_EMIT_TRASH 1      // 1: lanes without a k-mer increment a per-lane trash word.  ptxas cannot predicate
                             // ATOMS.POPC.INC (it branches around it, BSSY/BRA/BSYNC per increment), so 0 is slower
#endif
template <int K, bool SMEM>
__device__ __forceinline__ void emit16(const uint64_t W4, const uint32_t E, const uint32_t hist_addr,
                                       const uint32_t trash_addr, unsigned long long* const gh)
{
    constexpr uint32_t KMASK = (1u << (2 * K)) - 1;
    constexpr uint32_t fmask = KMASK << 2;
    const uint32_t Wl = (uint32_t)W4, Wh = (uint32_t)(W4 >> 32);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint32_t sh = __funnelshift_r(Wl, Wh, 2 * j);
        if (SMEM) {
#if VK_EMIT_TRASH
            smem_inc((E >> j) & 1u ? and_or(sh, fmask, hist_addr) : trash_addr);
#else
            if ((E >> j) & 1u) smem_inc(and_or(sh, fmask, hist_addr));
#endif
        } else {
            if ((E >> j) & 1u) atomicAdd(gh + ((sh & fmask) >> 2), 1ull);
        }
    }
}

// One unit = 32 sorted reads of the CTA's segment, one per lane, with the chunk numbering of the unit.
struct Unit {
    uint64_t ent;      // start << 24 | len   (0: no read in this lane)
    uint32_t excl;     // chunks of the unit's reads in lower lanes
    uint32_t total;    // chunks in the unit (warp-uniform)
    uint32_t nz;       // lanes that hold a read (warp-uniform; reads are a prefix of the lanes)
};
__device__ __forceinline__ Unit make_unit(uint64_t ent, uint32_t lane)
{
    constexpr uint32_t FULL = 0xffffffffu;
    Unit u;
    u.ent = ent;
    const uint32_t len = (uint32_t)(ent & kEntryLenMask);
    const uint32_t lo16 = (uint32_t)(ent >> kEntryLenBits) & 15u;
    const uint32_t n = len ? (lo16 + len + 31u) >> 5 : 0u;           // 32-byte chunks from the 16-byte word of the first base
    uint32_t incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, incl, d);
        if (lane >= (uint32_t)d) incl += t;
    }
    u.excl = incl - n;
    u.total = __shfl_sync(FULL, incl, 31);
    u.nz = __popc(__ballot_sync(FULL, n != 0));
    return u;
}

// What one lane needs to count its chunk: fetched one iteration ahead of its use.
struct Chunk {
    uint4 wa, wb;      // the 32 text bytes
    uint32_t range;    // valid bytes of the chunk (bit mask); 0 for an idle lane
    uint32_t j;        // chunk number inside its read (0: no bases to the left)
    uint32_t rlen;     // length of the read (break points)
    int32_t q0;        // read position of byte 0 of the chunk (negative in a read's first chunk)
};

// SMEM = true : histogram of the CTA's segment in shared memory (k <= 7), flushed to slabs[blockIdx.x]
// SMEM = false: increments go straight to the global (L2-resident) segment histogram (k = 8, 9)
//
// Work distribution ("flat lanes").  A warp takes the segment's sorted reads in units of 32 (one entry per lane)
// from the segment's global counter.  Every read is cut into 32-byte chunks that start at a 16-byte boundary of
// the text; the chunks of successive units form one stream, and in every iteration lane l works on chunk
// pos + l of that stream, whichever read it belongs to.  The owner of a chunk is found without a search: the lanes
// OR together one bit per read that starts inside the 32-chunk window (REDUX), and a lane's owner is the read of
// the highest such bit at or below it (or the read that owned the end of the previous window).  All 32 lanes
// classify and emit in every iteration (except at the very end of the segment), no lane waits for a longer read
// of a neighbour, and there is no divergent control flow in the loop: ranges (read start / end inside the chunk),
// invalid bytes and reformat.sh break points are bit masks.  The K-1 bases that precede a chunk come from the
// lane to the left (one shuffle of the packed tail; lane 0 keeps the tail of lane 31 of the previous iteration).
// The text of iteration i+1 is requested before iteration i is counted.
template <int K, bool SMEM>
__global__ void __launch_bounds__(kCountThreads)
count_kernel(const uint4* __restrict__ text16, const uint64_t* __restrict__ sorted, Plan* __restrict__ plan,
             uint32_t* __restrict__ slabs, unsigned long long* __restrict__ seg_hist, int breaklen)
{
    constexpr uint32_t NK = 1u << (2 * K);
    constexpr int KM1 = K - 1;
    constexpr uint32_t FULL = 0xffffffffu;
    extern __shared__ uint32_t s_raw[];           // SMEM: [pad to a 64 KiB shared address][NK bins][32 trash words]
    const uint32_t tid = threadIdx.x, lane = tid & 31;

    // ---- which segment does this CTA serve?
    int seg = -1;
    {
        const uint32_t b0 = plan->seg_cta_begin[lane], b1 = plan->seg_cta_begin[lane + 1];
        const uint32_t b2 = plan->seg_cta_begin[lane + 32], b3 = plan->seg_cta_begin[lane + 33];
        const uint32_t m0 = __ballot_sync(FULL, blockIdx.x >= b0 && blockIdx.x < b1);
        const uint32_t m1 = __ballot_sync(FULL, blockIdx.x >= b2 && blockIdx.x < b3);
        if (m0) seg = __ffs(m0) - 1;
        else if (m1) seg = 32 + __ffs(m1) - 1;
    }
    if (seg < 0) return;
    const uint64_t* const seg_sorted = sorted + plan->seg_begin[seg];
    // reads of this segment (< 2^32); never beyond the region (after a bucket overflow the step is repeated)
    const uint32_t seg_len = (uint32_t)(plan->seg_reads[seg] < plan->seg_cap[seg] ? plan->seg_reads[seg] : plan->seg_cap[seg]);
    unsigned long long* const gh = SMEM ? nullptr : seg_hist + (size_t)seg * NK;

    const uint32_t raw_addr = (uint32_t)__cvta_generic_to_shared(s_raw);
    const uint32_t hist_addr = SMEM ? (raw_addr + 0xFFFFu) & ~0xFFFFu : 0u;
    uint32_t* const s_hist = s_raw + ((hist_addr - raw_addr) >> 2);
    const uint32_t trash_addr = hist_addr + (NK + lane) * 4u;
    if (SMEM) {
        for (uint32_t i = tid; i < NK + 32; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
    }

    unsigned long long* const seg_counter = &plan->seg_next[seg];
    // units: A is being consumed, B follows it in the chunk stream, C is on its way from memory
    auto claim = [&]() -> uint64_t {
        unsigned long long r0 = 0;
        if (lane == 0) r0 = atomicAdd(seg_counter, 32ull);
        const uint32_t base = (uint32_t)__shfl_sync(FULL, r0, 0);
        return (base < seg_len && base + lane < seg_len) ? seg_sorted[base + lane] : 0ull;
    };
    Unit A = make_unit(claim(), lane);
    Unit B = make_unit(claim(), lane);
    uint64_t entC = claim();
    uint32_t pos = 0;                  // stream position of lane 0, in A's chunk numbering
    uint32_t own = 0;                  // owner (0..31: lane of A, 32..63: lane of B) and chunk number of the last
    uint32_t own_j = 0;                //   chunk of the previous window
    const bool has_break = breaklen > 0;

    auto fetch = [&]() -> Chunk {
        // rotate while A is used up
        while (pos >= A.total && A.nz != 0) {
            pos -= A.total;
            own -= 32;                 // an owner in B keeps its lane
            A = B;
            B = make_unit(entC, lane);
            entC = claim();
        }
        Chunk c;
        const uint32_t f = pos + lane;
        const uint32_t endAB = A.total + B.total;
        const bool act = f < endAB;
        // one bit per read whose first chunk lies in the window [pos, pos + 32)
        const uint32_t sa = A.excl - pos, sb = A.total + B.excl - pos;
        const bool hasA = (uint32_t)(A.ent & kEntryLenMask) != 0, hasB = (uint32_t)(B.ent & kEntryLenMask) != 0;
        const uint32_t bits = ((hasA && sa < 32u) ? 1u << sa : 0u) | ((hasB && sb < 32u) ? 1u << sb : 0u);
        const uint32_t marker = __reduce_or_sync(FULL, bits);
        const uint32_t mle = marker & (0xffffffffu >> (31u - lane));
        // reads that start in the window are numbered in lane order: A's lanes first, then B's
        const uint32_t firstA = __popc(__ballot_sync(FULL, hasA && A.excl < pos));      // reads of A that started earlier
        uint32_t o, j;
        if (mle == 0) { o = own; j = own_j + 1 + lane; }
        else {
            const uint32_t hb = 31u - __clz(mle);
            const uint32_t ord = firstA + __popc(mle) - 1;          // index among the reads of A then B
            o = ord < A.nz ? ord : ord - A.nz + 32;
            j = lane - hb;
        }
        own = __shfl_sync(FULL, o, 31);
        own_j = __shfl_sync(FULL, j, 31);
        const uint64_t ea = __shfl_sync(FULL, A.ent, (int)(o & 31u));
        uint64_t e = ea;
        if (pos + 32u > A.total) {                                  // warp-uniform: the window reaches into B
            const uint64_t eb = __shfl_sync(FULL, B.ent, (int)(o & 31u));
            if (o >= 32u) e = eb;
        }
        const uint64_t rstart = e >> kEntryLenBits;
        const uint32_t rlen = (uint32_t)(e & kEntryLenMask);
        const uint32_t rlo = (uint32_t)rstart & 15u;
        const uint4* const ptr = text16 + (rstart >> 4) + 2ull * j;
        const uint32_t lo = j == 0 ? rlo : 0u;
        const uint32_t endrel = rlo + rlen - 32u * j;               // > 0 for an active lane
        const uint32_t hi = endrel < 32u ? endrel : 32u;
        c.wa = make_uint4(0, 0, 0, 0);
        c.wb = c.wa;
        if (act) {
            c.wa = __ldg(ptr);
            if (hi > 16u) c.wb = __ldg(ptr + 1);
        }
        c.range = act ? (0xffffffffu >> (32u - hi)) & (0xffffffffu << lo) : 0u;
        c.j = j;
        c.rlen = rlen;
        c.q0 = (int32_t)(32u * j) - (int32_t)rlo;
        pos += 32;
        return c;
    };

    uint32_t carry = 0;                                             // tail of lane 31 of the previous iteration
    Chunk cur = fetch();
    while (__ballot_sync(FULL, cur.range != 0) != 0) {
        const Chunk nxt = fetch();

        const Cls4z c0 = classify4z(cur.wa.x), c1 = classify4z(cur.wa.y), c2 = classify4z(cur.wa.z), c3 = classify4z(cur.wa.w);
        const Cls4z c4 = classify4z(cur.wb.x), c5 = classify4z(cur.wb.y), c6 = classify4z(cur.wb.z), c7 = classify4z(cur.wb.w);
        const uint32_t v01 = gather8(c0.z, c1.z), v23 = gather8(c2.z, c3.z);
        const uint32_t v45 = gather8(c4.z, c5.z), v67 = gather8(c6.z, c7.z);
        const uint32_t V = __byte_perm(__byte_perm(v01, v23, 0x0073), __byte_perm(v45, v67, 0x0073), 0x5410) & cur.range;
        const uint32_t Plo = __byte_perm(__byte_perm(c0.packed_hi, c1.packed_hi, 0x0073),
                                         __byte_perm(c2.packed_hi, c3.packed_hi, 0x0073), 0x5410);
        const uint32_t Phi = __byte_perm(__byte_perm(c4.packed_hi, c5.packed_hi, 0x0073),
                                         __byte_perm(c6.packed_hi, c7.packed_hi, 0x0073), 0x5410);

        // ---- the K-1 bases before this chunk: from the lane to the left when it holds the same read
        const uint32_t tail = (Phi >> (32 - 2 * KM1)) | ((V >> (32 - KM1)) << 16);
        uint32_t hist = __shfl_up_sync(FULL, tail, 1);
        if (lane == 0) hist = carry;
        if (cur.j == 0) hist = 0;
        carry = __shfl_sync(FULL, tail, 31);
        const uint32_t Cc = hist & 0xFFFFu, Vc = hist >> 16;

        const uint64_t VW = (uint64_t)Vc | ((uint64_t)V << KM1);    // bit i <-> base i - (K-1) of the chunk
        uint32_t E = (uint32_t)runs_of_k64<K>(VW);
        if (has_break && __ballot_sync(FULL, cur.rlen > (uint32_t)breaklen) != 0) {
            // reformat.sh breaklength: no k-mer may span a multiple of breaklen counted from the read's first base.
            // byte b of the chunk is base q0 + b of the read; a window ending at base q spans the cut c
            // (c = m * breaklen, 1 <= m, c < rlen) iff q - (K-1) < c <= q, i.e. q in [c, c + K - 2].
            uint32_t dead = 0;
            if (cur.rlen > (uint32_t)breaklen && cur.range != 0) {
                const int32_t q0 = cur.q0;
                int32_t c = (q0 > 0 ? q0 / breaklen : 0) * breaklen;
                if (c < breaklen) c = breaklen;
                for (; c - q0 < 32 && c < (int32_t)cur.rlen; c += breaklen) {
                    const int32_t b = c - q0;                       // chunk byte that starts the new piece
                    if (b > -KM1) {
                        const uint32_t run = (1u << KM1) - 1u;     // K-1 window ends: b .. b+K-2
                        dead |= b >= 0 ? run << b : run >> (-b);
                    }
                }
            }
            E &= ~dead;
        }
        const uint64_t Wa = ((uint64_t)Cc | ((uint64_t)Plo << (2 * KM1))) << 2;
        const uint64_t Wb = ((uint64_t)(Plo >> (32 - 2 * KM1)) | ((uint64_t)Phi << (2 * KM1))) << 2;
        emit16<K, SMEM>(Wa, E & 0xFFFFu, hist_addr, trash_addr, gh);
        emit16<K, SMEM>(Wb, E >> 16, hist_addr, trash_addr, gh);
        cur = nxt;
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
