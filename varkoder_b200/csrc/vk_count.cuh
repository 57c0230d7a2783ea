// vk_count.cuh -- K2: k-mer counting of every ladder segment in one pass over the sequence lines.
//
// Stands in for `dsk -kmer-size k -abundance-min 1` run once per sub-sample file
// (varKoder/commands/image.py:771-796), restated per SURVEY.md section 8c:
//   D1 k-mers never span records (nor the <=500-base pieces reformat.sh breaklength=500 cuts, image.py:586)
//   D2 2-bit code (ascii >> 1) & 3;  D3 a window holding a non-ACGT byte is dropped;
//   D4/D5 canonical abundance = forward count of K + forward count of rc(K): the fold happens after
//   counting (vk_image.cuh), so the hot loop only builds a forward-strand histogram.
//
// Design (DESIGN.md "K2").  Reads arrive sorted by segment (vk_bucket.cuh); a CTA serves one segment and keeps
// that segment's histogram in shared memory, written to its private slab (plain stores; no global atomics).
//
// Work distribution ("flat lanes", struct ChunkStream).  A warp takes the segment's sorted reads in units of 32
// (one entry per lane) from the segment's global counter.  Every read is cut into 32-byte chunks that start at a
// 16-byte boundary of the text; the chunks of successive units form one stream, and in every iteration lane l
// works on chunk pos + l of that stream, whichever read it belongs to.  The owner of a chunk is found without a
// search: the lanes OR together one bit per read that starts inside the 32-chunk window (REDUX), and a lane's owner
// is the read of the highest such bit at or below it (or the read that owned the end of the previous window).  All
// 32 lanes classify and emit in every iteration (except at the very end of the segment), no lane waits for a longer
// read of a neighbour, and there is no divergent control flow in the loop: ranges (read start / end inside the
// chunk), invalid bytes and reformat.sh break points are bit masks.  The K-1 bases that precede a chunk come from
// the lane to the left (one shuffle of the packed tail; lane 0 keeps the tail of lane 31 of the previous
// iteration).  The text of iteration i+1 is requested before iteration i is counted.
//
// Per chunk: two LDG.128; the 32 bases are packed to 2 bits with SIMD-in-register arithmetic and joined to the
// previous K-1 codes in 64-bit windows, so each k-mer is one funnel shift + mask; validity is a bit mask whose
// K-long runs give the 32-bit "a k-mer ends here" mask E.
//
// Four ways to count:
//   kSmem32   k <= 7: 4^k x u32 bins in shared memory, one ATOMS.POPC.INC per base.  Lanes without a k-mer
//             increment one common trash word (POPC.INC folds lanes with the same address, so they add a single bank
//             access; per-lane trash words measured 142.5 us against 139.3 us): ptxas cannot predicate ATOMS.POPC.INC
//             (it branches around it) and a sparsely populated ATOMS costs as much as a full one.
//   kSmem16   k = 7, 8: 4^8 16-bit bins (two per 32-bit word, 128 KiB).  k = 8 counts every 8-mer; k = 7 counts
//             PAIRS: the 8-mer that ends at an odd chunk position stands for the two 7-mers it contains, so a base
//             pair costs one shared-memory operation instead of two (the kernel is bound by the shared-memory data
//             pipe: 3.5 wavefronts per 32 random banks, profiles/r01_notes.md); 7-mers whose pair partner is not
//             countable (read ends, N, break points) go to a 4^7 x u32 table.  16-bit bins are kept exact by
//             watching the values the atomics return and draining hot words to the slab (see count16_kernel).
//   pairs/9   k = 9 (count9h_kernel): the 2 x 4^8 canonical classes of the odd k as two such 16-bit tables, one per CTA
//             of a pair that streams the same reads; the class falls out of the middle base without a comparison.
//   kGlobal   k = 8, 9 behind VK_COUNT16=0: increments go straight to the global (L2-resident) segment histogram.
#pragma once
#include "vk_common.cuh"

namespace vk {

constexpr int kCountThreads = 1024;
enum { kSmem32 = 0, kGlobal = 1, kSmem16 = 2 };

template <int K>
__device__ __forceinline__ uint64_t runs_of_k64(uint64_t m)
{
    // bit j of the result = bits j .. j+K-1 of m are all set
    uint64_t r = m;
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
// 16 text bytes through the read-only path, asking L2 to fetch the whole 128-byte line on a miss.  A lane's chunk is 32
// bytes of a ~150-byte sequence line in a 400 MB text: with sector-sized (32-byte) fills the count kernel's DRAM
// accesses are scattered 32-byte pieces and the HBM delivers 2.7 TB/s of them however many are in flight; the other
// sectors of the line are wanted a moment later by the neighbouring lanes anyway.
__device__ __forceinline__ uint4 ldg_text(const uint4* p)
{
    uint4 v;
    asm volatile("ld.global.nc.L2::128B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void smem_inc(uint32_t shared_addr)
{
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(shared_addr) : "memory");      // SASS: ATOMS.POPC.INC
}
__device__ __forceinline__ void smem_add(uint32_t shared_addr, uint32_t v)
{
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(shared_addr), "r"(v) : "memory");      // SASS: ATOMS.ADD, nothing comes back
}
__device__ __forceinline__ uint32_t mad_hi_u32(uint32_t a, uint32_t b, uint32_t c)      // (a * b >> 32) + c, one IMAD.HI
{
    uint32_t d;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t smem_add_ret(uint32_t shared_addr, uint32_t v)
{
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(shared_addr), "r"(v) : "memory");
    return old;
}

// SIMD-in-register classification of 4 text bytes (one 32-bit word x).  The loop is bound by the ALU pipe (LOP3 / SHF /
// PRMT) and by the shared-memory data pipe while the FMA pipe idles, so everything that can be a multiply-add is one:
//   y     = x & 0x06060606: the 2-bit code (ascii >> 1) & 3 (A0 C1 T2 G3) still at bits 1-2 of each byte
//   pack  : y * 0x00820820 puts the four codes, densely packed, into byte 3 (no two partial products share a bit there)
//   T?    : y * 3 is 0, 6, 12, 18 for A, C, T, G -- only 12 has bit 3 set
//   valid : a byte is one of ACGTacgt iff, case bit cleared, it is the letter its own code names.  'A' 0x41, 'C' 0x43,
//           'G' 0x47 agree outside bits 1-2, so (x & 0xD9) ^ 0x41 is 0 for them (mask 0xD9 drops bits 1, 2 and the case
//           bit 5) and 0x11 for 'T' 0x54: d = ((x & 0xD9) ^ 0x41) ^ (T? ? 0x11 : 0) is 0 exactly for the eight letters.
//           Zero bytes of d by the carry-free test; bit 7 of d is bit 7 of x, so it is taken from x directly.
//   6 ALU-pipe + 4 FMA-pipe instructions per word (the first version of this kernel: 9 + 2).
struct Cls4z {
    uint32_t packed_hi;   // byte 3 = 4 packed codes
    uint32_t z;           // 0x80 in every byte that is one of ACGTacgt
};
__device__ __forceinline__ uint32_t mad_lo_op(uint32_t a, uint32_t b, uint32_t c)      // a * b + c as IMAD even when b is 1
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ Cls4z classify4z(uint32_t x, uint32_t one)
{
    const uint32_t y = x & 0x06060606u;
    const uint32_t t8 = (y * 3u) & 0x08080808u;
    const uint32_t t11 = (t8 >> 3) * 0x11u;
    uint32_t d1, dm;
    asm("lop3.b32 %0, %1, %2, %3, 0x6A;" : "=r"(d1) : "r"(x), "r"(0xD9D9D9D9u), "r"(0x41414141u));      // (x & M) ^ C
    asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(dm) : "r"(d1), "r"(t11), "r"(0x7F7F7F7Fu));            // (d1 ^ t11) & 0x7F..
    const uint32_t sv = mad_lo_op(dm, one, 0x7F7F7F7Fu);                                                 // + 0x7F.. on the FMA pipe
    Cls4z c;
    asm("lop3.b32 %0, %1, %2, %3, 0x02;" : "=r"(c.z) : "r"(sv), "r"(x), "r"(0x80808080u));             // ~(sv | x) & 0x80..
    c.packed_hi = y * 0x00820820u;
    return c;
}
// 8 validity flags of two classified words in byte 3: bits 24..27 = word a (bytes 0..3), bits 28..31 = word b.
// za / zb carry 0x80 in every valid byte; one multiply gathers both nibbles (no two partial products share a bit).
__device__ __forceinline__ uint32_t gather8(uint32_t za, uint32_t zb) { return (zb | (za >> 4)) * 0x00204081u; }

// ---------------------------------------------------------------------------------------- chunk stream
// One unit = 32 sorted reads of the CTA's segment, one per lane, with the chunk numbering of the unit.
struct Unit {
    uint64_t ent;      // start << 24 | len   (0: no read in this lane)
    uint32_t excl;     // chunks of the unit's reads in lower lanes
    uint32_t total;    // chunks in the unit (warp-uniform)
    uint32_t nz;       // lanes that hold a read (warp-uniform; reads are a prefix of the lanes)
};
// ALIGN: text alignment of a read's first chunk -- 16 when the chunk is two 16-byte text loads, 32 when it is one
// 32-byte block of the packed arrays
template <uint32_t ALIGN>
__device__ __forceinline__ Unit make_unit(uint64_t ent, uint32_t lane)
{
    constexpr uint32_t FULL = 0xffffffffu;
    Unit u;
    u.ent = ent;
    const uint32_t len = (uint32_t)(ent & kEntryLenMask);
    const uint32_t lo16 = (uint32_t)(ent >> kEntryLenBits) & (ALIGN - 1u);
    const uint32_t n = len ? (lo16 + len + 31u) >> 5 : 0u;           // 32-byte chunks from the aligned word of the first base
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
    uint4 wa, wb;      // the 32 text bytes (PACKED: wa.x, wa.y = their 2-bit codes, wa.z = their validity bits)
    uint32_t range;    // valid bytes of the chunk (bit mask); 0 for an idle lane
    uint32_t j;        // chunk number inside its read (0: no bases to the left)
    uint32_t rlen;     // length of the read (break points)
    int32_t q0;        // read position of byte 0 of the chunk (negative in a read's first chunk)
    uint32_t last;     // last lane of the window that holds a chunk (warp-uniform; 31 except where units run out)
};

// PACKED: the chunks come from the 2-bit codes + validity bits that parse_mask_kernel<true> wrote (vk_parse.cuh, one
// 8-byte and one 4-byte load per chunk, nothing to classify) instead of from the text (two 16-byte loads per chunk).
struct PackedSrc {
    const uint2* codes64;      // 32 bases per element
    const uint32_t* valid32;
};
template <bool PACKED>
struct ChunkStream {
    static constexpr uint32_t ALIGN = PACKED ? 32u : 16u;
    PackedSrc pk;
    const uint4* text16;
    const uint64_t* seg_sorted;
    unsigned long long* seg_counter;
    uint32_t seg_len, lane;
    uint32_t n_warps, last_base;   // warps serving the segment; counter value seen by the last claim
    uint64_t text_words;   // 16-byte words of the text buffer (VK_ASSERT only)
    Unit A, B;         // A is being consumed, B follows it in the chunk stream
    uint64_t entC;     // the unit after B, on its way from memory
    uint32_t pos;      // stream position of lane 0, in A's chunk numbering
    uint32_t own, own_j;   // owner (0..31: lane of A, 32..63: lane of B) and chunk number of the last chunk of the previous window

    // Guided self-scheduling: a unit is 32 reads while the segment has plenty left, 16 / 8 / 4 towards its end, so that
    // the warps of a segment finish within a few reads of each other (a full unit is ~5 iterations = ~16 us of a warp's
    // time when 32 warps share the SM; with fixed units that was the kernel's tail).  The chunk stream runs across
    // units, so small units do not leave lanes idle.
    __device__ __forceinline__ uint64_t claim()
    {
        const uint32_t left = last_base < seg_len ? seg_len - last_base : 0u;
        const uint32_t per_warp = left / n_warps;                    // reads still unclaimed per warp, roughly
        const uint32_t size = per_warp >= 96u ? 32u : per_warp >= 32u ? 16u : per_warp >= 8u ? 8u : 4u;
        unsigned long long r0 = 0;
        if (lane == 0) r0 = atomicAdd(seg_counter, (unsigned long long)size);
        const uint32_t base = (uint32_t)__shfl_sync(0xffffffffu, r0, 0);
        last_base = base;
        return (lane < size && base < seg_len && base + lane < seg_len) ? seg_sorted[base + lane] : 0ull;
    }
    // n_warps_seg: warps that serve this segment (all CTAs of it)
    __device__ __forceinline__ void init(const uint4* t, PackedSrc src, const uint64_t* ss, uint32_t sl, unsigned long long* sc,
                                         uint32_t ln, uint64_t n_bytes, uint32_t n_warps_seg)
    {
        text16 = t; pk = src; seg_sorted = ss; seg_len = sl; seg_counter = sc; lane = ln;
        text_words = (n_bytes + 15) >> 4;
        pos = 0; own = 0; own_j = 0;
        n_warps = n_warps_seg ? n_warps_seg : 1u;
        last_base = 0;
        if ((uint64_t)seg_len >= 192ull * n_warps) {
            // plenty of reads per warp: the first three units with ONE round trip to the counter and three independent
            // loads (three dependent claims cost ~5 us of start-up latency in every warp)
            unsigned long long r0 = 0;
            if (lane == 0) r0 = atomicAdd(seg_counter, 96ull);
            const uint32_t base = (uint32_t)__shfl_sync(0xffffffffu, r0, 0);
            last_base = base + 64;
            const uint64_t ea = (base < seg_len && base + lane < seg_len) ? seg_sorted[base + lane] : 0ull;
            const uint64_t eb = (base + 32 < seg_len && base + 32 + lane < seg_len) ? seg_sorted[base + 32 + lane] : 0ull;
            entC = (base + 64 < seg_len && base + 64 + lane < seg_len) ? seg_sorted[base + 64 + lane] : 0ull;
            A = make_unit<ALIGN>(ea, lane);
            B = make_unit<ALIGN>(eb, lane);
        } else {                                                        // small segment: unit by unit keeps the warps balanced
            A = make_unit<ALIGN>(claim(), lane);
            B = make_unit<ALIGN>(claim(), lane);
            entC = claim();
        }
    }
    __device__ __forceinline__ Chunk fetch()
    {
        constexpr uint32_t FULL = 0xffffffffu;
        while (pos >= A.total && A.nz != 0) {          // rotate while A is used up
            pos -= A.total;
            own -= 32;                                  // an owner in B keeps its lane
            A = B;
            B = make_unit<ALIGN>(entC, lane);
            entC = claim();
        }
        Chunk c;
        const uint32_t endAB = A.total + B.total;
        // chunks available to this window: 32 except where A and B together hold fewer (small units towards the end of
        // the segment); the window after this one starts right behind the last chunk taken
        const uint32_t nact = endAB > pos ? (endAB - pos < 32u ? endAB - pos : 32u) : 0u;
        const bool act = lane < nact;
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
        const int last = nact ? (int)nact - 1 : 31;
        own = __shfl_sync(FULL, o, last);
        own_j = __shfl_sync(FULL, j, last);
        uint64_t e = __shfl_sync(FULL, A.ent, (int)(o & 31u));
        if (pos + 32u > A.total) {                                  // warp-uniform: the window reaches into B
            const uint64_t eb = __shfl_sync(FULL, B.ent, (int)(o & 31u));
            if (o >= 32u) e = eb;
        }
        const uint64_t rstart = e >> kEntryLenBits;
        const uint32_t rlen = (uint32_t)(e & kEntryLenMask);
        const uint32_t rlo = (uint32_t)rstart & (ALIGN - 1u);
        const uint32_t lo = j == 0 ? rlo : 0u;
        const uint32_t endrel = rlo + rlen - 32u * j;               // > 0 for an active lane
        const uint32_t hi = endrel < 32u ? endrel : 32u;
        c.wa = make_uint4(0, 0, 0, 0);
        c.wb = c.wa;
        if (PACKED) {
            if (act) {
                const uint64_t blk = (rstart >> 5) + j;             // 32-byte block of the text = one element of each array
                VK_ASSERT(blk < ((text_words + 1) >> 1) && rlen != 0 && j < ((rlo + rlen + 31u) >> 5));
                const uint2 cd = __ldg(pk.codes64 + blk);
                c.wa.x = cd.x; c.wa.y = cd.y;
                c.wa.z = __ldg(pk.valid32 + blk);
            }
        } else {
            const uint4* const ptr = text16 + (rstart >> 4) + 2ull * j;
            if (act) {
                VK_ASSERT((uint64_t)(ptr - text16) + (hi > 16u ? 1 : 0) < text_words && rlen != 0 && j < ((rlo + rlen + 31u) >> 5));
                c.wa = ldg_text(ptr);
                if (hi > 16u) c.wb = ldg_text(ptr + 1);
            }
        }
        c.range = act ? (0xffffffffu >> (32u - hi)) & (0xffffffffu << lo) : 0u;
        c.j = j;
        c.rlen = rlen;
        c.q0 = (int32_t)(32u * j) - (int32_t)rlo;
        c.last = (uint32_t)last;
        pos += nact ? nact : 32u;
        return c;
    }
};

// Codes and countable positions of one chunk.
struct Decoded {
    uint32_t Plo, Phi;   // 2-bit codes of bases 0..15 / 16..31
    uint32_t Cc;         // codes of the K-1 bases to the left (first at bit 0)
    uint32_t E;          // bit b: the K-mer that ENDS at base b is to be counted
};
// one: the constant 1 in a register ptxas cannot see through (from a kernel argument), see classify4z
template <int K, bool PACKED = false>
__device__ __forceinline__ Decoded decode_chunk(const Chunk& cur, uint32_t& carry, uint32_t lane, int breaklen, uint32_t one)
{
    constexpr int KM1 = K - 1;
    constexpr uint32_t FULL = 0xffffffffu;
    Decoded d;
    uint32_t V;
    if (PACKED) {
        d.Plo = cur.wa.x;
        d.Phi = cur.wa.y;
        V = cur.wa.z & cur.range;
    } else {
        const Cls4z c0 = classify4z(cur.wa.x, one), c1 = classify4z(cur.wa.y, one), c2 = classify4z(cur.wa.z, one), c3 = classify4z(cur.wa.w, one);
        const Cls4z c4 = classify4z(cur.wb.x, one), c5 = classify4z(cur.wb.y, one), c6 = classify4z(cur.wb.z, one), c7 = classify4z(cur.wb.w, one);
        const uint32_t v01 = gather8(c0.z, c1.z), v23 = gather8(c2.z, c3.z);
        const uint32_t v45 = gather8(c4.z, c5.z), v67 = gather8(c6.z, c7.z);
        V = __byte_perm(__byte_perm(v01, v23, 0x0073), __byte_perm(v45, v67, 0x0073), 0x5410) & cur.range;
        d.Plo = __byte_perm(__byte_perm(c0.packed_hi, c1.packed_hi, 0x0073), __byte_perm(c2.packed_hi, c3.packed_hi, 0x0073), 0x5410);
        d.Phi = __byte_perm(__byte_perm(c4.packed_hi, c5.packed_hi, 0x0073), __byte_perm(c6.packed_hi, c7.packed_hi, 0x0073), 0x5410);
    }
    // ---- the K-1 bases before this chunk: from the lane to the left when it holds the same read
    const uint32_t tail = (d.Phi >> (32 - 2 * KM1)) | ((V >> (32 - KM1)) << 16);
    uint32_t hist = __shfl_up_sync(FULL, tail, 1);
    if (lane == 0) hist = carry;
    if (cur.j == 0) hist = 0;
    carry = __shfl_sync(FULL, tail, (int)cur.last);
    d.Cc = hist & 0xFFFFu;
    const uint32_t Vc = hist >> 16;
    const uint64_t VW = (uint64_t)Vc | ((uint64_t)V << KM1);        // bit i <-> base i - (K-1) of the chunk
    uint32_t E = (uint32_t)runs_of_k64<K>(VW);
    if (breaklen > 0 && __ballot_sync(FULL, cur.rlen > (uint32_t)breaklen) != 0) {
        // reformat.sh breaklength: no k-mer may span a multiple of breaklen counted from the read's first base.
        // byte b of the chunk is base q0 + b of the read; a window ending at base q spans the cut c
        // (c = m * breaklen, 1 <= m, c < rlen) iff q - (K-1) < c <= q, i.e. q in [c, c + K - 2].
        uint32_t dead = 0;
        if (cur.rlen > (uint32_t)breaklen && cur.range != 0) {
            const int32_t q0 = cur.q0;
            int32_t c = (q0 > 0 ? q0 / breaklen : 0) * breaklen;
            if (c < breaklen) c = breaklen;
            for (; c - q0 < 32 && c < (int32_t)cur.rlen; c += breaklen) {
                const int32_t b = c - q0;                           // chunk byte that starts the new piece
                if (b > -KM1) {
                    const uint32_t run = (1u << KM1) - 1u;         // K-1 window ends: b .. b+K-2
                    dead |= b >= 0 ? run << b : run >> (-b);
                }
            }
        }
        E &= ~dead;
    }
    d.E = E;
    return d;
}

// which segment does this CTA serve?  (-1: none)
__device__ __forceinline__ int cta_segment(const Plan* __restrict__ plan, uint32_t lane, uint32_t cta);
// CTAs are handed to the SMs in launch order and the plan puts the big segments first; the count kernels therefore
// number themselves BACKWARDS, so that the small last segments run on the first CTAs and the few CTAs launched beyond
// the SM count (vk_ctx::count_extra) join a big segment late, when a small one has finished: they take units off the
// segment's shared counter and level the tail.
__device__ __forceinline__ uint32_t logical_cta() { return gridDim.x - 1u - blockIdx.x; }
__device__ __forceinline__ int cta_segment(const Plan* __restrict__ plan, uint32_t lane) { return cta_segment(plan, lane, logical_cta()); }
__device__ __forceinline__ int cta_segment(const Plan* __restrict__ plan, uint32_t lane, const uint32_t cta)
{
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t b0 = plan->seg_cta_begin[lane], b1 = plan->seg_cta_begin[lane + 1];
    const uint32_t b2 = plan->seg_cta_begin[lane + 32], b3 = plan->seg_cta_begin[lane + 33];
    const uint32_t m0 = __ballot_sync(FULL, cta >= b0 && cta < b1);
    const uint32_t m1 = __ballot_sync(FULL, cta >= b2 && cta < b3);
    if (m0) return __ffs(m0) - 1;
    if (m1) return 32 + __ffs(m1) - 1;
    return -1;
}

// ------------------------------------------------------------------------------ kSmem32 / kGlobal
// 16 increments: window W4 holds K-1 carried codes then 16 new ones, pre-multiplied by 4 (byte offsets);
// bit j of E = the k-mer ending at new base j is to be counted.
template <int K, int MODE>
__device__ __forceinline__ void emit16(const uint64_t W4, const uint32_t E, const uint32_t hist_addr,
                                       const uint32_t trash_addr, unsigned long long* const gh)
{
    constexpr uint32_t KMASK = (1u << (2 * K)) - 1;
    constexpr uint32_t fmask = KMASK << 2;
    const uint32_t Wl = (uint32_t)W4, Wh = (uint32_t)(W4 >> 32);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint32_t sh = __funnelshift_r(Wl, Wh, 2 * j);
        if (MODE == kSmem32) {
            smem_inc((E >> j) & 1u ? and_or(sh, fmask, hist_addr) : trash_addr);
        } else {
            if ((E >> j) & 1u) atomicAdd(gh + ((sh & fmask) >> 2), 1ull);
        }
    }
}

// Is the sample one for the one-read-per-lane kernels (vk_countt.cuh, vk_countu.cuh)?  Reads of ONE length (raw Illumina
// reads: BASELINE configs 1, 2, 3, 5), short enough for a unit of 27 reads or more (13 text words each: 176 bases), none
// longer than the break length.  "One length" is judged by the mean: the longest counted read (scatter kernel) against
// bases / records of the framing pass -- a unit that holds a shorter read falls back to the queue for its last words, so
// a few stragglers (the odd last read of a file) cost nothing, while a sample whose mean is 0.4 % below its longest read
// may have one in every eighth unit and is left to the flat-lane kernel.  Evaluated on the device by both count kernels
// (exactly one of them counts): everything it needs is in the plan when the scatter kernel is done.
__device__ __forceinline__ bool countu_wanted(const Plan* __restrict__ plan, int breaklen, uint32_t policy)
{
    if (policy != 2u) return policy == 1u;       // 0: never, 1: always (tests), 2: by the sample
    const uint64_t hi = plan->len_max;
    return hi != 0u && hi <= 176u && plan->nsites_true * 256u >= hi * 255u * plan->n_reads &&
           (breaklen <= 0 || hi <= (uint64_t)breaklen);
}

template <int K, int MODE, bool PACKED>
__global__ void __launch_bounds__(kCountThreads)
count_kernel(const StepArgs* __restrict__ sa, PackedSrc pk, const uint64_t* __restrict__ sorted, Plan* __restrict__ plan,
             uint32_t* __restrict__ slabs, unsigned long long* __restrict__ seg_hist, uint32_t lanes_policy)
{
    pdl_wait();
    const uint4* __restrict__ text16 = reinterpret_cast<const uint4*>(sa->text);
    const int breaklen = sa->pa.p.breaklength;
    // lanes_policy 2: countt_kernel was launched in front of this kernel with the same rule -- exactly one of the two counts
    // lanes_policy 3: this kernel counts, and tells the host when the sample would have been one for countt_kernel (the
    // context then launches that kernel for its next sample, vk_capi.cu)
    if (K == 7 && MODE == kSmem32 && lanes_policy == 2u && countu_wanted(plan, breaklen, 2u)) return;
    if (K == 7 && MODE == kSmem32 && lanes_policy == 3u && blockIdx.x == 0 && threadIdx.x == 0 && countu_wanted(plan, breaklen, 2u))
        atomicOr(&plan->lanes_verdict, 1u);
    const uint32_t one = (uint32_t)(sa->n_bytes >> 62) + 1u;          // 1 (texts are shorter than 2^40), but not to ptxas
    static_assert(MODE == kSmem32 || MODE == kGlobal, "count16_kernel is the kSmem16 kernel");
    constexpr uint32_t NK = 1u << (2 * K);
    constexpr int KM1 = K - 1;
    constexpr uint32_t FULL = 0xffffffffu;
    extern __shared__ uint32_t s_raw[];           // kSmem32: [pad to a 64 KiB shared address][NK bins][trash word]
    const uint32_t tid = threadIdx.x, lane = tid & 31;

    const int seg = cta_segment(plan, lane);
    if (seg < 0) return;
    // reads of this segment (< 2^32); never beyond the region (after a bucket overflow the step is repeated)
    const uint32_t seg_len = (uint32_t)(plan->seg_reads[seg] < plan->seg_cap[seg] ? plan->seg_reads[seg] : plan->seg_cap[seg]);
    unsigned long long* const gh = MODE == kSmem32 ? nullptr : seg_hist + (size_t)seg * NK;

    // the histogram sits at a 64 KiB-aligned shared address so that "mask the k-mer, add the base" is one LOP3
    const uint32_t raw_addr = (uint32_t)__cvta_generic_to_shared(s_raw);
    const uint32_t hist_addr = MODE == kSmem32 ? (raw_addr + 0xFFFFu) & ~0xFFFFu : 0u;
    uint32_t* const s_hist = s_raw + ((hist_addr - raw_addr) >> 2);
    const uint32_t trash_addr = hist_addr + NK * 4u;
    if (MODE == kSmem32) {
        for (uint32_t i = tid; i < NK + 32; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
    }

    ChunkStream<PACKED> cs;
    cs.init(text16, pk, sorted + plan->seg_begin[seg], seg_len, &plan->seg_next[seg], lane, plan->n_bytes,
            (plan->seg_cta_begin[seg + 1] - plan->seg_cta_begin[seg]) * (blockDim.x >> 5));
    uint32_t carry = 0;                                             // tail of lane 31 of the previous iteration
    Chunk cur = cs.fetch();
    while (__ballot_sync(FULL, cur.range != 0) != 0) {
        const Chunk nxt = cs.fetch();
        const Decoded d = decode_chunk<K, PACKED>(cur, carry, lane, breaklen, one);
        const uint64_t Wa = ((uint64_t)d.Cc | ((uint64_t)d.Plo << (2 * KM1))) << 2;
        const uint64_t Wb = ((uint64_t)(d.Plo >> (32 - 2 * KM1)) | ((uint64_t)d.Phi << (2 * KM1))) << 2;
        emit16<K, MODE>(Wa, d.E & 0xFFFFu, hist_addr, trash_addr, gh);
        emit16<K, MODE>(Wb, d.E >> 16, hist_addr, trash_addr, gh);
        cur = nxt;
    }

    if (MODE == kSmem32) {
        __syncthreads();
        uint32_t* slab = slabs + (size_t)logical_cta() * NK;
        for (uint32_t i = tid; i < NK; i += blockDim.x) slab[i] = s_hist[i];
    }
}

// ------------------------------------------------------------------------------ chunk-table kernels (k <= 7)
// The chunk stream above works out, in every iteration, which read and which of its chunks a lane is looking at (~80
// of the ~250 ALU-pipe instructions of an iteration).  With the chunk TABLE (vk_bucket.cuh, chunk mode) that work is done
// once, by the scatter kernel, which knows it anyway: a segment is a flat array of 8-byte descriptors, a warp claims
// ranges of it from the segment's counter and lane l simply takes descriptor pos + l (one coalesced 8-byte load,
// requested an iteration ahead of the text it points at).  The chunks of a read are consecutive in the table, so the K-1
// bases to the left still come from the lane to the left; lane 0 of the first window of a claimed range reads the 8 text
// bytes in front of its chunk itself.
struct DescStream {
    const uint4* text16;
    const uint64_t* seg_chunks;
    unsigned long long* seg_counter;
    uint32_t n_chunks, lane, n_warps;
    uint32_t pos, end;          // where the descriptor loads are (two windows ahead of the window handed out) / end of that range
    uint32_t nbase, nsize;      // the range after that one, claimed a range ahead
    uint32_t claimed_upto;      // counter value after our last claim (guides the claim size)
    uint64_t dA, dB;            // descriptors of the next window and of the one after it
    uint64_t text_words;        // VK_ASSERT only
    bool fA, fB, fpos;          // "opens a claimed range" of dA, dB and of the window at pos
    bool is_last;               // the chunk handed out by the last fetch() is its read's last

    // The kernel's speed is set by how many text bytes it keeps in flight: a lane's 32 bytes sit somewhere in a 400 MB
    // text, and with ONE window requested ahead (32 KB per SM) the loads' DRAM latency capped the kernel at 2.7 TB/s
    // whatever the arithmetic did.  The descriptors therefore run two windows ahead, and the text of the window after
    // next is pulled into L2 (prefetch.global.L2: no register waits for it) while the next window's demand loads fly.
    __device__ __forceinline__ void claim()
    {
        const uint32_t left = claimed_upto < n_chunks ? n_chunks - claimed_upto : 0u;
        const uint32_t per_warp = left / (32u * n_warps);                 // windows still unclaimed per warp, roughly
        const uint32_t its = per_warp >= 64u ? 16u : per_warp >= 8u ? 4u : 1u;
        unsigned long long r0 = 0;
        if (lane == 0) r0 = atomicAdd(seg_counter, (unsigned long long)(its * 32u));
        r0 = __shfl_sync(0xffffffffu, r0, 0);
        nbase = r0 < 0xffffffe0ull ? (uint32_t)r0 : 0xffffffe0u;          // a range at or beyond n_chunks is empty
        nsize = its * 32u;
        claimed_upto = nbase + nsize;
    }
    __device__ __forceinline__ uint64_t load_desc(uint32_t idx) const { return idx < n_chunks ? __ldg(seg_chunks + idx) : 0ull; }
    __device__ __forceinline__ void advance()
    {
        pos += 32u;
        fpos = false;
        if (pos >= end) {                    // the range is used up: go on in the one claimed a range ago, claim another
            pos = nbase;
            end = nbase + nsize;
            fpos = true;
            claim();
        }
    }
    __device__ __forceinline__ void prefetch_text(uint64_t d) const
    {
        if (d & kChunkValid) {
            const uint4* const ptr = text16 + chunk_word16(d);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
            if (chunk_hi(d) > 16u) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + 1));
        }
    }
    __device__ __forceinline__ void init(const uint4* t, const uint64_t* sc, uint32_t n, unsigned long long* ctr, uint32_t ln,
                                         uint64_t n_bytes, uint32_t n_warps_seg)
    {
        text16 = t; seg_chunks = sc; n_chunks = n; seg_counter = ctr; lane = ln;
        text_words = (n_bytes + 15) >> 4;
        n_warps = n_warps_seg ? n_warps_seg : 1u;
        claimed_upto = 0;
        claim();
        pos = nbase; end = nbase + nsize;
        fpos = true;
        claim();
        dA = load_desc(pos + lane); fA = fpos; advance();
        dB = load_desc(pos + lane); fB = fpos; advance();
        prefetch_text(dA);
        prefetch_text(dB);
    }
    // this lane's chunk of the coming window (c.range == 0: none).  first: the window opens a claimed range; left8: for
    // lane 0 of such a window, the 8 text bytes in front of its chunk (when the chunk is not its read's first)
    __device__ __forceinline__ Chunk fetch(bool& first, uint2& left8)
    {
        Chunk c;
        const uint64_t d = dA;
        first = fA;
        c.wa = make_uint4(0, 0, 0, 0);
        c.wb = c.wa;
        left8 = make_uint2(0, 0);
        const bool act = (d & kChunkValid) != 0;
        const uint64_t word16 = chunk_word16(d);
        const uint32_t rlo = chunk_rlo(d), hi = chunk_hi(d), j = chunk_j(d);
        const uint32_t lo = j == 0 ? rlo : 0u;
        if (act) {
            const uint4* const ptr = text16 + word16;
            VK_ASSERT(word16 + (hi > 16u ? 1 : 0) < text_words && hi > lo);
            c.wa = ldg_text(ptr);
            if (hi > 16u) c.wb = ldg_text(ptr + 1);
            if (first && lane == 0 && j != 0) left8 = __ldg(reinterpret_cast<const uint2*>(ptr) - 1);
        }
        c.range = act ? (0xffffffffu >> (32u - hi)) & (0xffffffffu << lo) : 0u;
        c.j = j;
        c.rlen = chunk_long(d) ? 0xFFFFFFu : 0u;          // only "longer than the break length" is known here
        c.q0 = (int32_t)(32u * j) - (int32_t)rlo;
        c.last = 31u;
        is_last = chunk_last(d);
        // shift the pipeline: the window after next gets its text pulled into L2, a new descriptor is requested
        dA = dB; fA = fB;
        dB = load_desc(pos + lane); fB = fpos;
        advance();
        prefetch_text(dA);
        return c;
    }
};

// the K-1 bases (and their validity) in front of a chunk, from the 8 text bytes in front of it: the `tail` format of
// decode_chunk (codes of the last K-1 bases from bit 0, their validity bits from bit 16)
template <int K>
__device__ __forceinline__ uint32_t tail_from_left8(uint2 l8, uint32_t one)
{
    constexpr int KM1 = K - 1;
    const Cls4z a = classify4z(l8.x, one), b = classify4z(l8.y, one);
    const uint32_t codes16 = (a.packed_hi >> 24) | ((b.packed_hi >> 24) << 8);      // base -8 at bits 0-1 ... base -1 at bits 14-15
    const uint32_t v8 = gather8(a.z, b.z) >> 24;                                    // bit 0 = byte -8 ... bit 7 = byte -1
    return (codes16 >> (16 - 2 * KM1)) | ((v8 >> (8 - KM1)) << 16);
}

template <int K>
__global__ void __launch_bounds__(kCountThreads)
countd_kernel(const StepArgs* __restrict__ sa, const uint64_t* __restrict__ chunks, Plan* __restrict__ plan,
              uint32_t* __restrict__ slabs)
{
    pdl_wait();
    static_assert(K <= 7, "u32 bins of 4^K fit shared memory up to k = 7");
    const uint4* __restrict__ text16 = reinterpret_cast<const uint4*>(sa->text);
    const int breaklen = sa->pa.p.breaklength;
    const uint32_t one = (uint32_t)(sa->n_bytes >> 62) + 1u;          // 1 (texts are shorter than 2^40), but not to ptxas
    constexpr uint32_t NK = 1u << (2 * K);
    constexpr int KM1 = K - 1;
    constexpr uint32_t FULL = 0xffffffffu;
    extern __shared__ uint32_t s_raw[];           // [pad to a 64 KiB shared address][NK bins][trash word]
    const uint32_t tid = threadIdx.x, lane = tid & 31;

    const int seg = cta_segment(plan, lane);
    if (seg < 0) return;
    const unsigned long long nc64 = plan->seg_chunks[seg] < plan->seg_ccap[seg] ? plan->seg_chunks[seg] : plan->seg_ccap[seg];
    const uint32_t n_chunks = nc64 < 0xffffff00ull ? (uint32_t)nc64 : 0xffffff00u;

    const uint32_t raw_addr = (uint32_t)__cvta_generic_to_shared(s_raw);
    const uint32_t hist_addr = (raw_addr + 0xFFFFu) & ~0xFFFFu;
    uint32_t* const s_hist = s_raw + ((hist_addr - raw_addr) >> 2);
    const uint32_t trash_addr = hist_addr + NK * 4u;
    for (uint32_t i = tid; i < NK + 32; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();

    DescStream ds;
    ds.init(text16, chunks + plan->seg_cbegin[seg], n_chunks, &plan->seg_next[seg], lane, plan->n_bytes,
            (plan->seg_cta_begin[seg + 1] - plan->seg_cta_begin[seg]) * (blockDim.x >> 5));
    uint32_t carry = 0;
    bool cur_first, nxt_first;
    uint2 cur_left, nxt_left;
    Chunk cur = ds.fetch(cur_first, cur_left);
    while (__ballot_sync(FULL, cur.range != 0) != 0) {
        const Chunk nxt = ds.fetch(nxt_first, nxt_left);
        if (cur_first) carry = tail_from_left8<K>(cur_left, one);      // warp-uniform; only lane 0's value is used
        const Decoded d = decode_chunk<K, false>(cur, carry, lane, breaklen, one);
        const uint64_t Wa = ((uint64_t)d.Cc | ((uint64_t)d.Plo << (2 * KM1))) << 2;
        const uint64_t Wb = ((uint64_t)(d.Plo >> (32 - 2 * KM1)) | ((uint64_t)d.Phi << (2 * KM1))) << 2;
        emit16<K, kSmem32>(Wa, d.E & 0xFFFFu, hist_addr, trash_addr, nullptr);
        emit16<K, kSmem32>(Wb, d.E >> 16, hist_addr, trash_addr, nullptr);
        cur = nxt;
        cur_first = nxt_first;
        cur_left = nxt_left;
    }
    __syncthreads();
    uint32_t* slab = slabs + (size_t)logical_cta() * NK;
    for (uint32_t i = tid; i < NK; i += blockDim.x) slab[i] = s_hist[i];
}

// ------------------------------------------------------------------------------------------ kSmem16
// 4^8 bins of 16 bits, two per 32-bit word: bin b lives in word (b & 0x7FFF); bins with bit 15 clear are counted
// by adding 1, bins with bit 15 set by adding 0x10001.  So the LOW half of a word is the total of its two bins and
// the HIGH half the count of the upper bin: the increment is (sh & 0x20000) / 2 + 1 (one LOP3 + one IMAD.HI), and
// "the low half stays below 2^16" is the only overflow condition.
//
// Exactness.  Every increment is an ATOMS.ADD that returns the old word; a lane ORs what it gets back and looks at
// the OR after every 16 increments.  A returned low half of 0x4000 or more means that word is running hot: the lane
// then re-reads the (at most 16) words it has just touched and DRAINS the hot ones -- compare-and-swap the word to 0
// and credit what it held to the CTA's u32 slab in global memory (rare global atomics).  The swap makes a drain happen
// once however many lanes notice the same word.  Bound: after a word reaches 0x4000, every warp can add at most the
// 16 increments per lane it may be in the middle of plus 16 more before its own next look, 32 warps x 32 lanes x 32 =
// 32 768, so a low half never passes 0x4000 + 0x8000 < 2^16.  No barrier is involved; on ordinary reads nothing is
// ever drained (a CTA sees ~10^6 bases; one 8-mer would have to make up more than 1 % of them).
template <int K>
__device__ __forceinline__ void credit16(uint32_t* __restrict__ slab, uint32_t word_idx, uint32_t w)
{
    const uint32_t hi = w >> 16, lo = (w & 0xFFFFu) - hi;           // counts of the upper / lower bin of the word
    const uint32_t bl = word_idx, bu = word_idx | 0x8000u;
    if (K == 8) {
        if (lo) atomicAdd(slab + bl, lo);
        if (hi) atomicAdd(slab + bu, hi);
    } else {                                                        // an 8-mer stands for its first and its last 7-mer
        if (lo) { atomicAdd(slab + (bl & 0x3FFFu), lo); atomicAdd(slab + (bl >> 2), lo); }
        if (hi) { atomicAdd(slab + (bu & 0x3FFFu), hi); atomicAdd(slab + (bu >> 2), hi); }
    }
}
template <int K>
__device__ __forceinline__ void drain16(uint32_t* __restrict__ h8, uint32_t* __restrict__ slab, uint32_t word_idx)
{
    VK_ASSERT(word_idx < 32768u);
    uint32_t w = *reinterpret_cast<volatile uint32_t*>(h8 + word_idx);
    while ((w & 0xFFFFu) >= 0x4000u) {
        const uint32_t seen = atomicCAS(h8 + word_idx, w, 0u);
        if (seen == w) { credit16<K>(slab, word_idx, w); break; }
        w = seen;
    }
}

template <int K>
__device__ __forceinline__ void count16_flush(uint32_t* __restrict__ h8, uint32_t* __restrict__ h7, uint32_t* __restrict__ slab,
                                              uint32_t tid, uint32_t nthr)
{
    if (K == 8) {
        for (uint32_t w = tid; w < 32768u; w += nthr) {
            const uint32_t v = h8[w];
            const uint32_t hi = v >> 16, lo = (v & 0xFFFFu) - hi;
            slab[w] += lo;
            slab[w | 0x8000u] += hi;
        }
    } else {
        // 7-mer x: singles + 8-mers that start with it (x | c << 14) + 8-mers that end with it ((x << 2 | c) & 0xFFFF)
        for (uint32_t x = tid; x < 16384u; x += nthr) {
            uint32_t v = h7[x];
            v += (h8[x] & 0xFFFFu) + (h8[x | 0x4000u] & 0xFFFFu);         // low halves = totals of both bins of a word
            const uint32_t w0 = (x & 0x1FFFu) << 2;
            const bool upper = (x & 0x2000u) != 0;                          // bit 15 of (x << 2 | c)
#pragma unroll
            for (uint32_t c = 0; c < 4; ++c) {
                const uint32_t w = h8[w0 | c];
                v += upper ? (w >> 16) : (w & 0xFFFFu) - (w >> 16);
            }
            slab[x] += v;
        }
    }
}

// FAST: the increments are fire-and-forget (red.shared.add, no value comes back, nothing is watched in the loop) and the
// bins are checked ONCE, when the CTA is done: every increment adds 1 to the low half of its word, so the low halves must
// sum to the number of increments the CTA made; a low half that passed 2^16 carried into its high half and the sum
// comes out 2^16 short.  A mismatch raises plan->count_overflow and the host repeats the count with the exact kernel
// (returning adds + drains, below; for k = 7 the u32 kernel).  It takes a word that receives more than 65 535 of one
// CTA's increments -- a flood of one k-mer (poly-A libraries); ordinary reads never get near it.
template <int K, bool PACKED, bool FAST>
__global__ void __launch_bounds__(kCountThreads)
count16_kernel(const StepArgs* __restrict__ sa, PackedSrc pk, const uint64_t* __restrict__ sorted, Plan* __restrict__ plan,
               uint32_t* __restrict__ slabs)
{
    pdl_wait();
    const uint4* __restrict__ text16 = reinterpret_cast<const uint4*>(sa->text);
    const int breaklen = sa->pa.p.breaklength;
    const uint32_t one = (uint32_t)(sa->n_bytes >> 62) + 1u;          // 1 (texts are shorter than 2^40), but not to ptxas
    static_assert(K == 7 || K == 8, "16-bit bins: k = 8 directly, k = 7 through pairs");
    constexpr uint32_t NK = 1u << (2 * K);
    constexpr int KM1 = K - 1;
    constexpr uint32_t FULL = 0xffffffffu;
    extern __shared__ uint32_t s_raw[];           // [h8: 32768 words][h7: 16384 words (k = 7)]
    const uint32_t tid = threadIdx.x, lane = tid & 31, nthr = blockDim.x;

    const int seg = cta_segment(plan, lane);
    if (seg < 0) return;
    const uint32_t seg_len = (uint32_t)(plan->seg_reads[seg] < plan->seg_cap[seg] ? plan->seg_reads[seg] : plan->seg_cap[seg]);

    uint32_t* const h8 = s_raw;
    uint32_t* const h7 = s_raw + 32768;
    constexpr uint32_t kWords = 32768u + (K == 7 ? 16384u : 0u);
    const uint32_t h8_addr = (uint32_t)__cvta_generic_to_shared(s_raw);
    const uint32_t h7_addr = h8_addr + 32768u * 4u;
    uint32_t* const slab = slabs + (size_t)logical_cta() * NK;
    __shared__ unsigned long long s_chk[2];                            // FAST: increments made / sum of the low halves
    for (uint32_t i = tid; i < kWords; i += nthr) s_raw[i] = 0;
    for (uint32_t i = tid; i < NK; i += nthr) slab[i] = 0;             // drains and the final fold ADD to the slab
    if (tid < 2) s_chk[tid] = 0;
    __syncthreads();

    ChunkStream<PACKED> cs;
    cs.init(text16, pk, sorted + plan->seg_begin[seg], seg_len, &plan->seg_next[seg], lane, plan->n_bytes,
            (plan->seg_cta_begin[seg + 1] - plan->seg_cta_begin[seg]) * (blockDim.x >> 5));
    uint32_t carry = 0;
    uint32_t made = 0;                                                 // FAST: increments of this lane (< 2^32: < 2^37 bases per lane)
    Chunk cur = cs.fetch();
    while (__ballot_sync(FULL, cur.range != 0) != 0) {
        const Chunk nxt = cs.fetch();
        const Decoded d = decode_chunk<K, PACKED>(cur, carry, lane, breaklen, one);
        const uint64_t Wa = ((uint64_t)d.Cc | ((uint64_t)d.Plo << (2 * KM1))) << 2;
        const uint64_t Wb = ((uint64_t)(d.Plo >> (32 - 2 * KM1)) | ((uint64_t)d.Phi << (2 * KM1))) << 2;
        if (K == 8 && FAST) {
            made += __popc(d.E);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint64_t W4 = h ? Wb : Wa;
                const uint32_t Wl = (uint32_t)W4, Wh = (uint32_t)(W4 >> 32), E = h ? d.E >> 16 : d.E & 0xFFFFu;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t sh = __funnelshift_r(Wl, Wh, 2 * j);
                    const uint32_t inc = ((sh >> 1) & 0x10000u) | 1u;           // 1, or 0x10001 for the upper bin of the word
                    smem_add(h8_addr + (sh & 0x1FFFCu), (E >> j) & 1u ? inc : 0u);
                }
            }
        } else if (K == 7 && FAST) {
            // pairs: positions (2m, 2m+1); Eb: both 7-mers countable -> one 8-mer; Es: exactly one -> a single 7-mer
            const uint32_t Ee = d.E & 0x55555555u, Eo = (d.E >> 1) & 0x55555555u;
            const uint32_t Eb = Ee & Eo;
            uint32_t Es = Ee ^ Eo;                                      // bit 2m: pair m holds exactly one 7-mer
            made += __popc(Eb);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint64_t W4 = h ? Wb : Wa;
                const uint32_t Wl = (uint32_t)W4, Wh = (uint32_t)(W4 >> 32), E = h ? Eb >> 16 : Eb & 0xFFFFu;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const uint32_t sh = __funnelshift_r(Wl, Wh, 4 * m);
                    const uint32_t inc = ((sh >> 1) & 0x10000u) | 1u;
                    smem_add(h8_addr + (sh & 0x1FFFCu), (E >> (2 * m)) & 1u ? inc : 0u);
                }
            }
            while (__ballot_sync(FULL, Es != 0) != 0) {                 // single 7-mers (read ends, N, break points)
                if (Es != 0) {
                    const uint32_t b2 = __ffs(Es) - 1;                  // = 2m
                    Es &= Es - 1;
                    const uint64_t W4 = b2 >= 16 ? Wb : Wa;
                    const uint32_t sh = (uint32_t)(W4 >> (2 * (b2 & 15u)));
                    const bool second = (Eo >> b2) & 1u;
                    const uint32_t off7 = (second ? sh >> 2 : sh) & 0xFFFCu;
                    smem_inc(h7_addr + off7);
                }
            }
        } else if (K == 8) {
            // every 8-mer: window position j holds the 8-mer that ends at base j (7 carried codes in front).  A lane
            // without an 8-mer at j adds 0 to whatever word the window names (always inside the table).
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint64_t W4 = h ? Wb : Wa;
                const uint32_t Wl = (uint32_t)W4, Wh = (uint32_t)(W4 >> 32), E = h ? d.E >> 16 : d.E & 0xFFFFu;
                uint32_t acc = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t sh = __funnelshift_r(Wl, Wh, 2 * j);
                    const uint32_t inc = __umulhi(sh & 0x20000u, 0x80000000u) + 1u;
                    acc |= smem_add_ret(h8_addr + (sh & 0x1FFFCu), (E >> j) & 1u ? inc : 0u);
                }
                if (__ballot_sync(FULL, (acc & 0xC000u) != 0) != 0) {          // rare: some word is running hot
                    if (acc & 0xC000u)
                        for (int j = 0; j < 16; ++j)
                            drain16<K>(h8, slab, ((uint32_t)(W4 >> (2 * j)) & 0x1FFFCu) >> 2);
                }
            }
        } else {
            // pairs: positions (2m, 2m+1); Eb: both 7-mers countable -> one 8-mer; Es: exactly one -> a single 7-mer
            const uint32_t Ee = d.E & 0x55555555u, Eo = (d.E >> 1) & 0x55555555u;
            const uint32_t Eb = Ee & Eo;
            uint32_t Es = Ee ^ Eo;                                      // bit 2m: pair m holds exactly one 7-mer
            uint32_t acc = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint64_t W4 = h ? Wb : Wa;
                const uint32_t Wl = (uint32_t)W4, Wh = (uint32_t)(W4 >> 32), E = h ? Eb >> 16 : Eb & 0xFFFFu;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    // the 8-mer that ends at base 2m+1 starts at window position 2m (6 carried codes in front)
                    const uint32_t sh = __funnelshift_r(Wl, Wh, 4 * m);
                    const uint32_t inc = __umulhi(sh & 0x20000u, 0x80000000u) + 1u;
                    acc |= smem_add_ret(h8_addr + (sh & 0x1FFFCu), (E >> (2 * m)) & 1u ? inc : 0u);
                }
            }
            if (__ballot_sync(FULL, (acc & 0xC000u) != 0) != 0) {              // rare: some word is running hot
                if (acc & 0xC000u)
                    for (int m = 0; m < 16; ++m)
                        drain16<K>(h8, slab, ((uint32_t)((m < 8 ? Wa : Wb) >> (4 * (m & 7))) & 0x1FFFCu) >> 2);
            }
            // single 7-mers (read ends, N, break points): a few per warp and iteration
            while (__ballot_sync(FULL, Es != 0) != 0) {
                if (Es != 0) {
                    const uint32_t b2 = __ffs(Es) - 1;                  // = 2m
                    Es &= Es - 1;
                    const uint64_t W4 = b2 >= 16 ? Wb : Wa;
                    const uint32_t sh = (uint32_t)(W4 >> (2 * (b2 & 15u)));     // window position 2m, as above
                    const bool second = (Eo >> b2) & 1u;                // the 7-mer that ends at 2m+1: last 7 bases
                    const uint32_t off7 = (second ? sh >> 2 : sh) & 0xFFFCu;
                    smem_inc(h7_addr + off7);
                }
            }
        }
        cur = nxt;
    }
    __syncthreads();
    if (FAST) {
        unsigned long long low = 0;
        for (uint32_t w = tid; w < 32768u; w += nthr) low += h8[w] & 0xFFFFu;
        unsigned long long mine = made;
#pragma unroll
        for (int dlt = 16; dlt > 0; dlt >>= 1) {
            low += __shfl_xor_sync(FULL, low, dlt);
            mine += __shfl_xor_sync(FULL, mine, dlt);
        }
        if (lane == 0) { atomicAdd(&s_chk[0], mine); atomicAdd(&s_chk[1], low); }
        __syncthreads();
        if (tid == 0 && s_chk[0] != s_chk[1]) atomicOr(&plan->count_overflow, 1u);
    }
    count16_flush<K>(h8, h7, slab, tid, nthr);
}

// ---- k = 7 in PAIRS from the chunk table.  Two consecutive 7-mers are one 8-mer: one shared-memory increment per base
// PAIR, into 4^8 16-bit bins (two per 32-bit word; count16_kernel below describes the bin format, its FAST form the
// fire-and-forget increments and the end-of-kernel checksum that this kernel shares).  What makes pairs pay here:
//   * pairs are aligned to the READ, not to the text: the first 7-mer of a read (or of a 500-base piece: the break length
//     is even) ends at chunk position rlo + 6, so with parity = rlo & 1 every 7-mer of an N-free read has its partner and
//     only an odd count of 7-mers leaves ONE single at the read's end -- text-aligned pairs left a single at one end of
//     every other read and a loop over singles in every iteration;
//   * a pair is owned by the chunk that holds its SECOND 7-mer: the chunk carries 7 bases (not 6) from its left neighbour
//     and counts the pair (-1, 0) itself; a first 7-mer at position 31 is left to the next chunk, unless the chunk is its
//     read's last (descriptor flag), then it is a single;
//   * singles (read ends, around N) go to a 4^7 x u32 table in a loop that most iterations skip.
// Per pair: funnel shift, address mask, half-select bit -> increment (multiply-add on the FMA pipe), one ATOMS.ADD.
__global__ void __launch_bounds__(kCountThreads)
countp_kernel(const StepArgs* __restrict__ sa, const uint64_t* __restrict__ chunks, Plan* __restrict__ plan,
              uint32_t* __restrict__ slabs)
{
    pdl_wait();
    constexpr int K = 7;
    constexpr uint32_t NK = 1u << (2 * K);
    constexpr uint32_t FULL = 0xffffffffu;
    const uint4* __restrict__ text16 = reinterpret_cast<const uint4*>(sa->text);
    const int breaklen = sa->pa.p.breaklength;
    const uint32_t zero = (uint32_t)(sa->n_bytes >> 62);              // 0 (texts are shorter than 2^40), but not to ptxas
    const uint32_t one = zero + 1u, two31 = 0x80000000u >> zero;
    extern __shared__ uint32_t s_raw[];           // [h8: 32768 words][h7: 16384 words]
    __shared__ unsigned long long s_chk[2];
    const uint32_t tid = threadIdx.x, lane = tid & 31, nthr = blockDim.x;

    const int seg = cta_segment(plan, lane);
    if (seg < 0) return;
    const unsigned long long nc64 = plan->seg_chunks[seg] < plan->seg_ccap[seg] ? plan->seg_chunks[seg] : plan->seg_ccap[seg];
    const uint32_t n_chunks = nc64 < 0xffffff00ull ? (uint32_t)nc64 : 0xffffff00u;

    uint32_t* const h8 = s_raw;
    uint32_t* const h7 = s_raw + 32768;
    const uint32_t h8_addr = (uint32_t)__cvta_generic_to_shared(s_raw);
    const uint32_t h7_addr = h8_addr + 32768u * 4u;
    uint32_t* const slab = slabs + (size_t)logical_cta() * NK;
    for (uint32_t i = tid; i < 32768u + 16384u; i += nthr) s_raw[i] = 0;
    for (uint32_t i = tid; i < NK; i += nthr) slab[i] = 0;
    if (tid < 2) s_chk[tid] = 0;
    __syncthreads();

    DescStream ds;
    ds.init(text16, chunks + plan->seg_cbegin[seg], n_chunks, &plan->seg_next[seg], lane, plan->n_bytes,
            (plan->seg_cta_begin[seg + 1] - plan->seg_cta_begin[seg]) * (blockDim.x >> 5));
    uint32_t carry = 0;             // tail of lane 31 of the previous window: 7 codes (bits 0..13) + 7 validity bits (16..22)
    uint32_t made = 0;              // pair increments of this lane
    bool cur_first, nxt_first;
    uint2 cur_left, nxt_left;
    Chunk cur = ds.fetch(cur_first, cur_left);
    bool cur_last = ds.is_last;
    while (__ballot_sync(FULL, cur.range != 0) != 0) {
        const Chunk nxt = ds.fetch(nxt_first, nxt_left);
        const bool nxt_last = ds.is_last;
        // ---- codes and validity of the 32 bytes (as decode_chunk), 7 bases of left context
        const Cls4z c0 = classify4z(cur.wa.x, one), c1 = classify4z(cur.wa.y, one), c2 = classify4z(cur.wa.z, one), c3 = classify4z(cur.wa.w, one);
        const Cls4z c4 = classify4z(cur.wb.x, one), c5 = classify4z(cur.wb.y, one), c6 = classify4z(cur.wb.z, one), c7 = classify4z(cur.wb.w, one);
        const uint32_t v01 = gather8(c0.z, c1.z), v23 = gather8(c2.z, c3.z), v45 = gather8(c4.z, c5.z), v67 = gather8(c6.z, c7.z);
        const uint32_t V = __byte_perm(__byte_perm(v01, v23, 0x0073), __byte_perm(v45, v67, 0x0073), 0x5410) & cur.range;
        const uint32_t Plo = __byte_perm(__byte_perm(c0.packed_hi, c1.packed_hi, 0x0073), __byte_perm(c2.packed_hi, c3.packed_hi, 0x0073), 0x5410);
        const uint32_t Phi = __byte_perm(__byte_perm(c4.packed_hi, c5.packed_hi, 0x0073), __byte_perm(c6.packed_hi, c7.packed_hi, 0x0073), 0x5410);
        if (cur_first) {            // warp-uniform; only lane 0's value is used: the 8 text bytes in front of its chunk
            const Cls4z a = classify4z(cur_left.x, one), b = classify4z(cur_left.y, one);
            const uint32_t codes16 = (a.packed_hi >> 24) | ((b.packed_hi >> 24) << 8);
            carry = (codes16 >> 2) | (((gather8(a.z, b.z) >> 24) >> 1) << 16);
        }
        const uint32_t tail = (Phi >> 18) | ((V >> 25) << 16);
        uint32_t hist = __shfl_up_sync(FULL, tail, 1);
        if (lane == 0) hist = carry;
        if (cur.j == 0) hist = 0;
        carry = __shfl_sync(FULL, tail, 31);
        const uint32_t C7 = hist & 0x3FFFu;
        const uint64_t VW = (uint64_t)(hist >> 16) | ((uint64_t)V << 7);        // bit i <-> base i - 7 of the chunk
        uint64_t EE = runs_of_k64<K>(VW) & 0x1FFFFFFFFull;                      // bit p + 1: the 7-mer that ends at base p, p = -1..31
        if (breaklen > 0 && __ballot_sync(FULL, cur.rlen != 0) != 0) {
            // reformat.sh breaklength: no 7-mer may span a multiple of breaklen counted from the read's first base
            uint64_t dead = 0;
            if (cur.rlen != 0 && cur.range != 0) {
                const int32_t q0 = cur.q0;
                int32_t cpos = (q0 > 0 ? q0 / breaklen : 0) * breaklen;
                if (cpos < breaklen) cpos = breaklen;
                for (; cpos - q0 < 32; cpos += breaklen) {
                    const int32_t b = cpos - q0;                        // chunk byte that starts the new piece: ends b .. b+5 are dead
                    if (b + 1 > -(K - 1)) dead |= b + 1 >= 0 ? (uint64_t)((1u << (K - 1)) - 1u) << (b + 1) : (uint64_t)((1u << (K - 1)) - 1u) >> (-(b + 1));
                }
            }
            EE &= ~dead;
        }
        // ---- pairs in read parity
        const uint32_t par = (uint32_t)(-cur.q0) & 1u;                 // rlo & 1: first elements sit at chunk positions of this parity
        const uint32_t Q = par ? (uint32_t)EE : (uint32_t)(EE >> 1);   // bit 2m: first, bit 2m + 1: second 7-mer of pair slot m
        const uint32_t F = Q & 0x55555555u, S = (Q >> 1) & 0x55555555u;
        const uint32_t both = F & S;
        uint32_t s1 = F & ~S, s2 = S & ~F;                             // singles: first without second / second without first
        const bool dangling = par && ((EE >> 32) & 1u) && cur_last;    // a first 7-mer at position 31 of a read's last chunk
        made += __popc(both);
        // window with 7 bases of context, shifted so that pair slot m is the 16-bit field at bit 4m; times 4 (byte offsets)
        const uint32_t sft = par ? 0u : 2u;
        const uint64_t WA = (((uint64_t)C7 | ((uint64_t)Plo << 14)) >> sft) << 2;
        const uint64_t WB = ((((uint64_t)(Plo >> 18)) | ((uint64_t)Phi << 14)) >> sft) << 2;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint64_t W4 = h ? WB : WA;
            const uint32_t Wl = (uint32_t)W4, Wh = (uint32_t)(W4 >> 32), B = h ? both >> 16 : both & 0xFFFFu;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const uint32_t sh = __funnelshift_r(Wl, Wh, 4 * m);
                uint32_t inc = 0;
                if ((B >> (2 * m)) & 1u) inc = mad_hi_u32(sh & 0x20000u, two31, 1u);      // 1, or 0x10001 for the upper bin of the word
                smem_add(h8_addr + (sh & 0x1FFFCu), inc);
            }
        }
        // ---- singles: rare (read ends with an odd number of 7-mers, neighbours of N)
        uint32_t sing = s1 | (s2 << 1);                                // bit 2m: first of slot m, bit 2m + 1: second
        bool dang = dangling;
        while (__ballot_sync(FULL, sing != 0 || dang) != 0) {
            if (sing != 0) {
                const uint32_t bit = __ffs(sing) - 1;
                sing &= sing - 1;
                const uint32_t m = bit >> 1;
                const uint64_t W4 = m >= 8 ? WB : WA;
                const uint32_t f16 = (uint32_t)(W4 >> (4 * (m & 7u)));          // (8-mer of slot m) << 2
                const uint32_t off7 = ((bit & 1u) ? f16 >> 2 : f16) & 0xFFFCu;   // its last / its first 7 bases
                smem_inc(h7_addr + off7);
            } else if (dang) {
                smem_inc(h7_addr + ((Phi >> 18) << 2));                          // the 7-mer that ends at position 31
                dang = false;
            }
        }
        cur = nxt;
        cur_first = nxt_first;
        cur_left = nxt_left;
        cur_last = nxt_last;
    }
    __syncthreads();
    {
        unsigned long long low = 0;
        for (uint32_t w = tid; w < 32768u; w += nthr) low += h8[w] & 0xFFFFu;
        unsigned long long mine = made;
#pragma unroll
        for (int dlt = 16; dlt > 0; dlt >>= 1) {
            low += __shfl_xor_sync(FULL, low, dlt);
            mine += __shfl_xor_sync(FULL, mine, dlt);
        }
        if (lane == 0) { atomicAdd(&s_chk[0], mine); atomicAdd(&s_chk[1], low); }
        __syncthreads();
        if (tid == 0 && s_chk[0] != s_chk[1]) atomicOr(&plan->count_overflow, 1u);
    }
    count16_flush<7>(h8, h7, slab, tid, nthr);
}

// ------------------------------------------------------------------------------------------ k = 9: canonical halves
// 4^9 forward bins do not fit shared memory even at 16 bits (512 KiB).  The 2 x 4^8 CANONICAL classes of an odd k do,
// as two tables of 4^8 16-bit bins: CTAs work in PAIRS, both CTAs of a pair stream the same reads (the second read of
// a text line is an L2 hit while the two stay close), and each counts the classes of its half with the exact
// returning-add / drain scheme of count16_kernel<8>.
//
// The class of a 9-mer needs no comparison: a k-mer and its reverse complement have complementary MIDDLE bases, and
// under the code A0 C1 T2 G3 the complement is XOR 2, so exactly one of the two has bit 9 (the high bit of base 4)
// clear -- that one is the representative x, and dropping the zero bit gives a dense 17-bit class id:
//     id = x[0..8] | x[10..17] << 9,   bin = id & 0xFFFF (word id & 0x7FFF, upper bin iff id bit 15),   half = id bit 16.
// The reverse-complement index comes from a second window: the stream with its 2-bit groups reversed and complemented
// (two BREV per 16 positions), in which the reverse complement of the 9-mer ending at position j is again a contiguous
// field.  reduce_slabs9h_kernel writes a class total to seg_hist[x] and 0 to seg_hist[rc x]; fold_kernel's
// fwd[i] + fwd[rc i] then yields the canonical abundance for both, unchanged.
__device__ __forceinline__ uint64_t revcomp_groups64(uint64_t s)
{
    uint32_t a = __brev((uint32_t)(s >> 32)), b = __brev((uint32_t)s);          // a: new low word, b: new high word
    a = ((a & 0x55555555u) << 1) | ((a >> 1) & 0x55555555u);                   // bit order inside each group back
    b = ((b & 0x55555555u) << 1) | ((b >> 1) & 0x55555555u);
    return (uint64_t)(a ^ 0xAAAAAAAAu) | ((uint64_t)(b ^ 0xAAAAAAAAu) << 32);   // complement = XOR 2 per group
}
// x << 2 of the representative of the 9-mer that ends at window position j (garbage above bit 19)
template <int J>
__device__ __forceinline__ uint32_t rep9_x4(uint32_t Wl, uint32_t Wh, uint32_t Rl, uint32_t Rh)
{
    const uint32_t f4 = __funnelshift_r(Wl, Wh, 2 * J);
    const uint32_t r4 = J == 0 ? Rh : __funnelshift_r(Rl, Rh, 32 - 2 * J);
    return (f4 & 0x800u) ? r4 : f4;
}
__device__ __forceinline__ uint32_t rep9_offset(uint32_t x4) { return (x4 & 0x7FCu) + ((x4 >> 1) & 0x1F800u); }
__device__ __forceinline__ uint32_t mad_hi(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// The kernel is bound by the ALU pipe (97 % busy in the first version, FMA pipe 5 %), so everything that can be a
// multiply-add is one: the multipliers are powers of two held in registers the compiler cannot see through
// (Opaque), which keeps ptxas from turning the IMADs back into shifts.
struct Opaque9 {
    uint32_t two31;      // x * 2^31 >> 32 = x >> 1
    uint32_t two11;      // x * 2^11
    uint32_t h_addr;     // table base, 2 KiB aligned: (x4 & 0x7FC) | h_addr is one LOP3
    uint32_t other19;    // (1 - half) << 19: bit 19 of x4 ^ other19 is set for a k-mer of this CTA's half
};

template <int J>
__device__ __forceinline__ void count9_steps(uint32_t Wl, uint32_t Wh, uint32_t Rl, uint32_t Rh, uint32_t E, const Opaque9& q,
                                             uint32_t& acc)
{
    if constexpr (J < 16) {
        const uint32_t x4 = rep9_x4<J>(Wl, Wh, Rl, Rh);
        // address: bits 2..10 stay, bits 12..17 move down by one (the hole is the representative's zero bit)
        const uint32_t addr = mad_hi(x4 & 0x3F000u, q.two31, and_or(x4, 0x7FCu, q.h_addr));
        // increment << 13: 1 (lower bin of the word) or 0x10001 (upper bin, bit 18), then x 2^-13 if the half is ours
        const uint32_t inc13 = mad_lo(x4 & 0x40000u, q.two11, 0x2000u);
        // 2^19 iff the k-mer is countable (bit J of E) AND of this CTA's half: (x4 ^ other19) & (E bit J moved to bit 19)
        const uint32_t ej = (E << (19 - J)) & 0x80000u;
        uint32_t mine;
        asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(mine) : "r"(x4), "r"(q.other19), "r"(ej));
        acc |= smem_add_ret(addr, mad_hi(mine, inc13, 0u));
        count9_steps<J + 1>(Wl, Wh, Rl, Rh, E, q, acc);
    }
}
template <int J>
__device__ __forceinline__ void drain9_steps(uint32_t Wl, uint32_t Wh, uint32_t Rl, uint32_t Rh, uint32_t* h8, uint32_t* slab)
{
    if constexpr (J < 16) {
        drain16<8>(h8, slab, rep9_offset(rep9_x4<J>(Wl, Wh, Rl, Rh)) >> 2);
        drain9_steps<J + 1>(Wl, Wh, Rl, Rh, h8, slab);
    }
}

template <bool PACKED>
__global__ void __launch_bounds__(kCountThreads)
count9h_kernel(const StepArgs* __restrict__ sa, PackedSrc pk, const uint64_t* __restrict__ sorted, Plan* __restrict__ plan,
               uint32_t* __restrict__ slabs, uint32_t lanes_policy)
{
    pdl_wait();
    const uint4* __restrict__ text16 = reinterpret_cast<const uint4*>(sa->text);
    const int breaklen = sa->pa.p.breaklength;
    // lanes_policy 3: tell the host when the sample would have been one for countt9_kernel (vk_countt9.cuh)
    if (lanes_policy == 3u && blockIdx.x == 0 && threadIdx.x == 0 && countu_wanted(plan, breaklen, 2u)) atomicOr(&plan->lanes_verdict, 1u);
    const uint32_t one = (uint32_t)(sa->n_bytes >> 62) + 1u;          // 1 (texts are shorter than 2^40), but not to ptxas
    constexpr int K = 9;
    constexpr uint32_t NB = 65536u;               // bins of one half
    constexpr uint32_t FULL = 0xffffffffu;
    extern __shared__ uint32_t s_raw9[];          // [pad to a 2 KiB shared address][32768 words: two 16-bit bins each]
    const uint32_t tid = threadIdx.x, lane = tid & 31, nthr = blockDim.x;
    const uint32_t pair = logical_cta() >> 1, half = logical_cta() & 1u;
    const uint32_t raw_addr = (uint32_t)__cvta_generic_to_shared(s_raw9);
    const uint32_t h_addr = (raw_addr + 2047u) & ~2047u;
    uint32_t* const s_raw = s_raw9 + ((h_addr - raw_addr) >> 2);

    uint32_t* const slab = slabs + (size_t)logical_cta() * NB;
    for (uint32_t i = tid; i < NB; i += nthr) slab[i] = 0;             // also for CTAs without a segment (reduce reads all)
    const int seg = cta_segment(plan, lane, pair);
    if (seg < 0) return;
    const uint32_t seg_len = (uint32_t)(plan->seg_reads[seg] < plan->seg_cap[seg] ? plan->seg_reads[seg] : plan->seg_cap[seg]);
    uint32_t* const h8 = s_raw;
    for (uint32_t i = tid; i < 32768u; i += nthr) s_raw[i] = 0;
    __syncthreads();

    ChunkStream<PACKED> cs;
    cs.init(text16, pk, sorted + plan->seg_begin[seg], seg_len, half ? &plan->seg_next2[seg] : &plan->seg_next[seg], lane,
            plan->n_bytes, (plan->seg_cta_begin[seg + 1] - plan->seg_cta_begin[seg]) * (blockDim.x >> 5));
    const uint32_t zero = (uint32_t)(plan->n_bytes >> 63);                   // 0 (texts are shorter than 2^40), but not to ptxas
    const Opaque9 q = {0x80000000u >> zero, 2048u >> zero, h_addr, (1u - half) << 19};
    uint32_t carry = 0;
    Chunk cur = cs.fetch();
    while (__ballot_sync(FULL, cur.range != 0) != 0) {
        const Chunk nxt = cs.fetch();
        const Decoded d = decode_chunk<K, PACKED>(cur, carry, lane, breaklen, one);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            // 24 bases: the 9-mers that end at chunk positions 16h .. 16h+15 (8 bases in front of the first one)
            const uint64_t W = h ? (uint64_t)(d.Plo >> 16) | ((uint64_t)d.Phi << 16) : (uint64_t)d.Cc | ((uint64_t)d.Plo << 16);
            const uint64_t W4 = W << 2;
            const uint64_t R4 = (revcomp_groups64(W) >> 14) << 2;      // rc of the 9-mer ending at j: R4 >> (32 - 2j)
            const uint32_t Wl = (uint32_t)W4, Wh = (uint32_t)(W4 >> 32), Rl = (uint32_t)R4, Rh = (uint32_t)(R4 >> 32);
            const uint32_t E = h ? d.E >> 16 : d.E & 0xFFFFu;
            uint32_t acc = 0;
            count9_steps<0>(Wl, Wh, Rl, Rh, E, q, acc);
            if (__ballot_sync(FULL, (acc & 0xC000u) != 0) != 0) {          // rare: some word is running hot
                if (acc & 0xC000u) drain9_steps<0>(Wl, Wh, Rl, Rh, h8, slab);
            }
        }
        cur = nxt;
    }
    __syncthreads();
    count16_flush<8>(h8, nullptr, slab, tid, nthr);
}

// segment histograms of the k = 9 pairs: class totals at the representative's forward index, 0 at its reverse complement
__global__ void __launch_bounds__(256)
reduce_slabs9h_kernel(const uint32_t* __restrict__ slabs, const Plan* __restrict__ plan, unsigned long long* __restrict__ seg_hist)
{
    pdl_wait();
    constexpr uint32_t NK = 1u << 18;
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (uint64_t)kMaxLevels * NK) return;
    const uint32_t s = (uint32_t)(g / NK), x = (uint32_t)(g % NK);
    unsigned long long sum = 0;
    if ((x & 0x200u) == 0) {
        const uint32_t id = (x & 0x1FFu) | ((x >> 10) << 9);
        const uint32_t half = id >> 16, bin = id & 0xFFFFu;
        const uint32_t c0 = plan->seg_cta_begin[s], c1 = plan->seg_cta_begin[s + 1];
        for (uint32_t c = c0; c < c1; ++c) sum += slabs[(size_t)(2 * c + half) * 65536u + bin];
    }
    seg_hist[g] = sum;
}

// K3: per-segment histograms (uint64) = sum of the slabs of the CTAs that served the segment.
// grid covers kMaxLevels * 4^k bins; segments beyond the ladder are written as zero.
__global__ void __launch_bounds__(256)
reduce_slabs_kernel(const uint32_t* __restrict__ slabs, const Plan* __restrict__ plan, uint32_t nk,
                    unsigned long long* __restrict__ seg_hist, int zero_unused)
{
    pdl_wait();
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (uint64_t)kMaxLevels * nk) return;
    const uint32_t s = (uint32_t)(g / nk), i = (uint32_t)(g % nk);
    // zero_unused = 0 (the fused step): the rows beyond the ladder are left alone -- the fold only reads the ladder's rows
    if (!zero_unused && (int)s >= plan->n_levels) return;
    unsigned long long sum = 0;
    const uint32_t c0 = plan->seg_cta_begin[s], c1 = plan->seg_cta_begin[s + 1];
    for (uint32_t c = c0; c < c1; ++c) sum += slabs[(size_t)c * nk + i];
    seg_hist[g] = sum;
}

// read-sharded samples: the per-segment reads / bases and the two overflow flags of this shard ride in the all-reduce of
// the histograms, right behind the rows that are exchanged (pack before, unpack after: the Plan then holds sample-wide
// totals and "some shard overflowed" flags, and the host code that reads it does not care whether the sample was sharded)
constexpr uint32_t kShardTail = 2 * kMaxLevels + 3;
__global__ void __launch_bounds__(kShardTail <= 256 ? 256 : 512)
shard_tail_kernel(Plan* __restrict__ plan, unsigned long long* __restrict__ tail, int unpack)
{
    pdl_wait();
    const uint32_t i = threadIdx.x;
    if (i >= kShardTail) return;
    if (!unpack) {
        unsigned long long v;
        if (i < (uint32_t)kMaxLevels) v = plan->seg_reads[i] - plan->seg_extra[i];      // reads, not entries
        else if (i < 2u * kMaxLevels) v = plan->seg_bases[i - kMaxLevels];
        else v = i == 2u * kMaxLevels ? plan->table_overflow : (i == 2u * kMaxLevels + 1 ? plan->bucket_overflow : plan->count_overflow);
        tail[i] = v;
    } else {
        const unsigned long long v = tail[i];
        if (i < (uint32_t)kMaxLevels) { plan->seg_reads[i] = v; plan->seg_extra[i] = 0; }
        else if (i < 2u * kMaxLevels) plan->seg_bases[i - kMaxLevels] = v;
        else if (i == 2u * kMaxLevels) plan->table_overflow = v ? 1u : 0u;
        else if (i == 2u * kMaxLevels + 1) plan->bucket_overflow = v ? 1u : 0u;
        else plan->count_overflow = v ? 1u : 0u;
    }
}

__global__ void __launch_bounds__(256)
zero_u64_kernel(unsigned long long* __restrict__ p, uint64_t n)
{
    pdl_wait();
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) p[g] = 0;
}

}  // namespace vk
