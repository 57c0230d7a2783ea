// vk_countt.cuh -- K2t: k = 7 counted in 8-mer PAIRS, ONE READ PER LANE, the text staged by asynchronous copies.
//
// Stands in for `dsk -kmer-size 7 -abundance-min 1` on every sub-sample file (varKoder/commands/image.py:771-796) exactly
// as count_kernel / count16_kernel / countu_kernel do (vk_count.cuh: rules D1-D8, forward-strand histogram per ladder
// segment, fold later); bins, increments, checksum and flush are countu_kernel's, the slab it writes is interchangeable.
//
// 106.6 us per 200 Mbp = 0.605 of the measured HBM roofline (the flat-lane kernel: 139 us, countu_kernel: 188 us); what was
// measured on the way is in profiles/r02c_notes.md, the design in DESIGN.md "K2".
//
// What is different from countu_kernel (vk_countu.cuh):
//  * STAGING costs no classification, no register traffic and little arithmetic: the 16-byte text words of a unit of 32
//    reads go from global to shared memory with cp.async (LDGSTS, the data never passes through registers), in read-major
//    order (consecutive lanes copy consecutive words of one read: about three reads per request).  Which word of which
//    read a lane copies in round n is fixed for the whole kernel (the stride is the sample's), so a round is two shuffles
//    (the read's address), an add and the copy.  The buffer of a warp is refilled in two HALVES (words [0, H) and [H, S) of
//    every read), each while the other is being counted: one buffer per warp, 16 warps per SM.  No L2 prefetch (it made
//    every line travel twice and held the kernel at the flat-lane kernel's time).
//  * The lane that owns a read classifies ITS words when it counts them (one LDS.128 per 16 bases): the SIMD
//    classification of vk_count.cuh, once per text word, nothing stored back.
//  * Words are aligned to the read's 7-MERS, not to its bases: word v holds the sixteen 7-mers that END at bases
//    16v + 6 .. 16v + 21 (window = bases 16v .. 16v + 21).  A read of 150 bases is then nine words of eight pairs and
//    nothing else: no first word without its six leading ends, no short last word.
//  * COUNT FIRST, LOOK AFTERWARDS: every lane adds the eight pairs of a word unconditionally (the increment is one LOP3);
//    a word with an N is queued with a flag and the drain takes the pairs that should not have counted back (shared-memory
//    arithmetic is modular).  Words that reach beyond the shortest read of the unit (other lengths; the odd 7-mer of a
//    151-base read; reads with cut points) are not pre-counted and take the queue with their exact masks.
//  * The queue entry is two words (window low | window high + the 16 "a 7-mer ends here" bits), filled under one
//    ballot; drained 32 at a time by countu's rule (complete pairs to the 8-mer bins, widowed 7-mers to the slab).
//  * Units are claimed from a counter in SHARED memory (the CTAs of a segment take equal contiguous shares).
//  * Large texts (the EPOCHS form): the 16-bit table goes to the slab every kTEpochUnits units of a CTA.
#pragma once
#include "vk_countu.cuh"

namespace vk {

constexpr uint32_t kTQuads = 352;                       // 16-byte text words a warp stages per unit
constexpr uint32_t kTBufWords = (kTQuads + 2u) * 4u;    // + one quad of slack on either side (read, never used)
constexpr uint32_t kTEpochUnits = 1536;                 // units of a CTA between two flushes of its table (3.5 M pairs: a word holds 2^15)
constexpr uint32_t kTQueue = 64;                        // irregular words a warp can hold (drained 32 at a time)
constexpr uint32_t kTRoundsF = 8, kTRoundsS = 6;        // copy rounds of the first / second half of a unit (at most 256 / 192 pieces)
constexpr uint32_t kNoPiece = 0x80000000u | (kTQuads << 22);      // a copy round in which the lane has no word to move: 16 zero bytes to the slack quad behind the buffer
template <int NW>
constexpr uint32_t countt_words() { return 32768u + (uint32_t)NW * kTBufWords + (uint32_t)NW * 2u * kTQueue; }
template <int NW>
constexpr size_t countt_smem_bytes() { return (size_t)countt_words<NW>() * sizeof(uint32_t); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void red_global_add(uint32_t* p, uint32_t v) { asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// x >> n on the FMA pipe (IMAD.HI): the loop is bound by the ALU pipe (LOP3 / SHF / PRMT), the FMA pipe idles
// (mul = 2^(32 - n) in a register ptxas cannot see through: a constant power of two comes back as LEA.HI, an ALU-pipe shift)
template <bool FMA, int N>
__device__ __forceinline__ uint32_t shr_fma(uint32_t x, uint32_t mul, uint32_t zero)
{
    return FMA ? mad_hi_u32(x, mul, zero) : x >> N;
}
// classify4z (vk_count.cuh) with its one shift as a multiply: t11 = 0x11 in every byte that holds a 'T' code
template <bool FMA>
__device__ __forceinline__ Cls4z classify4t(uint32_t x, uint32_t one, uint32_t zero, uint32_t m29)
{
    const uint32_t y = x & 0x06060606u;
    const uint32_t t8 = (y * 3u) & 0x08080808u;
    const uint32_t t11 = mad_lo_op(t8, one + one, shr_fma<FMA, 3>(t8, m29, zero));           // t8 << 1 | t8 >> 3
    uint32_t d1, dm;
    asm("lop3.b32 %0, %1, %2, %3, 0x6A;" : "=r"(d1) : "r"(x), "r"(0xD9D9D9D9u), "r"(0x41414141u));      // (x & M) ^ C
    asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(dm) : "r"(d1), "r"(t11), "r"(0x7F7F7F7Fu));            // (d1 ^ t11) & 0x7F..
    const uint32_t sv = mad_lo_op(dm, one, 0x7F7F7F7Fu);                                                 // + 0x7F.. on the FMA pipe
    Cls4z c;
    asm("lop3.b32 %0, %1, %2, %3, 0x02;" : "=r"(c.z) : "r"(sv), "r"(x), "r"(0x80808080u));             // ~(sv | x) & 0x80..
    c.packed_hi = y * 0x00820820u;
    return c;
}
template <bool FMA>
__device__ __forceinline__ uint32_t gather8t(uint32_t za, uint32_t zb, uint32_t zero, uint32_t m28)
{
    // (zb | za >> 4) * M as zb * M + (za >> 4) * M: the OR of disjoint bits is an add, and an add folds into the multiply (FMA pipe)
    return mad_lo_op(zb, 0x00204081u, shr_fma<FMA, 4>(za, m28, zero) * 0x00204081u);
}

template <int NW, int FM, bool EPOCHS>
__global__ void __launch_bounds__(NW * 32, 1)
countt_kernel(const StepArgs* __restrict__ sa, const uint64_t* __restrict__ sorted, Plan* __restrict__ plan,
              uint32_t* __restrict__ slabs, uint32_t policy)
{
    pdl_wait();
    constexpr int K = 7;
    constexpr uint32_t NK = 1u << (2 * K);
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr bool FMA_CLS = (FM & 1) != 0, FMA_PAIR = (FM & 2) != 0;      // which shifts are multiplies (measured: none is fastest)
    const uint8_t* __restrict__ text = sa->text;
    const int breaklen = sa->pa.p.breaklength;
    const uint32_t knobs = policy >> 8;          // bit 0: this kernel is the only one launched (a refusal must be reported); bits 4..7: H, bits 8..: units per epoch / 16 (experiments and tests, VK_COUNTT_KNOBS)
    policy &= 0xFFu;
    if (!countu_wanted(plan, breaklen, policy)) {
        if ((knobs & 1u) && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&plan->lanes_verdict, 2u);      // nobody else counts this step: say so
        return;
    }
    const uint32_t zero = (uint32_t)(sa->n_bytes >> 62);              // 0 (texts are shorter than 2^40), but not to ptxas
    const uint32_t one = zero + 1u;
    const uint32_t m29 = (1u << 29) + zero, m28 = (1u << 28) + zero, m24 = (1u << 24) + zero, m20 = (1u << 20) + zero;      // shifts by 3, 4, 8, 12 as multiplies
    const uint64_t total_quads = (sa->n_bytes + 15u) >> 4;
    extern __shared__ __align__(16) uint32_t s_rawt[];
    uint32_t* const s_raw = s_rawt;
    __shared__ unsigned long long s_chk[2];
    __shared__ uint32_t s_next;                  // next unclaimed read of this CTA's share of the segment
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;

    const int seg = cta_segment(plan, lane);
    if (seg < 0) return;
    const uint32_t seg_len = (uint32_t)(plan->seg_reads[seg] < plan->seg_cap[seg] ? plan->seg_reads[seg] : plan->seg_cap[seg]);
    const uint64_t* __restrict__ seg_sorted = sorted + plan->seg_begin[seg];
    // text words per read (odd: lane r reads at r * S, conflict-free) and reads per unit
    const uint32_t S = ((plan->len_max + 30u) >> 4) | 1u;
    const uint32_t R = kTQuads / S >= 32u ? 32u : kTQuads / S;
    // the buffer is refilled in two halves, words [0, H) and [H, S) of every read, each while the other is being counted
    // H: the second half (asked for when the unit starts) is needed from word H - 5 on, the first half of the next unit is
    // asked for then and has the remaining words to arrive: H = (S + 7) / 2 gives both about the same lead
    uint32_t H = (S + 7u) >> 1;
    if (H > S - 1u) H = S - 1u;
    if (R && H > 256u / R) H = 256u / R;
    if ((knobs >> 4) & 15u) H = (knobs >> 4) & 15u;

    uint32_t* const h8 = s_raw;
    const uint32_t h8_addr = (uint32_t)__cvta_generic_to_shared(s_raw);
    const uint32_t buf_addr = h8_addr + (32768u + warp * kTBufWords + 4u) * 4u;      // quad 0 of the warp's buffer
    uint32_t* const slab = slabs + (size_t)logical_cta() * NK;
    if (tid < 2) s_chk[tid] = 0;
    // the CTAs of a segment take equal contiguous shares of its sorted reads; inside a CTA the warps claim units of R reads
    // from a counter in shared memory (a counter per segment in global memory: 2368 warps x 18 claims on one address,
    // measured 25 % of all stall samples on the shuffle that reads the claim's answer)
    const uint32_t n_ctas_seg = plan->seg_cta_begin[seg + 1] - plan->seg_cta_begin[seg];
    const uint32_t per_cta = (seg_len + n_ctas_seg - 1u) / n_ctas_seg;
    const uint64_t lo64 = (uint64_t)(logical_cta() - plan->seg_cta_begin[seg]) * per_cta;
    const uint32_t cta_lo = lo64 < seg_len ? (uint32_t)lo64 : seg_len;
    const uint32_t cta_hi = seg_len - cta_lo < per_cta ? seg_len : cta_lo + per_cta;
    if (tid == 0) s_next = cta_lo;
    __syncthreads();
    if (R == 0u) {                                // a read too long for a staging buffer (forced mode only: countu_wanted keeps such samples away)
        if (tid == 0) atomicOr(&plan->count_overflow, 1u);      // -> the host repeats the step with the exact flat-lane kernel
        return;
    }

    // ---- units of R reads are claimed three units ahead of their use (the entries of a unit are needed when the unit
    // before it is half done: that is when its first copies are issued)
    auto entry_at = [&](uint32_t base) -> uint64_t {
        return (lane < R && base < cta_hi && base + lane < cta_hi) ? __ldg(seg_sorted + base + lane) : 0ull;
    };
    auto clamp_hi = [&](uint32_t b) -> uint32_t { return b < cta_hi ? b : cta_hi; };      // (shares are below 2^32 - 2^20: claims do not wrap)
    uint32_t claim0 = 0;
    if (lane == 0) claim0 = atomicAdd(&s_next, 3u * R);
    uint32_t baseA = clamp_hi(__shfl_sync(FULL, claim0, 0));
    uint32_t baseB = clamp_hi(baseA + R), baseC = clamp_hi(baseA + 2u * R);
    uint64_t entA = entry_at(baseA), entB = entry_at(baseB), entC = entry_at(baseC);
    uint32_t pending = 0;                                              // lane 0: the claim whose answer is read a unit later
    if (lane == 0) pending = atomicAdd(&s_next, R);
    // the tables are cleared while the first entries are on their way
    for (uint32_t i = tid; i < 8192u; i += nthr) reinterpret_cast<uint4*>(s_raw)[i] = make_uint4(0, 0, 0, 0);
    for (uint32_t i = tid; i < NK / 4u; i += nthr) reinterpret_cast<uint4*>(slab)[i] = make_uint4(0, 0, 0, 0);      // singles and the final fold ADD to the slab
    __syncthreads();

    // copy round n of a half: this lane moves word w of read r to quad r * S + w -- fixed for the whole kernel.
    // packed: r (5 bits) | w * 16 (13 bits, byte offset in the text) << 5 | (r * S + w) * 16 (13 bits, byte offset in the buffer) << 18
    uint32_t rwF[kTRoundsF], rwS[kTRoundsS];
    {
        const uint32_t S2 = S - H;
        const uint32_t invH = ((1u << 20) + H - 1u) / H, invS = S2 ? ((1u << 20) + S2 - 1u) / S2 : 0u;      // i / d = i * inv >> 20 for i < 416
#pragma unroll
        for (uint32_t n = 0; n < kTRoundsF; ++n) {
            const uint32_t i = 32u * n + lane;
            const uint32_t r1 = (i * invH) >> 20, w1 = i - r1 * H;
            rwF[n] = r1 < R ? (r1 | (w1 << 9) | ((r1 * S + w1) << 22)) : kNoPiece;
        }
#pragma unroll
        for (uint32_t n = 0; n < kTRoundsS; ++n) {
            const uint32_t i = 32u * n + lane;
            const uint32_t r2 = (i * invS) >> 20, w2 = i - r2 * S2 + H;
            rwS[n] = (S2 && r2 < R) ? (r2 | (w2 << 9) | ((r2 * S + w2) << 22)) : kNoPiece;
        }
    }
    const uint8_t* const text_end = text + (total_quads << 4);
    auto issue_half = [&](uint64_t ent, const auto& rw) {
        constexpr uint32_t NR = sizeof(rw) / sizeof(rw[0]);
        const uint32_t len = (uint32_t)(ent & kEntryLenMask);
        VK_ASSERT(!len || (ent >> kEntryLenBits) + len <= sa->n_bytes);            // a sorted-table entry names bytes of the text
        const uint8_t* const p0 = len ? text + ((ent >> kEntryLenBits) & ~15ull) : text_end;   // the read's first text word
        const uint32_t plo = (uint32_t)(uintptr_t)p0, phi = (uint32_t)((uintptr_t)p0 >> 32);
#pragma unroll
        for (uint32_t n = 0; n < NR; ++n) {
            const uint32_t wb = (rw[n] >> 5) & 0x1FF0u, db = (rw[n] >> 18) & 0x1FF0u;
            const uint32_t blo = __shfl_sync(FULL, plo, (int)rw[n]), bhi = __shfl_sync(FULL, phi, (int)rw[n]);      // (the shuffle looks at the low five bits: r)
            const uint8_t* const src = reinterpret_cast<const uint8_t*>(((uint64_t)bhi << 32) | blo) + wb;
            const bool ok = src < text_end && (int32_t)rw[n] >= 0;
            VK_ASSERT(db <= kTQuads * 16u);                              // quads 0 .. kTQuads - 1 of the warp's buffer, or the slack quad behind it
            VK_ASSERT(!ok || (src >= text && src + 16 <= text_end));      // the source word lies inside the text allocation
            cp_async16(buf_addr + db, ok ? src : text, ok ? 16u : 0u);
        }
        cp_async_commit();
    };
    // (No L2 prefetch: asking the lines of the units ahead into L2 made the kernel 20 % slower -- 600 instead of 410 MB of
    // DRAM reads, the copies' own lead is enough: profiles/r02c_notes.md.)

    // ---- the warp's queue of irregular words (window low | window high + E << 16)
    uint32_t* const qx = s_raw + 32768u + NW * kTBufWords + warp * (2u * kTQueue);
    uint32_t* const qy = qx + kTQueue;
    const uint32_t idle_x = (lane & 15u) * 0x11111111u;
    uint32_t qn = 0;                                                    // entries queued (warp-uniform)
    uint32_t made = 0;                                                  // pair increments of this lane
    // Bins: 8-mer b lives in word b & 0x7FFF; an increment adds 1 (bit 15 of b clear) or 0x20001 (set): the low half of a
    // word is the total of its two bins, the high half TWICE the upper bin -- the increment is one LOP3, (sh & 0x20000) | 1.
    // A word of a fast step is counted by every lane before anybody looks at its validity; the lanes whose word was
    // irregular queue it with the `pre` flag and the drain takes the pairs that should not have been counted back
    // (shared-memory arithmetic is modular: the order does not matter).
    auto drain = [&](uint32_t first, uint32_t n) {
        __syncwarp();
        const bool have = lane < n;
        const uint32_t Xl = have ? qx[first + lane] : idle_x;
        const uint32_t qyv = have ? qy[first + lane] : 0u;
        const uint32_t Xh = have ? (qyv & 0x3FFFu) : idle_x;
        const uint32_t pre = (qyv >> 15) & 1u;
        const uint32_t E = qyv >> 16;
        const uint32_t Ee = E & 0x5555u, Eo = (E >> 1) & 0x5555u;
        const uint32_t Eb = Ee & Eo;                                    // bit 2m: pair m complete -> one 8-mer
        uint32_t Es = Ee ^ Eo;                                          // bit 2m: pair m holds exactly one 7-mer
        made += __popc(Eb) - 8u * pre;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const uint32_t sh = __funnelshift_r(Xl, Xh, 4 * m);
            const uint32_t f = ((Eb >> (2 * m)) & 1u) - pre;            // +1: count it now, -1: take it back, 0: leave it
            smem_add(h8_addr + (sh & 0x1FFFCu), f * and_or(sh, 0x20000u, 1u));
        }
        while (Es != 0u) {                                              // a few lanes, once or twice
            const uint32_t b2 = __ffs(Es) - 1;                          // = 2m
            Es &= Es - 1;
            const uint32_t sh = __funnelshift_r(Xl, Xh, 2 * b2);
            const bool second = (Eo >> b2) & 1u;                        // the 7-mer that ends at 2m + 1: the 8-mer's last 7 bases
            atomicAdd(slab + (((second ? sh >> 2 : sh) & 0xFFFCu) >> 2), 1u);
        }
        __syncwarp();
    };

    // codes (32 bits) and validity (16 bits) of one staged text word
    auto classify16 = [&](const uint4 q, uint32_t& P, uint32_t& V) {
        const Cls4z c0 = classify4t<FMA_CLS>(q.x, one, zero, m29), c1 = classify4t<FMA_CLS>(q.y, one, zero, m29), c2 = classify4t<FMA_CLS>(q.z, one, zero, m29), c3 = classify4t<FMA_CLS>(q.w, one, zero, m29);
        P = __byte_perm(__byte_perm(c0.packed_hi, c1.packed_hi, 0x0073), __byte_perm(c2.packed_hi, c3.packed_hi, 0x0073), 0x5410);
        V = __byte_perm(gather8t<FMA_CLS>(c0.z, c1.z, zero, m28), gather8t<FMA_CLS>(c2.z, c3.z, zero, m28), 0x0073) & 0xFFFFu;
    };

    // Per unit: the second half of its text is asked for when the unit before it is done (the first half arrived while
    // that unit's last words were counted), waited for when the words in flight reach into it (v + 5 >= H: a lane reads two
    // quads ahead of the two words of an iteration), and from then on no lane reads below H any more: the first half of
    // the NEXT unit is asked for after that iteration.
    // The table goes to the slab (7-mer x: 8-mers that start with it, x | c << 14, + 8-mers that end with it, (x << 2 | c) &
    // 0xFFFF; x and x | 0x2000 end the same four words (x & 0x1FFF) << 2 | c: one 16-byte load serves both -- and, summed over x,
    // is the checksum: every increment added 1 to the low half of its word, so the low halves must sum to the increments made)
    // at the end of the CTA and, in the form for large texts (EPOCHS), after every kTEpochUnits units: 16-bit bins hold a CTA's
    // share of a 200 Mbp sample many times over, but a 15 Gbp shard puts 50 M pairs into the 2^15 words and a 7-mer at ten
    // times the mean wraps one (measured: every step of BASELINE configs[4] on two GPUs fell back to the exact kernel).
    auto flush_table = [&]() {
        uint32_t big = 0;
        unsigned long long low = 0;
        for (uint32_t x = tid; x < NK / 2u; x += nthr) {
            const uint4 q = reinterpret_cast<const uint4*>(h8)[x];
            const uint32_t up = (q.x >> 17) + (q.y >> 17) + (q.z >> 17) + (q.w >> 17);
            const uint32_t all = (q.x & 0xFFFFu) + (q.y & 0xFFFFu) + (q.z & 0xFFFFu) + (q.w & 0xFFFFu);
            big |= q.x | q.y | q.z | q.w;
            low += all;
            const uint32_t x1 = x | 0x2000u;
            red_global_add(slab + x, (h8[x] & 0xFFFFu) + (h8[x | 0x4000u] & 0xFFFFu) + (all - up));
            red_global_add(slab + x1, (h8[x1] & 0xFFFFu) + (h8[x1 | 0x4000u] & 0xFFFFu) + up);
        }
#pragma unroll
        for (int dlt = 16; dlt > 0; dlt >>= 1) low += __shfl_xor_sync(FULL, low, dlt);
        if (lane == 0) atomicAdd(&s_chk[1], low);
        // a word whose total reached 2^15 may have wrapped its high half (twice the upper bin): exact recount
        if (big & 0x8000u) atomicOr(&plan->count_overflow, 1u);
    };
    const uint32_t epoch_reads = ((knobs >> 8) ? (knobs >> 8) * 16u : kTEpochUnits) * R;      // (tests: short epochs)
    issue_half(entA, rwF);
    // epochs: every warp leaves the unit loop when the unit it holds lies beyond the epoch's end (a warp that has run out of
    // units leaves at once), the queue is applied, the table goes to the slab and is cleared, the next epoch begins
    for (uint32_t epoch_end = (EPOCHS && cta_hi - cta_lo > epoch_reads) ? cta_lo + epoch_reads : cta_hi;;) {
    while (baseA < epoch_end) {
        // ---- the claim made a unit ago has its answer; the entries of that unit are asked for now (used three units on)
        const uint32_t baseD = clamp_hi(__shfl_sync(FULL, pending, 0));
        const uint64_t entD = entry_at(baseD);
        if (lane == 0 && baseD < cta_hi) pending = atomicAdd(&s_next, R);
        issue_half(entA, rwS);                                          // (the unit before this one is done with those quads)
        cp_async_wait<1>();                                             // first half of this unit
        __syncwarp();
        // ---- this lane's read
        const uint32_t len = (uint32_t)(entA & kEntryLenMask);
        const uint32_t o = (uint32_t)(entA >> kEntryLenBits) & 15u;
        const bool active = len != 0u;
        const bool brk = breaklen > 0 && len > (uint32_t)breaklen;
        const uint32_t lmax = __reduce_max_sync(FULL, len);
        const uint32_t lmin = __reduce_min_sync(FULL, active ? len : 0xFFFFFFFFu);
        const uint32_t nW = lmax >= (uint32_t)K ? (lmax - 6u + 15u) >> 4 : 0u;      // words of the longest read
        bool waited = false, filled = false;
        if (H < 4u) { cp_async_wait<0>(); __syncwarp(); waited = true; }           // (the first four quads reach into the second half)
        if (nW != 0u) {
            // words below vfast lie wholly inside every read of the unit (16 v + 22 <= lmin); none when a read has cut points
            const uint32_t vfast = (__any_sync(FULL, brk) || lmin < 22u) ? 0u : (lmin - 6u) >> 4;
            const uint32_t s = (o + 6u) & 15u, tb = (o + 6u) >> 4;
            // text word T of the read sits at quad lane * S + T; word v needs T = v + tb and v + tb + 1
            uint32_t qa = buf_addr + ((lane < R ? lane : 0u) * S + tb) * 16u;      // (a lane beyond the unit walks read 0: nothing it sees counts)
            uint32_t Pp, Vp, P, V;
            classify16(lds128(qa - 16u), Pp, Vp);                       // T = tb - 1 (tb = 0: the slack quad, no bit of it is used)
            classify16(lds128(qa), P, V);                               // T = tb
            uint32_t C = __funnelshift_r(Pp, P, 2u * s);
            uint32_t Vv = ((Vp | (V << 16)) >> s) & 0xFFFFu;
            uint32_t Cc = C >> 20, Vc = Vv >> 10;                       // bases 0..5 of the read: context of word 0
            Pp = P;
            Vp = V;
            const uint32_t inc_mask = active ? 0x20000u : 0u, inc_one = active ? 1u : 0u;
            // eight pair increments of a fast word: counted first (every lane that holds a read), looked at afterwards.
            // Pairs 0..3 lie in the low word of the window, pairs 4..7 in the word that starts at its bit 16
            auto count8 = [&](const uint32_t Xl, const uint32_t C) {
                const uint32_t X2 = C >> 2;                              // the window from its bit 16 on: (Cc << 2 | C << 14) >> 16
                uint32_t sh[8];
                sh[0] = Xl; sh[1] = shr_fma<FMA_PAIR, 4>(Xl, m28, zero); sh[2] = shr_fma<FMA_PAIR, 8>(Xl, m24, zero); sh[3] = shr_fma<FMA_PAIR, 12>(Xl, m20, zero);
                sh[4] = X2; sh[5] = shr_fma<FMA_PAIR, 4>(X2, m28, zero); sh[6] = shr_fma<FMA_PAIR, 8>(X2, m24, zero); sh[7] = shr_fma<FMA_PAIR, 12>(X2, m20, zero);
                uint32_t ad[8], in[8];
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    ad[m] = h8_addr + (sh[m] & 0x1FFFCu);
                    in[m] = and_or(sh[m], inc_mask, inc_one);
                }
                asm volatile("red.shared.add.u32 [%0], %8;\n\tred.shared.add.u32 [%1], %9;\n\tred.shared.add.u32 [%2], %10;\n\t"
                             "red.shared.add.u32 [%3], %11;\n\tred.shared.add.u32 [%4], %12;\n\tred.shared.add.u32 [%5], %13;\n\t"
                             "red.shared.add.u32 [%6], %14;\n\tred.shared.add.u32 [%7], %15;"
                             :: "r"(ad[0]), "r"(ad[1]), "r"(ad[2]), "r"(ad[3]), "r"(ad[4]), "r"(ad[5]), "r"(ad[6]), "r"(ad[7]),
                                "r"(in[0]), "r"(in[1]), "r"(in[2]), "r"(in[3]), "r"(in[4]), "r"(in[5]), "r"(in[6]), "r"(in[7]) : "memory");
            };
            // is word v of this lane's read one for the queue?  fast: it was counted and is not clean; else: it holds a 7-mer
            auto wants_push = [&](const uint32_t v, const bool fast, uint32_t& VW) -> bool {
                if (fast) return active && VW != 0x3FFFFFu;
                const uint32_t b0w = 16u * v;
                const bool push = active && b0w + 6u < len;
                const uint32_t left = push ? len - b0w : 0u;
                VW &= left < 22u ? (1u << left) - 1u : 0x3FFFFFu;
                return push;
            };
            auto do_push = [&](const uint32_t v, const bool fast, const bool push, const uint32_t Xl, const uint32_t Xh, const uint32_t VW) {
                const uint32_t bal = __ballot_sync(FULL, push);
                if (bal == 0u) return;
                if (push) {
                    const uint32_t r2 = VW & (VW >> 1), r4 = r2 & (r2 >> 2);
                    uint32_t E = r4 & (r2 >> 4) & (VW >> 6) & 0xFFFFu;         // bit e: the 7-mer that ENDS at base 16 v + 6 + e is countable
                    if (brk) {
                        // reformat.sh breaklength: no 7-mer may span a multiple of breaklen counted from the read's first base
                        const int32_t qb = (int32_t)(16u * v) + 6;              // base of E's bit 0
                        int32_t c = (qb / breaklen) * breaklen;
                        if (c < breaklen) c = breaklen;
                        uint32_t dead = 0;
                        for (; c - qb < 16 && c < (int32_t)len; c += breaklen) {
                            const int32_t b = c - qb;                           // base that starts the new piece: ends b .. b+5 are dead
                            if (b > -(K - 1)) dead |= b >= 0 ? 0x3Fu << b : 0x3Fu >> (-b);
                        }
                        E &= ~dead;
                    }
                    const uint32_t slot = qn + __popc(bal & ((1u << lane) - 1u));
                    qx[slot] = Xl;
                    qy[slot] = Xh | (fast ? 0x8000u : 0u) | (E << 16);
                }
                qn += __popc(bal);
                if (qn >= 32u) { qn -= 32u; drain(qn, 32u); }
            };
            // Two words per iteration (their classifications are independent chains: the loop is short of warps, not of
            // work); an odd word count starts with a word on its own.  The lane reads two quads ahead: quads up to
            // tb + v + 4 are in flight before words v, v + 1, so the second half must have arrived when v + 5 >= H, and
            // from then on no lane reads below H: the first half of the next unit can be asked for.
            uint4 nq0 = lds128(qa + 16u), nq1 = lds128(qa + 32u);
            uint32_t v = 0;
            if (nW & 1u) {
                classify16(nq0, P, V);
                qa += 16u;
                nq0 = nq1;
                nq1 = lds128(qa + 32u);
                C = __funnelshift_r(Pp, P, 2u * s);
                Vv = ((Vp | (V << 16)) >> s) & 0xFFFFu;
                uint32_t VW = Vc | (Vv << 6);                           // bit b <-> base 16 v + b of the read, b = 0..21
                const uint32_t Xl = (Cc << 2) | (C << 14), Xh = C >> 18;
                const bool fast = 0u < vfast;
                if (fast) count8(Xl, C);
                const bool push = wants_push(0u, fast, VW);
                do_push(0u, fast, push, Xl, Xh, VW);
                Cc = C >> 20; Vc = Vv >> 10; Pp = P; Vp = V;
                v = 1;
            }
            for (; v < nW; v += 2) {
                if (!waited && v + 5u >= H) { cp_async_wait<0>(); __syncwarp(); waited = true; }      // second half of this unit
                uint32_t Pa, Va, Pb, Vb;
                classify16(nq0, Pa, Va);
                classify16(nq1, Pb, Vb);
                qa += 32u;
                nq0 = lds128(qa + 16u);                                 // (beyond the last word: the next read, the slack quad, the next buffer)
                nq1 = lds128(qa + 32u);
                const uint32_t Ca = __funnelshift_r(Pp, Pa, 2u * s), Cb = __funnelshift_r(Pa, Pb, 2u * s);
                const uint32_t Vva = ((Vp | (Va << 16)) >> s) & 0xFFFFu, Vvb = ((Va | (Vb << 16)) >> s) & 0xFFFFu;
                uint32_t VWa = Vc | (Vva << 6), VWb = (Vva >> 10) | (Vvb << 6);
                const uint32_t Xla = (Cc << 2) | (Ca << 14), Xha = Ca >> 18;
                const uint32_t Xlb = ((Ca >> 20) << 2) | (Cb << 14), Xhb = Cb >> 18;
                const bool fast_a = v < vfast, fast_b = v + 1u < vfast;
                if (fast_a) count8(Xla, Ca);
                if (fast_b) count8(Xlb, Cb);
                const bool push_a = wants_push(v, fast_a, VWa), push_b = wants_push(v + 1u, fast_b, VWb);
                if (__any_sync(FULL, push_a || push_b)) {
                    do_push(v, fast_a, push_a, Xla, Xha, VWa);
                    do_push(v + 1u, fast_b, push_b, Xlb, Xhb, VWb);
                }
                Cc = Cb >> 20; Vc = Vvb >> 10; Pp = Pb; Vp = Vb;
                if (!filled && v + 5u >= H) { __syncwarp(); issue_half(entB, rwF); filled = true; }      // first half of the next unit
            }
            made += active ? 8u * vfast : 0u;
        }
        if (!waited) cp_async_wait<0>();
        __syncwarp();                                                   // every lane has read its words: the buffer may be overwritten
        if (!filled) issue_half(entB, rwF);
        entA = entB; entB = entC; entC = entD;
        baseA = baseB; baseB = baseC; baseC = baseD;
    }
        if (!EPOCHS || epoch_end >= cta_hi) break;
        if (qn != 0u) { drain(0u, qn); qn = 0u; }
        __syncthreads();
        flush_table();
        __syncthreads();
        for (uint32_t i = tid; i < 8192u; i += nthr) reinterpret_cast<uint4*>(s_raw)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        epoch_end = cta_hi - epoch_end < epoch_reads ? cta_hi : epoch_end + epoch_reads;
    }
    cp_async_wait<0>();
    if (qn != 0u) drain(0u, qn);
    __threadfence();                                                    // the singles' atomics have reached the slab
    __syncthreads();
    flush_table();
    {
        unsigned long long mine = (unsigned long long)(long long)(int32_t)made;      // (a lane drains other lanes' words: its own balance may be negative)
#pragma unroll
        for (int dlt = 16; dlt > 0; dlt >>= 1) mine += __shfl_xor_sync(FULL, mine, dlt);
        if (lane == 0) atomicAdd(&s_chk[0], mine);
        __syncthreads();
        if (tid == 0 && s_chk[0] != s_chk[1]) atomicOr(&plan->count_overflow, 1u);
    }
}

}  // namespace vk
