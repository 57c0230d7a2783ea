// vk_countu.cuh -- K2u: k = 7 counted in 8-mer PAIRS with ONE READ PER LANE, for samples whose reads all have (about) the
// same length (raw Illumina reads; BASELINE configs 1, 2, 3, 5).
//
// Stands in for `dsk -kmer-size 7 -abundance-min 1` on every sub-sample file (varKoder/commands/image.py:771-796) exactly
// as count_kernel / count16_kernel do (vk_count.cuh: rules D1-D8, forward-strand histogram per ladder segment, fold later);
// bins, increments, checksum and flush are count16_kernel<7, ., FAST>'s, so the slab it writes is interchangeable.
//
// Why another kernel.  The flat-lane kernels (vk_count.cuh) spend ~425 warp instructions per 32 chunks of 32 bytes, ~300
// of them on the ALU pipe (one warp instruction per two cycles and scheduler) and 32 x 3.5 shared-memory wavefronts on the
// increments: both pipes are level at ~100 us per 200 Mbp and the kernel takes 139 us.  Their bookkeeping (which read owns a
// chunk, validity runs, range masks, "count or trash" selects) is paid on every chunk because every 32-chunk window holds
// half a dozen read starts and ends.  Here a lane owns a read and REALIGNS it: the 2-bit codes of two consecutive
// 16-byte text words are funnel-shifted by the read's offset inside its first word (one SHF), so that "virtual word" j of a
// lane holds bases 16j .. 16j+15 of ITS read whatever the read's address is.  Reads of one length are then in the same
// phase in all 32 lanes: which pairs of a virtual word exist is a warp-uniform mask (bases 0..5 open no 7-mer; the last
// word is short), the increments are unpredicated, and pairs are aligned to the READ, so an N-free read of even 7-mer count
// has no single 7-mers at all: 72 increments for 144 7-mers.  What the uniform mask gets wrong in a lane -- a pair that
// holds an N or a cut point, a shorter read of the unit -- is put right afterwards in a loop most iterations skip: the
// increment is taken back (shared-memory arithmetic is modular, the order does not matter) and a 7-mer that lost its
// partner goes to the 4^7 x u32 table.
//
// Loads.  One read per lane must not mean one LOAD per lane: 32 lanes fetching 16 bytes from 32 different reads touch 32
// cache lines per request, keep two 16-byte words per lane in flight and leave the SM idle (measured: 27 % issue, 310 us).
// A warp therefore handles a unit of reads in two phases.  STAGE: the unit's text words are numbered read-major (word w of
// read r is piece r * stride + w); lane l takes pieces l, l + 32, ... -- consecutive lanes read consecutive 16-byte words of
// the same read, three reads per request -- classifies its word and stores the 32-bit codes and the 16 validity bits at
// the piece's index in the warp's staging area in shared memory (consecutive indices: no bank conflicts).  COUNT: lane r
// walks read r, word t at index r * stride + t (stride odd: no bank conflicts either).  The 16-byte loads of a stage are
// issued four at a time and the unit after next is pulled into L2 meanwhile.
//
// Work distribution: static.  The reads of a segment (sorted table, vk_bucket.cuh) are dealt to the warps that serve the
// segment in equal contiguous shares; nothing is claimed at run time.
//
// Shared memory: 128 KiB of 8-mer bins + 72 KiB of staging + 24 KiB of queues; the rare single 7-mers go straight to the
// CTA's slab in global memory (atomics on L2).
#pragma once
#include <type_traits>
#include "vk_count.cuh"

namespace vk {

constexpr uint32_t kQueue = 64;         // irregular words a warp can hold (drained 32 at a time)
constexpr uint32_t kStagePieces = 384;  // 16-byte text words a warp stages per unit (32 reads of up to 176 text bytes)
constexpr uint32_t kCountuWarps = kCountThreads / 32;
// dynamic shared memory in 32-bit words: 8-mer bins | codes of the staged words | their validity bits (16 each) | queues
constexpr uint32_t kCuCodes = 32768;
constexpr uint32_t kCuValid = kCuCodes + kCountuWarps * kStagePieces;
constexpr uint32_t kCuQueue = kCuValid + kCountuWarps * kStagePieces / 2;
constexpr uint32_t kCuWords = kCuQueue + kCountuWarps * 3 * kQueue;
constexpr size_t countu_smem_bytes() { return (size_t)kCuWords * sizeof(uint32_t); }

// 16 validity bits of a classified text word (bit b: byte b is one of ACGTacgt)
__device__ __forceinline__ uint32_t valid16(const Cls4z& c0, const Cls4z& c1, const Cls4z& c2, const Cls4z& c3)
{
    return __byte_perm(gather8(c0.z, c1.z), gather8(c2.z, c3.z), 0x0073) & 0xFFFFu;
}

__global__ void __launch_bounds__(kCountThreads)
countu_kernel(const StepArgs* __restrict__ sa, const uint64_t* __restrict__ sorted, Plan* __restrict__ plan,
              uint32_t* __restrict__ slabs, uint32_t policy)
{
    pdl_wait();
    constexpr int K = 7;
    constexpr uint32_t NK = 1u << (2 * K);
    constexpr uint32_t FULL = 0xffffffffu;
    const uint4* __restrict__ text16 = reinterpret_cast<const uint4*>(sa->text);
    const int breaklen = sa->pa.p.breaklength;
    if (!countu_wanted(plan, breaklen, policy)) return;
    const uint32_t zero = (uint32_t)(sa->n_bytes >> 62);              // 0 (texts are shorter than 2^40), but not to ptxas
    const uint32_t one = zero + 1u, two31 = 0x80000000u >> zero;
    extern __shared__ uint32_t s_raw[];
    __shared__ unsigned long long s_chk[2];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;

    const int seg = cta_segment(plan, lane);
    if (seg < 0) return;
    const uint32_t seg_len = (uint32_t)(plan->seg_reads[seg] < plan->seg_cap[seg] ? plan->seg_reads[seg] : plan->seg_cap[seg]);
    const uint64_t* __restrict__ seg_sorted = sorted + plan->seg_begin[seg];
    // static shares: warp g of the segment's warps walks reads [g * per, (g + 1) * per)
    const uint32_t n_warps = (plan->seg_cta_begin[seg + 1] - plan->seg_cta_begin[seg]) * (nthr >> 5);
    const uint32_t gw = (logical_cta() - plan->seg_cta_begin[seg]) * (nthr >> 5) + warp;
    const uint32_t per = (seg_len + n_warps - 1u) / n_warps;
    const uint64_t wb64 = (uint64_t)gw * per;
    const uint32_t wbase = wb64 < seg_len ? (uint32_t)wb64 : seg_len;
    const uint32_t wend = seg_len - wbase < per ? seg_len : wbase + per;
    // reads per unit: as many as the staging area holds words of the sample's longest read (32 up to 160 bases)
    const uint32_t cap_nt = ((plan->len_max + 30u) >> 4) | 1u;
    const uint32_t R = kStagePieces / cap_nt >= 32u ? 32u : kStagePieces / cap_nt;
    if (R == 0u) return;                          // (countu_wanted keeps such samples away; policy 1 with huge reads)
    const uint32_t n_units = (wend - wbase + R - 1u) / R;

    uint32_t* const h8 = s_raw;
    const uint32_t h8_addr = (uint32_t)__cvta_generic_to_shared(s_raw);
    uint32_t* const Ps = s_raw + kCuCodes + warp * kStagePieces;
    uint16_t* const Vs = reinterpret_cast<uint16_t*>(s_raw + kCuValid) + warp * kStagePieces;
    uint32_t* const slab = slabs + (size_t)logical_cta() * NK;
    for (uint32_t i = tid; i < 32768u; i += nthr) s_raw[i] = 0;
    for (uint32_t i = tid; i < NK; i += nthr) slab[i] = 0;             // singles and the final fold ADD to the slab
    if (tid < 2) s_chk[tid] = 0;
    __syncthreads();

    auto entry = [&](uint32_t u) -> uint64_t {
        const uint32_t idx = wbase + R * u + lane;
        return (u < n_units && lane < R && idx < wend) ? __ldg(seg_sorted + idx) : 0ull;
    };
    // lanes without a pair to count add 0 to words spread over the banks (every 4-bit group of the window = lane & 15:
    // the bank bits of all eight pair addresses then differ between lanes that differ in their low four bits)
    const uint32_t idle_x = (lane & 15u) * 0x11111111u;
    // ---- the warp's queue of irregular words (window low / high / countable mask)
    uint32_t* const qx = s_raw + kCuQueue + warp * (3u * kQueue);
    uint32_t* const qy = qx + kQueue;
    uint32_t* const qe = qy + kQueue;
    uint32_t qn = 0;                                                    // entries queued (warp-uniform)
    uint32_t made = 0;                                                  // pair increments of this lane
    // one queued word per lane: pairs whose two 7-mers are countable go to the 8-mer bins, the others' 7-mers to the
    // slab (count16_kernel<7> does this for every chunk; here it runs once per 32 irregular words)
    auto drain = [&](uint32_t first, uint32_t n) {
        __syncwarp();
        const bool have = lane < n;
        const uint32_t Xl = have ? qx[first + lane] : idle_x, Xh = have ? qy[first + lane] : idle_x;
        const uint32_t E = have ? qe[first + lane] : 0u;
        const uint32_t Ee = E & 0x5555u, Eo = (E >> 1) & 0x5555u;
        const uint32_t Eb = Ee & Eo;                                    // bit 2m: pair m complete -> one 8-mer
        uint32_t Es = Ee ^ Eo;                                          // bit 2m: pair m holds exactly one 7-mer
        made += __popc(Eb);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const uint32_t sh = __funnelshift_r(Xl, Xh, 4 * m);
            uint32_t inc = 0;
            if ((Eb >> (2 * m)) & 1u) inc = mad_hi_u32(sh & 0x20000u, two31, one);
            smem_add(h8_addr + (sh & 0x1FFFCu), inc);
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

    // the text of a unit is pulled into L2 two units ahead: lane l asks for the first and the last line of its read
    auto prefetch_l2 = [&](uint64_t ent) {
        const uint32_t len = (uint32_t)(ent & kEntryLenMask);
        if (len != 0u) {
            const uint64_t start = ent >> kEntryLenBits;
            const uint4* const p = text16 + (start >> 4);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (((uint32_t)start & 15u) + len - 1u) / 16u));
            if (len > 128u) for (uint32_t w = 8; w * 16u < len; w += 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + w));
        }
    };

    uint64_t entA = entry(0), entB = entry(1);
    prefetch_l2(entB);
    for (uint32_t u = 0; u < n_units; ++u) {
        const uint64_t entC = entry(u + 2u);
        prefetch_l2(entC);
        // ---- this lane's read
        const uint32_t len = (uint32_t)(entA & kEntryLenMask);
        const uint64_t start = entA >> kEntryLenBits;
        const uint32_t o = (uint32_t)start & 15u;
        const uint32_t nt = len ? (o + len + 15u) >> 4 : 0u;           // text words the read touches
        const uint32_t w16lo = (uint32_t)(start >> 4), w16hi = (uint32_t)(start >> 36);
        const bool brk = breaklen > 0 && len > (uint32_t)breaklen;
        const uint32_t lmax = __reduce_max_sync(FULL, len);
        const uint32_t lmin = __reduce_min_sync(FULL, len ? len : 0xFFFFFFFFu);
        const uint32_t stride = __reduce_max_sync(FULL, nt) | 1u;      // <= cap_nt: R * stride pieces fit
        const uint32_t n_rounds = (R * stride + 31u) >> 5;
        const uint32_t inv = ((1u << 20) + stride - 1u) / stride;      // i / stride = i * inv >> 20 for i < 416 (checked)

        // ---- STAGE: piece i = word (i % stride) of read (i / stride); lane l takes pieces l, l + 32, ...
        __syncwarp();
        for (uint32_t n0 = 0; n0 < n_rounds; n0 += 4) {
            uint4 q[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t i = 32u * (n0 + k) + lane;
                const uint32_t r = (i * inv) >> 20, w = i - r * stride;
                const uint32_t rr = r < 32u ? r : 31u;
                const uint32_t blo = __shfl_sync(FULL, w16lo, rr), bhi = __shfl_sync(FULL, w16hi, rr);
                const uint32_t ntr = __shfl_sync(FULL, nt, rr);
                const bool ok = n0 + k < n_rounds && r < 32u && w < ntr;
                q[k] = ok ? ldg_text(text16 + (((uint64_t)bhi << 32) | blo) + w) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (n0 + k < n_rounds) {
                    const uint32_t i = 32u * (n0 + k) + lane;
                    const Cls4z c0 = classify4z(q[k].x, one), c1 = classify4z(q[k].y, one), c2 = classify4z(q[k].z, one), c3 = classify4z(q[k].w, one);
                    Ps[i] = __byte_perm(__byte_perm(c0.packed_hi, c1.packed_hi, 0x0073), __byte_perm(c2.packed_hi, c3.packed_hi, 0x0073), 0x5410);
                    Vs[i] = (uint16_t)valid16(c0, c1, c2, c3);
                }
            }
        }
        __syncwarp();

        // ---- COUNT: lane r walks read r; virtual word j = bases 16j .. 16j + 15 = staged words j and j + 1 shifted by o
        if (lmax != 0u) {
            const uint32_t* const myP = Ps + lane * stride;
            const uint16_t* const myV = Vs + lane * stride;
            const uint32_t nv = (lmax + 15u) >> 4;                      // virtual words of the longest read
            const uint32_t jfull = lmin >> 4;                           // words 1 .. jfull - 1 lie wholly inside every read
            uint32_t Pp = nt > 0u ? myP[0] : 0u, Vp = nt > 0u ? myV[0] : 0u;
            uint32_t Cc = 0, Vc = 0;    // codes / validity bits of the six bases in front of the current virtual word
            // MODE 1: an INTERIOR word of every read of the unit (sixteen bases of the read in every lane that holds one,
            // eight pairs).  MODE 2: any word (the first: bases 0..5 end no 7-mer; the last ones: reads end); the pairs a
            // read of lmax bases has there are a warp-uniform mask.
            auto step = [&](auto mode_c, const uint32_t j) {
                constexpr int MODE = decltype(mode_c)::value;
                const uint32_t t = j + 1u;
                const uint32_t P = t < nt ? myP[t] : 0u, Vt = t < nt ? myV[t] : 0u;
                const uint32_t j16 = 16u * j;
                const uint32_t C = __funnelshift_r(Pp, P, 2u * o);
                uint32_t V = (Vp | (Vt << 16)) >> o;
                bool active, irregular;
                uint32_t um, VW;
                if constexpr (MODE == 1) {
                    active = len != 0u;
                    V &= 0xFFFFu;
                    VW = Vc | (V << 6);                                         // bit b <-> base b - 6 of the word
                    um = 0x5555u;
                    irregular = active && (VW != 0x3FFFFFu || brk);
                } else {
                    active = j16 < len;
                    const uint32_t left = len - j16;                            // bases of the read from this word on
                    V &= active ? (left < 16u ? (1u << left) - 1u : 0xFFFFu) : 0u;
                    VW = Vc | (V << 6);
                    // warp-uniform: the pairs a read of lmax bases has in this word -- pair m = the 7-mers that end at bases
                    // 2m and 2m + 1 of the word (read parity: the first 7-mer of a read ends at base 6)
                    const uint32_t lleft = lmax - j16;                          // > 0
                    const uint32_t mhi = lleft >= 16u ? 8u : lleft >> 1;
                    um = (0x5555u >> (16u - 2u * mhi)) & (j == 0u ? 0x5540u : 0x5555u);         // bit 2m: pair m is counted here
                    const uint32_t need = ((1u << (2u * mhi + 6u)) - 1u) & (j == 0u ? ~0x3Fu : ~0u);    // bases those pairs are made of
                    // a lane is REGULAR when its read has exactly those pairs in this word and no 7-mer beside them
                    irregular = active && ((VW & need) != need || ((VW >> (2u * mhi + 6u)) & 1u) != 0u || brk);
                }
                // window: six bases of context, then the sixteen of the word; times 4 (byte offsets into the table)
                uint32_t Xl = (Cc << 2) | (C << 14), Xh = C >> 18;
                if (__any_sync(FULL, irregular)) {
                    // rare (an N, a cut point, a shorter read, an odd number of 7-mers): the lane's word goes to the queue
                    // with the exact mask of its countable 7-mers
                    const uint32_t bal = __ballot_sync(FULL, irregular);
                    if (irregular) {
                        const uint32_t r2 = VW & (VW >> 1), r4 = r2 & (r2 >> 2);
                        uint32_t E = r4 & (r2 >> 4) & (VW >> 6) & 0xFFFFu;     // bit s: the 7-mer that ENDS at base 16j + s is countable
                        if (brk) {
                            // reformat.sh breaklength: no 7-mer may span a multiple of breaklen counted from the read's first base
                            const int32_t qb = (int32_t)j16;
                            int32_t c = (qb / breaklen) * breaklen;
                            if (c < breaklen) c = breaklen;
                            uint32_t dead = 0;
                            for (; c - qb < 16 && c < (int32_t)len; c += breaklen) {
                                const int32_t b = c - qb;                       // base that starts the new piece: ends b .. b+5 are dead
                                if (b > -(K - 1)) dead |= b >= 0 ? 0x3Fu << b : 0x3Fu >> (-b);
                            }
                            E &= ~dead;
                        }
                        const uint32_t slot = qn + __popc(bal & ((1u << lane) - 1u));
                        qx[slot] = Xl; qy[slot] = Xh; qe[slot] = E;
                    }
                    qn += __popc(bal);
                    if (qn >= 32u) { qn -= 32u; drain(qn, 32u); }
                }
                const bool counts = active && !irregular;
                if (!counts) { Xl = idle_x; Xh = idle_x; }
                const uint32_t mul = counts ? two31 : 0u, add = counts ? one : 0u;
                // the 8-mer that ends at base 2m + 1 starts at window position 2m; 1, or 0x10001 for the upper bin of the word
                if constexpr (MODE == 1) {
                    made += counts ? 8u : 0u;
                    uint32_t ad[8], in[8];
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        const uint32_t sh = __funnelshift_r(Xl, Xh, 4 * m);
                        ad[m] = h8_addr + (sh & 0x1FFFCu);
                        in[m] = mad_hi_u32(sh & 0x20000u, mul, add);
                    }
                    // one statement: sixteen live registers, so that no increment waits for a register of the one before
                    asm volatile("red.shared.add.u32 [%0], %8;\n\tred.shared.add.u32 [%1], %9;\n\tred.shared.add.u32 [%2], %10;\n\t"
                                 "red.shared.add.u32 [%3], %11;\n\tred.shared.add.u32 [%4], %12;\n\tred.shared.add.u32 [%5], %13;\n\t"
                                 "red.shared.add.u32 [%6], %14;\n\tred.shared.add.u32 [%7], %15;"
                                 :: "r"(ad[0]), "r"(ad[1]), "r"(ad[2]), "r"(ad[3]), "r"(ad[4]), "r"(ad[5]), "r"(ad[6]), "r"(ad[7]),
                                    "r"(in[0]), "r"(in[1]), "r"(in[2]), "r"(in[3]), "r"(in[4]), "r"(in[5]), "r"(in[6]), "r"(in[7]) : "memory");
                } else {
                    made += counts ? __popc(um) : 0u;
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        if ((um >> (2 * m)) & 1u) {
                            const uint32_t sh = __funnelshift_r(Xl, Xh, 4 * m);
                            smem_add(h8_addr + (sh & 0x1FFFCu), mad_hi_u32(sh & 0x20000u, mul, add));
                        }
                    }
                }
                Cc = C >> 20;
                Vc = V >> 10;
                Pp = P;
                Vp = Vt;
            };
            using M1 = std::integral_constant<int, 1>;
            using M2 = std::integral_constant<int, 2>;
            step(M2{}, 0u);
            uint32_t j = 1;
            for (; j < jfull && j < nv; ++j) step(M1{}, j);
            for (; j < nv; ++j) step(M2{}, j);
        }
        entA = entB;
        entB = entC;
    }
    if (qn != 0u) drain(0u, qn);
    __threadfence();                                                    // the singles' atomics have reached the slab
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
    // 7-mer x: 8-mers that start with it (x | c << 14) + 8-mers that end with it ((x << 2 | c) & 0xFFFF); low half of a
    // word = total of its two bins, high half = the upper bin (count16_flush<7> without a table of singles)
    for (uint32_t x = tid; x < NK; x += nthr) {
        uint32_t v = (h8[x] & 0xFFFFu) + (h8[x | 0x4000u] & 0xFFFFu);
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

}  // namespace vk
