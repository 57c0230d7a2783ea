// vk_countt9.cuh -- K2t9: k = 9 with countt_kernel's front end (one read per lane, text staged by cp.async in half units,
// the owner classifies its words from shared memory) and count9h_kernel's back end (CTA pairs, the 2 x 4^8 canonical
// classes as two tables of 16-bit bins, exact returning adds + drains of hot words: vk_count.cuh "k = 9").
//
// Stands in for `dsk -kmer-size 9 -abundance-min 1` on every sub-sample file (varKoder/commands/image.py:771-796) exactly
// as count9h_kernel does; the slabs it writes are count9h_kernel's (reduce_slabs9h_kernel follows either).
//
// Word v of a read = the sixteen 9-mers that END at bases 16v + 8 .. 16v + 23 (window = bases 16v .. 16v + 23: eight bases
// of context + sixteen, the 48-bit window count9_steps wants).  No queue: the "a 9-mer ends here" mask E of a word is exact
// in every lane (validity, read length, cut points) and an increment of a 9-mer that is not to be counted is 0, as in
// count9h_kernel.  Both CTAs of a pair walk the same contiguous share of the segment's reads, each counting its half of
// the classes; units are claimed from a counter in the CTA's shared memory.
//
// 2.19 ms per Gbp against 2.44 ms for count9h_kernel (profiles/r02c_notes.md); the default for samples of one read length
// (the context goes by its last sample, as for k = 7: vk_capi.cu launch_count).
#pragma once
#include "vk_countt.cuh"

namespace vk {

constexpr uint32_t countt9_smem_bytes() { return (32768u + 16u * kTBufWords) * 4u + 2048u; }      // table | 16 staging buffers, + alignment of the table

__global__ void __launch_bounds__(512, 1)
countt9_kernel(const StepArgs* __restrict__ sa, const uint64_t* __restrict__ sorted, Plan* __restrict__ plan,
               uint32_t* __restrict__ slabs, uint32_t policy)
{
    pdl_wait();
    constexpr int K = 9;
    constexpr uint32_t NB = 65536u;               // bins of one half
    constexpr uint32_t FULL = 0xffffffffu;
    const uint8_t* __restrict__ text = sa->text;
    const int breaklen = sa->pa.p.breaklength;
    // policy 2: by the sample (launched alone: a sample that is not one for this kernel is refused and the host repeats the
    // step with count9h_kernel, vk_capi.cu); 1: every sample
    if (!countu_wanted(plan, breaklen, policy)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&plan->lanes_verdict, 2u);
        return;
    }
    const uint32_t zero = (uint32_t)(sa->n_bytes >> 62);              // 0 (texts are shorter than 2^40), but not to ptxas
    const uint32_t one = zero + 1u;
    const uint64_t total_quads = (sa->n_bytes + 15u) >> 4;
    extern __shared__ __align__(16) uint32_t s_rawt9[];
    __shared__ uint32_t s_next;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
    const uint32_t pair = logical_cta() >> 1, half = logical_cta() & 1u;
    const uint32_t raw_addr = (uint32_t)__cvta_generic_to_shared(s_rawt9);
    const uint32_t h_addr = (raw_addr + 2047u) & ~2047u;              // count9_steps: the table at a 2 KiB-aligned shared address
    uint32_t* const h8 = s_rawt9 + ((h_addr - raw_addr) >> 2);
    const uint32_t buf_addr = h_addr + (32768u + warp * kTBufWords + 4u) * 4u;      // quad 0 of the warp's buffer

    uint32_t* const slab = slabs + (size_t)logical_cta() * NB;
    for (uint32_t i = tid; i < NB / 4u; i += nthr) reinterpret_cast<uint4*>(slab)[i] = make_uint4(0, 0, 0, 0);      // also for CTAs without a segment (reduce reads all)
    const int seg = cta_segment(plan, lane, pair);
    if (seg < 0) return;
    const uint32_t seg_len = (uint32_t)(plan->seg_reads[seg] < plan->seg_cap[seg] ? plan->seg_reads[seg] : plan->seg_cap[seg]);
    const uint64_t* __restrict__ seg_sorted = sorted + plan->seg_begin[seg];
    const uint32_t S = ((plan->len_max + 30u) >> 4) | 1u;
    const uint32_t R = kTQuads / S >= 32u ? 32u : kTQuads / S;
    uint32_t H = (S + 5u) >> 1;                   // (one word per iteration: the second half is needed from word H - 3 on)
    if (H > S - 1u) H = S - 1u;
    if (R && H > 256u / R) H = 256u / R;
    // the pairs of a segment take equal contiguous shares of its sorted reads; both CTAs of a pair walk the same share
    const uint32_t n_pairs_seg = plan->seg_cta_begin[seg + 1] - plan->seg_cta_begin[seg];
    const uint32_t per_pair = (seg_len + n_pairs_seg - 1u) / n_pairs_seg;
    const uint64_t lo64 = (uint64_t)(pair - plan->seg_cta_begin[seg]) * per_pair;
    const uint32_t cta_lo = lo64 < seg_len ? (uint32_t)lo64 : seg_len;
    const uint32_t cta_hi = seg_len - cta_lo < per_pair ? seg_len : cta_lo + per_pair;
    if (tid == 0) s_next = cta_lo;
    for (uint32_t i = tid; i < 8192u; i += nthr) reinterpret_cast<uint4*>(h8)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (R == 0u) {                                // a read too long for a staging buffer: the exact flat-lane kernel counts the step
        if (tid == 0) atomicOr(&plan->count_overflow, 1u);
        return;
    }

    auto entry_at = [&](uint32_t base) -> uint64_t {
        return (lane < R && base < cta_hi && base + lane < cta_hi) ? __ldg(seg_sorted + base + lane) : 0ull;
    };
    auto clamp_hi = [&](uint32_t b) -> uint32_t { return b < cta_hi ? b : cta_hi; };
    uint32_t claim0 = 0;
    if (lane == 0) claim0 = atomicAdd(&s_next, 3u * R);
    uint32_t baseA = clamp_hi(__shfl_sync(FULL, claim0, 0));
    uint32_t baseB = clamp_hi(baseA + R), baseC = clamp_hi(baseA + 2u * R);
    uint64_t entA = entry_at(baseA), entB = entry_at(baseB), entC = entry_at(baseC);
    uint32_t pending = 0;
    if (lane == 0) pending = atomicAdd(&s_next, R);

    // copy plan of the two halves of a unit (vk_countt.cuh)
    uint32_t rwF[kTRoundsF], rwS[kTRoundsS];
    {
        const uint32_t S2 = S - H;
        const uint32_t invH = ((1u << 20) + H - 1u) / H, invS = S2 ? ((1u << 20) + S2 - 1u) / S2 : 0u;
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
        const uint8_t* const p0 = len ? text + ((ent >> kEntryLenBits) & ~15ull) : text_end;
        const uint32_t plo = (uint32_t)(uintptr_t)p0, phi = (uint32_t)((uintptr_t)p0 >> 32);
#pragma unroll
        for (uint32_t n = 0; n < NR; ++n) {
            const uint32_t wb = (rw[n] >> 5) & 0x1FF0u, db = (rw[n] >> 18) & 0x1FF0u;
            const uint32_t blo = __shfl_sync(FULL, plo, (int)rw[n]), bhi = __shfl_sync(FULL, phi, (int)rw[n]);
            const uint8_t* const src = reinterpret_cast<const uint8_t*>(((uint64_t)bhi << 32) | blo) + wb;
            const bool ok = src < text_end && (int32_t)rw[n] >= 0;
            VK_ASSERT(db <= kTQuads * 16u);                              // quads 0 .. kTQuads - 1 of the warp's buffer, or the slack quad behind it
            VK_ASSERT(!ok || (src >= text && src + 16 <= text_end));      // the source word lies inside the text allocation
            cp_async16(buf_addr + db, ok ? src : text, ok ? 16u : 0u);
        }
        cp_async_commit();
    };
    auto classify16 = [&](const uint4 q, uint32_t& P, uint32_t& V) {
        const Cls4z c0 = classify4z(q.x, one), c1 = classify4z(q.y, one), c2 = classify4z(q.z, one), c3 = classify4z(q.w, one);
        P = __byte_perm(__byte_perm(c0.packed_hi, c1.packed_hi, 0x0073), __byte_perm(c2.packed_hi, c3.packed_hi, 0x0073), 0x5410);
        V = valid16(c0, c1, c2, c3);
    };
    // (the multipliers of count9_steps as plain constants: ptxas makes LEA.HI / shifts of them; hiding them as count9h_kernel
    // does -- IMAD.HI on the FMA pipe -- measured the same here, 2.202 against 2.192 ms)
    const Opaque9 q9 = {0x80000000u, 2048u, h_addr, (1u - half) << 19};

    issue_half(entA, rwF);
    while (baseA < cta_hi) {
        const uint32_t baseD = clamp_hi(__shfl_sync(FULL, pending, 0));
        const uint64_t entD = entry_at(baseD);
        if (lane == 0 && baseD < cta_hi) pending = atomicAdd(&s_next, R);
        issue_half(entA, rwS);
        cp_async_wait<1>();                                             // first half of this unit
        __syncwarp();
        const uint32_t len = (uint32_t)(entA & kEntryLenMask);
        const uint32_t o = (uint32_t)(entA >> kEntryLenBits) & 15u;
        const bool active = len != 0u;
        const bool brk = breaklen > 0 && len > (uint32_t)breaklen;
        const uint32_t lmax = __reduce_max_sync(FULL, len);
        const uint32_t nW = lmax >= (uint32_t)K ? (lmax - 8u + 15u) >> 4 : 0u;      // words of the longest read
        bool waited = false, filled = false;
        if (H < 4u) { cp_async_wait<0>(); __syncwarp(); waited = true; }
        if (nW != 0u) {
            const uint32_t s = (o + 8u) & 15u, tb = (o + 8u) >> 4;
            uint32_t qa = buf_addr + ((lane < R ? lane : 0u) * S + tb) * 16u;
            uint32_t Pp, Vp, P, V;
            classify16(lds128(qa - 16u), Pp, Vp);                       // T = tb - 1
            classify16(lds128(qa), P, V);                               // T = tb
            uint32_t C = __funnelshift_r(Pp, P, 2u * s);
            uint32_t Vv = ((Vp | (V << 16)) >> s) & 0xFFFFu;
            uint32_t Cc = C >> 16, Vc = Vv >> 8;                        // bases 0..7 of the read: context of word 0
            Pp = P;
            Vp = V;
            uint4 nq = lds128(qa + 16u);
            for (uint32_t v = 0; v < nW; ++v) {
                if (!waited && v + 3u >= H) { cp_async_wait<0>(); __syncwarp(); waited = true; }      // second half of this unit
                classify16(nq, P, V);
                qa += 16u;
                nq = lds128(qa + 16u);
                C = __funnelshift_r(Pp, P, 2u * s);
                Vv = ((Vp | (V << 16)) >> s) & 0xFFFFu;
                // window: bases 16 v .. 16 v + 23; bit b of VW <-> base 16 v + b is a letter of the read
                const uint32_t b0 = 16u * v;
                const uint32_t left = (active && b0 < len) ? len - b0 : 0u;
                const uint32_t VW = (Vc | (Vv << 8)) & (left < 24u ? (1u << left) - 1u : 0xFFFFFFu);
                uint32_t E = (uint32_t)runs_of_k64<K>((uint64_t)VW) & 0xFFFFu;      // bit j: the 9-mer that ends at base 16 v + 8 + j is countable
                if (brk) {
                    // reformat.sh breaklength: no 9-mer may span a multiple of breaklen counted from the read's first base
                    const int32_t qb = (int32_t)b0 + 8;                 // base of E's bit 0
                    int32_t c = (qb / breaklen) * breaklen;
                    if (c < breaklen) c = breaklen;
                    uint32_t dead = 0;
                    for (; c - qb < 16 && c < (int32_t)len; c += breaklen) {
                        const int32_t b = c - qb;                       // base that starts the new piece: ends b .. b+7 are dead
                        if (b > -(K - 1)) dead |= b >= 0 ? 0xFFu << b : 0xFFu >> (-b);
                    }
                    E &= ~dead;
                }
                const uint64_t W = (uint64_t)Cc | ((uint64_t)C << 16);
                const uint64_t W4 = W << 2;
                const uint64_t R4 = (revcomp_groups64(W) >> 14) << 2;  // rc of the 9-mer ending at j: R4 >> (32 - 2j)
                const uint32_t Wl = (uint32_t)W4, Wh = (uint32_t)(W4 >> 32), Rl = (uint32_t)R4, Rh = (uint32_t)(R4 >> 32);
                uint32_t acc = 0;
                count9_steps<0>(Wl, Wh, Rl, Rh, E, q9, acc);
                if (__ballot_sync(FULL, (acc & 0xC000u) != 0) != 0) {  // rare: some word is running hot
                    if (acc & 0xC000u) drain9_steps<0>(Wl, Wh, Rl, Rh, h8, slab);
                }
                Cc = C >> 16;
                Vc = Vv >> 8;
                Pp = P;
                Vp = V;
                if (!filled && v + 2u >= H) { __syncwarp(); issue_half(entB, rwF); filled = true; }      // first half of the next unit
            }
        }
        if (!waited) cp_async_wait<0>();
        __syncwarp();
        if (!filled) issue_half(entB, rwF);
        entA = entB; entB = entC; entC = entD;
        baseA = baseB; baseB = baseC; baseC = baseD;
    }
    cp_async_wait<0>();
    __threadfence();                                                    // the drains' atomics have reached the slab
    __syncthreads();
    count16_flush<8>(h8, nullptr, slab, tid, nthr);
}

}  // namespace vk
