// vk_parse.cuh -- K1: FASTQ framing in ONE pass over the text, and K1L: the ladder on the device.
//
// Stands in for the gzip line loop of split_fastq (varKoder/commands/image.py:662-667: lines are split on
// '\n', line index % 4 == 1 is a sequence line, nsites = sum(len(line) - 1)) and for the ladder that follows
// it (image.py:669-695).
//
// K1 is HBM-bound: every byte is read exactly once with 16-byte loads.  A tile of 16 KiB is owned by one
// CTA; the line index of a tile's first byte (which decides what each newline in the tile terminates) is
// the exclusive prefix sum of newline counts over all earlier tiles, obtained with a single-pass
// decoupled look-back (one 64-bit status word per tile: 2 flag bits + 62 value bits), so there is no
// second pass over the text.  Output: for read r, starts[r] = offset of its sequence line and ends[r] =
// offset of the newline closing it; nsites falls out as sum(ends) - sum(starts) in modular arithmetic.
#pragma once
#include "vk_common.cuh"

namespace vk {

constexpr int kParseThreads = 256;
constexpr int kParseWordsPerThread = 4;                                   // 4 x 16 B = 64 contiguous bytes per thread
constexpr uint32_t kParseTileBytes = kParseThreads * kParseWordsPerThread * 16;   // 16 KiB
constexpr uint64_t kFlagAgg = 1ull << 62;
constexpr uint64_t kFlagIncl = 2ull << 62;
constexpr uint64_t kValMask = (1ull << 62) - 1;

// 4-bit mask of the bytes of x equal to '\n'
__device__ __forceinline__ uint32_t nl4(uint32_t x)
{
    uint32_t d = x ^ 0x0A0A0A0Au;
    uint32_t t = (d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    uint32_t z = ~(t | d | 0x7F7F7F7Fu);          // 0x80 in every byte of x that was '\n'
    return (z * 0x00204081u) >> 28;               // gather the four flag bits
}
__device__ __forceinline__ uint32_t nl16(uint4 w)
{
    return nl4(w.x) | (nl4(w.y) << 4) | (nl4(w.z) << 8) | (nl4(w.w) << 12);
}

__global__ void __launch_bounds__(kParseThreads)
parse_kernel(const uint4* __restrict__ text16, uint64_t n_bytes, uint64_t byte_base, uint32_t n_tiles,
             uint64_t* __restrict__ tile_status, uint64_t* __restrict__ starts, uint64_t* __restrict__ ends,
             uint64_t cap_reads, Plan* __restrict__ plan)
{
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_prefix;
    __shared__ uint32_t s_warp[kParseThreads / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t carry_in = plan->n_newlines;       // read once, before any tile of this launch can finish
    uint64_t sum_s = 0, sum_e = 0;
    uint32_t overflow = 0;

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(&plan->parse_ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= n_tiles) break;

        // ---- load 64 contiguous bytes, build the newline mask
        const uint64_t tbyte = (uint64_t)tile * kParseTileBytes + (uint64_t)tid * (kParseWordsPerThread * 16);
        uint64_t m = 0;
        if (tbyte < n_bytes) {
            uint4 w[kParseWordsPerThread];
#pragma unroll
            for (int i = 0; i < kParseWordsPerThread; ++i) {
                const uint64_t b = tbyte + 16ull * i;
                w[i] = (b < n_bytes) ? __ldg(text16 + (b >> 4)) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int i = 0; i < kParseWordsPerThread; ++i) m |= (uint64_t)nl16(w[i]) << (16 * i);
            const uint64_t left = n_bytes - tbyte;             // bytes of this thread's span inside the buffer
            if (left < 64) m &= (1ull << left) - 1;
        }
        const uint32_t cnt = __popcll(m);

        // ---- block exclusive scan of the counts
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();

        if (warp == 0) {
            uint32_t wv = (lane < kParseThreads / 32) ? s_warp[lane] : 0;
            uint32_t wi = wv;
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += t;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, wi, kParseThreads / 32 - 1);
            if (lane < kParseThreads / 32) s_warp[lane] = wi - wv;      // exclusive warp offsets
            // ---- decoupled look-back
            uint64_t excl = carry_in;
            if (tile == 0) {
                if (lane == 0) st_volatile_u64(tile_status + 0, kFlagIncl | (carry_in + total));
            } else {
                if (lane == 0) st_volatile_u64(tile_status + tile, kFlagAgg | (uint64_t)total);
                int64_t j = (int64_t)tile - 1;
                excl = 0;
                for (;;) {
                    const int64_t idx = j - (int64_t)lane;
                    uint64_t v = (idx >= 0) ? ld_volatile_u64(tile_status + idx) : (kFlagIncl | carry_in);
                    // before tile 0 sits a virtual inclusive prefix; only the first lane past it counts
                    if (idx < -1) v = kFlagIncl;
                    if (__any_sync(0xffffffffu, (v >> 62) == 0)) continue;     // a predecessor has not published yet
                    const uint32_t inc = __ballot_sync(0xffffffffu, (v >> 62) == 2);
                    uint64_t c = v & kValMask;
                    if (inc) {
                        const int first = __ffs(inc) - 1;                       // nearest predecessor with a full prefix
                        if ((int)lane > first) c = 0;
                    }
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
                    excl += c;
                    if (inc) break;
                    j -= 32;
                }
                if (lane == 0) st_volatile_u64(tile_status + tile, kFlagIncl | (excl + total));
            }
            if (lane == 0) {
                s_prefix = excl;
                if (tile == n_tiles - 1) plan->n_newlines = excl + total;       // carry for the next chunk / final total
            }
        }
        __syncthreads();

        // ---- what does each newline terminate?  line index = number of newlines before it
        uint64_t line = s_prefix + s_warp[warp] + (incl - cnt);
        const uint64_t pos0 = byte_base + tbyte;
        while (m) {
            const int j = __ffsll((long long)m) - 1;
            m &= m - 1;
            const uint64_t pos = pos0 + j;
            const uint32_t ph = (uint32_t)line & 3u;
            const uint64_t r = line >> 2;
            if (ph == 0) {                      // header line ends: the sequence line starts at pos + 1
                sum_s += pos + 1;
                if (r < cap_reads) starts[r] = pos + 1; else overflow = 1;
            } else if (ph == 1) {               // sequence line ends
                sum_e += pos;
                if (r < cap_reads) ends[r] = pos; else overflow = 1;
            }
            ++line;
        }
        // s_tile / s_prefix / s_warp are rewritten only after the next iteration's barriers
    }

#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sum_s += __shfl_xor_sync(0xffffffffu, sum_s, d);
        sum_e += __shfl_xor_sync(0xffffffffu, sum_e, d);
        overflow |= __shfl_xor_sync(0xffffffffu, overflow, d);
    }
    if (lane == 0) {
        if (sum_s) atomicAdd((unsigned long long*)&plan->sum_starts, (unsigned long long)sum_s);
        if (sum_e) atomicAdd((unsigned long long*)&plan->sum_ends, (unsigned long long)sum_e);
        if (overflow) atomicOr(&plan->table_overflow, 1u);
    }
}

// ---- K1L: finish the framing and build the ladder, single thread ---------------------------------------
struct PlanArgs {
    vk_params p;
    uint64_t n_bytes;       // total bytes of the buffer
    uint64_t cap_reads;
};

__device__ inline uint64_t div_2p64(uint64_t num, uint64_t den)
{
    // floor(num * 2^64 / den) for num < den: restoring long division, 64 quotient bits
    uint64_t q = 0, rem = num;
    for (int i = 0; i < 64; ++i) {
        const bool carry = rem >> 63;
        rem <<= 1;
        q <<= 1;
        if (carry || rem >= den) { rem -= den; q |= 1; }
    }
    return q;
}

__global__ void plan_kernel(const uint8_t* __restrict__ text, uint64_t* __restrict__ starts,
                            uint64_t* __restrict__ ends, PlanArgs a, Plan* __restrict__ plan)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint64_t n = a.n_bytes;
    const uint64_t T = plan->n_newlines;
    const bool last_nl = n > 0 && text[n - 1] == '\n';
    const uint64_t n_lines = T + ((n > 0 && !last_nl) ? 1 : 0);      // Python yields an unterminated last line too
    const uint64_t n_reads = (n_lines + 2) >> 2;                      // line indices 1 mod 4
    uint64_t S = plan->sum_starts, E = plan->sum_ends;
    uint64_t ref_adjust = 0;
    if ((T & 3) == 1) {
        // a header newline opened a sequence line that no newline closed
        if (!last_nl) {                    // unterminated, non-empty sequence line: closes at EOF
            if (n_reads >= 1 && n_reads - 1 < a.cap_reads) ends[n_reads - 1] = n;
            E += n;
            ref_adjust = 1;                // len(line) - 1 drops a real base here (image.py:666)
        } else {
            S -= n;                        // header newline was the last byte: no such line for Python
        }
    }
    plan->n_bytes = n;
    plan->n_lines = n_lines;
    plan->n_reads = n_reads;
    plan->nsites_true = E - S;
    plan->nsites_ref = E - S - ref_adjust;
    if (n_reads > a.cap_reads) plan->table_overflow = 1;

    // ---- ladder, image.py:669-695 in integers
    const uint64_t nsites = a.p.nsites_override ? a.p.nsites_override : plan->nsites_ref;
    plan->nsites_ladder = nsites;
    int nl = 0;
    int status = VK_LADDER_OK;
    uint64_t lv[kMaxLevels];
    if (!a.p.has_max_bp) lv[nl++] = nsites;
    else if (a.p.is_query || nsites > a.p.min_bp) lv[nl++] = nsites < a.p.max_bp ? nsites : a.p.max_bp;
    else status = VK_LADDER_LESS_THAN_MIN;
    if (status == VK_LADDER_OK && !a.p.is_query) {
        while (lv[nl - 1] > a.p.min_bp && nl < kMaxLevels) {
            const uint64_t oneless = lv[nl - 1] - 1;
            if (oneless == 0) break;                      // the reference would raise in log10(0); min_bp = 0 only
            uint64_t p10 = 1;
            while (oneless / p10 >= 10) p10 *= 10;        // 10^floor(log10(oneless))
            const uint64_t fd = oneless / p10;
            const uint64_t mult = fd >= 5 ? 5 : (fd >= 2 ? 2 : 1);   // largest of {1,2,5} <= first digit
            lv[nl++] = mult * p10;
        }
        if (lv[nl - 1] < a.p.min_bp) --nl;
    }
    if (status != VK_LADDER_OK) nl = 0;
    plan->status = status;
    plan->n_levels = nl;
    for (int l = 0; l < kMaxLevels; ++l) {
        plan->level_bp[l] = l < nl ? lv[l] : 0;
        const bool all = l < nl && (lv[l] >= nsites || nsites == 0);
        plan->level_all[l] = all ? 1u : 0u;
        plan->level_thr[l] = l < nl ? (all ? kThrAll : div_2p64(lv[l], nsites)) : 0;
        plan->seg_reads[l] = 0;
        plan->seg_bases[l] = 0;
        plan->seg_cursor[l] = 0;
        plan->seg_next[l] = 0;
    }
    plan->long_reads = 0;
}

}  // namespace vk
