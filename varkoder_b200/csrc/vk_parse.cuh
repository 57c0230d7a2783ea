// vk_parse.cuh -- K1: FASTQ framing in ONE pass over the text, and K1L: the ladder on the device.
//
// Stands in for the gzip line loop of split_fastq (varKoder/commands/image.py:662-667: lines are split on
// '\n', line index % 4 == 1 is a sequence line, nsites = sum(len(line) - 1)) and for the ladder that follows
// it (image.py:669-695).
//
// K1 is HBM-bound: every byte of the text is read exactly once with 16-byte loads.  Deciding what a newline
// terminates needs the number of newlines before it, a prefix sum over the whole file.  A single-pass
// decoupled look-back was tried first and measured at 1.8 TB/s (16 KiB tiles, 32-wide window) and 1.0 TB/s
// (32 KiB, 128-wide, static tiles): its throughput is tile bytes x window / L2 round trip and the tiles of a
// wave all wait at the same time (profiles/r01_notes.md).  The shipped form has no waiting at all:
//   K1a parse_mask_kernel   streams the text once, writes one 64-bit newline mask per 64 bytes (1/8 of the
//                           text) and a newline count per 32 KiB tile;
//   K1s parse_scan_kernel   one CTA: exclusive prefix of the tile counts (13 k tiles for 423 MB);
//   K1b parse_emit_kernel   reads the masks (not the text) + the tile prefix and writes, for read r,
//                           starts[r] = offset of its sequence line, ends[r] = offset of the closing newline;
// nsites falls out as sum(ends) - sum(starts) in modular arithmetic.
#pragma once
#include "vk_common.cuh"

namespace vk {

constexpr int kParseThreads = 512;
constexpr int kParseWarps = kParseThreads / 32;
constexpr int kParseWordsPerThread = 4;                                   // 4 x 16 B = 64 contiguous bytes per thread
constexpr uint32_t kParseTileBytes = kParseThreads * kParseWordsPerThread * 16;   // 32 KiB

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

// ---- K0: the first kernel of every step.  Copies the step's argument block from the pinned HOST buffer the caller has
// just written (the pointer is device-accessible: unified addressing) into device memory, where every later kernel
// reads it, and clears what the framing pass accumulates into.  A kernel rather than a copy-engine memcpy + two
// memsets: the hand-over from the copy engine to the first compute kernel cost ~45 us of a 300 us step.
constexpr uint32_t kStepArgWords = (uint32_t)(sizeof(StepArgs) / 4);
static_assert(sizeof(StepArgs) % 4 == 0, "StepArgs is copied as 32-bit words");
__global__ void __launch_bounds__(1024)
step_begin_kernel(const uint32_t* __restrict__ args_host, uint32_t* __restrict__ args_dev, Plan* __restrict__ plan,
                  uint32_t* __restrict__ tile_count, uint32_t tile_cap, int rescan, unsigned long long* __restrict__ prio_hist)
{
    pdl_wait();
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < kStepArgWords) args_dev[g] = args_host[g];
    for (uint32_t i = g; i < (uint32_t)VK_PRIO_BUCKETS; i += gridDim.x * blockDim.x) prio_hist[i] = 0;      // vk_sample.cuh
    if (!rescan) return;
    if (g < offsetof(Plan, n_lines) / 4) reinterpret_cast<uint32_t*>(plan)[g] = 0;        // parse fields: counters, sums, ticket
    for (uint32_t i = g; i < tile_cap; i += gridDim.x * blockDim.x) tile_count[i] = 0;
}

// ---- 2-bit pack (K1a with PACK): the same pass that finds the newlines also classifies every byte, blind to the
// framing: code = (byte >> 1) & 3 (A0 C1 T2 G3, the dsk code; vk_count.cuh) and valid = byte is one of ACGTacgt.
// Layout ("position layout"): text byte i has its code at bits [2 (i & 31), +2) of codes64[i >> 5] and its validity
// at bit (i & 31) of valid32[i >> 5] -- headers and quality lines are packed too (their words are never read: the
// count kernels only visit the 32-byte blocks of sequence lines, and only trust codes under a set validity bit).
// A newline is an invalid byte, so runs of k valid bytes can never span two reads.  Nothing here depends on k.
//
// Arithmetic per 32-bit word x, SIMD over its four bytes (pipe balance matters: the kernel sits at the ALU-pipe limit,
// so additions and the multiplications by small constants go to the FMA pipe as IMAD):
//   y  = x & 0x06060606                         codes at bits 1-2 of each byte
//   codes: top byte of y * 0x00820820           (no two partial products meet in bits 24..31)
//   T?  : (y * 3) & 0x08  -- y*3 is 0, 6, 12, 18 for codes A, C, T, G; only 12 has bit 3 set
//   d   = ((x & 0xD9) ^ 0x41) ^ (T? ? 0x11 : 0) -- 0 iff the byte, case bit cleared, is the letter its code names:
//         'A' 0x41, 'C' 0x43, 'G' 0x47 agree outside bits 1-2 (mask 0xD9 drops bits 1, 2 and 5 = case); 'T' 0x54 differs
//         from that pattern by 0x11.  Zero bytes of d are then found with the exact carry-free test.
__device__ __forceinline__ uint32_t mul_lo(uint32_t a, uint32_t b)      // IMAD (FMA pipe), even by a power of two
{
    uint32_t d;
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t add_fma(uint32_t a, uint32_t b, uint32_t one)      // a * 1 + b as IMAD; `one` opaque to ptxas
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
    return d;
}
struct Packed4 {
    uint32_t codes_hi;    // byte 3 = the four 2-bit codes
    uint32_t zv;          // 0x80 in every byte that is one of ACGTacgt
    uint32_t zn;          // 0x80 in every byte that is '\n'
};
__device__ __forceinline__ Packed4 pack4(uint32_t x, uint32_t one)
{
    Packed4 r;
    const uint32_t y = x & 0x06060606u;
    r.codes_hi = y * 0x00820820u;
    const uint32_t t8 = (y * 3u) & 0x08080808u;
    const uint32_t t11 = (t8 >> 3) * 0x11u;
    uint32_t d1;                                               // (x & 0xD9..) ^ 0x41..
    asm("lop3.b32 %0, %1, %2, %3, 0x6A;" : "=r"(d1) : "r"(x), "r"(0xD9D9D9D9u), "r"(0x41414141u));
    uint32_t dm;                                               // (d1 ^ t11) & 0x7F..
    asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(dm) : "r"(d1), "r"(t11), "r"(0x7F7F7F7Fu));
    const uint32_t sv = add_fma(dm, 0x7F7F7F7Fu, one);
    // bit 7 of d is bit 7 of x (neither 0x41 nor 0x11 touches it): valid <=> ~(sv | x) & 0x80
    asm("lop3.b32 %0, %1, %2, %3, 0x02;" : "=r"(r.zv) : "r"(sv), "r"(x), "r"(0x80808080u));
    uint32_t nm;                                               // (x ^ 0x0A..) & 0x7F..
    asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(nm) : "r"(x), "r"(0x0A0A0A0Au), "r"(0x7F7F7F7Fu));
    const uint32_t sn = add_fma(nm, 0x7F7F7F7Fu, one);
    asm("lop3.b32 %0, %1, %2, %3, 0x02;" : "=r"(r.zn) : "r"(sn), "r"(x), "r"(0x80808080u));
    return r;
}
// 8 flags (0x80 per byte in za, zb) -> byte 3: bits 24..27 = word a (bytes 0..3), bits 28..31 = word b
__device__ __forceinline__ uint32_t gather_flags8(uint32_t za, uint32_t zb) { return (zb | (za >> 4)) * 0x00204081u; }
// byte 3 of four registers -> one word (a lowest)
__device__ __forceinline__ uint32_t top_bytes4(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    return __byte_perm(__byte_perm(a, b, 0x0073), __byte_perm(c, d, 0x0073), 0x5410);
}

// K1a: thread t of tile T owns bytes [T*32Ki + 64t, +64): mask bit j = byte j is '\n'.
template <bool PACK>
__global__ void __launch_bounds__(kParseThreads)
parse_mask_kernel(const StepArgs* __restrict__ sa,
                  uint64_t* __restrict__ masks, uint32_t* __restrict__ tile_count, uint32_t* __restrict__ warp_count,
                  uint4* __restrict__ codes16, uint2* __restrict__ valid8)
{
    pdl_wait();
    const uint4* __restrict__ text16 = reinterpret_cast<const uint4*>(sa->text);
    const uint64_t n_bytes = sa->n_bytes;
    const uint32_t n_tiles = sa->n_tiles;
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t one = PACK ? (uint32_t)(n_bytes >> 62) + 1u : 1u;       // 1 (texts are shorter than 2^40), but not to ptxas
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t tbyte = (uint64_t)tile * kParseTileBytes + (uint64_t)tid * 64;
        uint64_t m = 0;
        if (tbyte < n_bytes) {
            uint4 w[kParseWordsPerThread];
            // streaming loads: the text must not evict the masks.  A tile that lies wholly inside the text takes four
            // unconditional loads, issued back to back (all 64 bytes of the thread in flight at once); only the last
            // tile checks every word against the end.  (With one predicated form for both, ptxas once sank the third and
            // fourth load behind the arithmetic of the second, halving the bytes in flight: 70 -> 117 us.)
            if ((uint64_t)(tile + 1) * kParseTileBytes <= n_bytes) {
                const uint4* const src = text16 + (tbyte >> 4);
#pragma unroll
                for (int i = 0; i < kParseWordsPerThread; ++i) w[i] = __ldcs(src + i);
            } else {
#pragma unroll
                for (int i = 0; i < kParseWordsPerThread; ++i) {
                    const uint64_t b = tbyte + 16ull * i;
                    w[i] = (b < n_bytes) ? __ldcs(text16 + (b >> 4)) : make_uint4(0, 0, 0, 0);
                }
            }
            if (PACK) {
                uint32_t cw[4], vb[4], nb[4];      // per 16 bytes: 32 bits of codes, validity byte pair, newline byte pair
                uint32_t vq[2][2], nq[2][2];
#pragma unroll
                for (int i = 0; i < kParseWordsPerThread; ++i) {
                    const Packed4 a = pack4(w[i].x, one), b = pack4(w[i].y, one), c = pack4(w[i].z, one), d = pack4(w[i].w, one);
                    cw[i] = top_bytes4(a.codes_hi, b.codes_hi, c.codes_hi, d.codes_hi);
                    vq[i & 1][0] = gather_flags8(a.zv, b.zv); vq[i & 1][1] = gather_flags8(c.zv, d.zv);
                    nq[i & 1][0] = gather_flags8(a.zn, b.zn); nq[i & 1][1] = gather_flags8(c.zn, d.zn);
                    if (i & 1) {
                        vb[i >> 1] = top_bytes4(vq[0][0], vq[0][1], vq[1][0], vq[1][1]);
                        nb[i >> 1] = top_bytes4(nq[0][0], nq[0][1], nq[1][0], nq[1][1]);
                    }
                }
                uint32_t v0 = vb[0], v1 = vb[1];
                m = (uint64_t)nb[0] | ((uint64_t)nb[1] << 32);
                if (n_bytes - tbyte < 64) {
                    const uint64_t keep = (1ull << (n_bytes - tbyte)) - 1;
                    m &= keep;
                    v0 &= (uint32_t)keep;
                    v1 &= (uint32_t)(keep >> 32);
                }
                codes16[tbyte >> 6] = make_uint4(cw[0], cw[1], cw[2], cw[3]);
                valid8[tbyte >> 6] = make_uint2(v0, v1);
            } else {
#pragma unroll
                for (int i = 0; i < kParseWordsPerThread; ++i) m |= (uint64_t)nl16(w[i]) << (16 * i);
                if (n_bytes - tbyte < 64) m &= (1ull << (n_bytes - tbyte)) - 1;
            }
        }
        masks[(uint64_t)tile * kParseThreads + tid] = m;
        const uint32_t c = __reduce_add_sync(0xffffffffu, (uint32_t)__popcll(m));
        if (lane == 0) {
            warp_count[(uint64_t)tile * kParseWarps + (tid >> 5)] = c;       // newlines in this warp's 2 KiB
            if (c) atomicAdd(tile_count + tile, c);
        }
    }
}

// K1s: one CTA; tile_prefix[T] = carry + newlines in tiles < T; plan->n_newlines = carry + all.
// Counts are staged through shared memory in chunks (coalesced loads/stores), each thread scans a contiguous slice.
constexpr uint32_t kScanChunk = 8192;            // 32 KiB of static shared memory
__global__ void __launch_bounds__(1024)
parse_scan_kernel(const uint32_t* __restrict__ tile_count, const StepArgs* __restrict__ sa, uint64_t* __restrict__ tile_prefix,
                  Plan* __restrict__ plan)
{
    pdl_wait();
    const uint32_t n_tiles = sa->n_tiles;
    __shared__ uint32_t s_cnt[kScanChunk];
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_total;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t carry = plan->n_newlines;
    constexpr uint32_t PER = kScanChunk / 1024;
    for (uint32_t c0 = 0; c0 < n_tiles; c0 += kScanChunk) {
        const uint32_t n = n_tiles - c0 < kScanChunk ? n_tiles - c0 : kScanChunk;
        for (uint32_t i = tid; i < kScanChunk; i += 1024) s_cnt[i] = i < n ? tile_count[c0 + i] : 0u;
        __syncthreads();
        uint32_t loc[PER], sum = 0;
#pragma unroll
        for (uint32_t i = 0; i < PER; ++i) { loc[i] = sum; sum += s_cnt[tid * PER + i]; }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t wv = s_warp[lane];
            uint32_t wi = wv;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= (uint32_t)d) wi += t;
            }
            s_warp[lane] = wi - wv;
            if (lane == 31) s_total = wi;
        }
        __syncthreads();
        const uint32_t base = s_warp[warp] + incl - sum;
#pragma unroll
        for (uint32_t i = 0; i < PER; ++i) s_cnt[tid * PER + i] = base + loc[i];      // exclusive prefix inside the chunk
        __syncthreads();
        for (uint32_t i = tid; i < n; i += 1024) tile_prefix[c0 + i] = carry + s_cnt[i];
        carry += s_total;
        __syncthreads();
    }
    if (tid == 0) plan->n_newlines = carry;
}

// K1b: one WARP per 4 KiB of text (64 masks, two per lane); no block-level synchronisation.
// A warp walks a CONTIGUOUS range of 4 KiB units, so the line index simply runs on from tile to tile (the tile prefix
// and the counts of the earlier warps of the tile are looked up once per warp, not per tile) and the masks are read
// through one pointer that advances by 256 bytes.  Everything per newline is 32-bit and relative to the warp-tile:
// line index = line_base + q, byte = tile0 + rel.  nsites = sum(ends) - sum(starts) is accumulated as one difference
// (added to plan->sum_ends; plan_kernel only ever uses sum_ends - sum_starts).
__device__ __forceinline__ uint32_t count_phase(uint32_t q0, uint32_t cnt, uint32_t ph)
{
    // #{ i in [q0, q0 + cnt) : i mod 4 == ph }
    return ((q0 + cnt + 3u - ph) >> 2) - ((q0 + 3u - ph) >> 2);
}

// One lane decodes kEmitMasks consecutive 64-byte masks (128 bytes of text); a warp-unit is 4 KiB.
constexpr int kEmitMasks = 2;      // 4 measured slower (339 vs 332 us per step): longer divergent loops
constexpr uint32_t kEmitUnitBytes = 32u * 64u * kEmitMasks;                      // 4 KiB
constexpr uint32_t kEmitUnitsPerTile = kParseTileBytes / kEmitUnitBytes;         // 8
constexpr uint32_t kEmitWarpsPerUnit = kParseWarps / kEmitUnitsPerTile;          // K1a warps (2 KiB each) per unit: 2

template <bool STORE>
__device__ __forceinline__ int32_t emit_unit(uint32_t (&w)[2 * kEmitMasks], uint32_t q, uint32_t rel0, uint64_t unit0,
                                             uint64_t* __restrict__ ps, uint64_t* __restrict__ pe)
{
    int32_t d = 0;                              // sum over sequence ends of rel - sum over sequence starts of rel
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < 2 * kEmitMasks; ++i) any |= w[i];
    while (any) {                               // one trip per newline of the busiest lane
        // lowest set bit over the words, in text order
        uint32_t cur = 0, base = 0;
        int which = -1;
#pragma unroll
        for (int i = 2 * kEmitMasks - 1; i >= 0; --i)
            if (w[i]) { cur = w[i]; base = 32u * i; which = i; }
        const uint32_t rel = rel0 + base + (uint32_t)__ffs(cur) - 1u;
        const uint32_t cleared = cur & (cur - 1);
        any = 0;
#pragma unroll
        for (int i = 0; i < 2 * kEmitMasks; ++i) {
            if (i == which) w[i] = cleared;
            any |= w[i];
        }
        const uint32_t ph = q & 3u, rr = q >> 2;
        if (ph == 0) {                          // header line ends: the sequence line starts at the next byte
            d -= (int32_t)(rel + 1);
            VK_ASSERT(rel < kEmitUnitBytes && rr < kEmitUnitBytes);
            if (STORE) ps[rr] = unit0 + rel + 1;
        } else if (ph == 1) {                   // sequence line ends
            d += (int32_t)rel;
            if (STORE) pe[rr] = unit0 + rel;
        }
        ++q;
    }
    return d;
}

__global__ void __launch_bounds__(256)
parse_emit_kernel(const uint64_t* __restrict__ masks, const uint64_t* __restrict__ tile_prefix,
                  const uint32_t* __restrict__ warp_count, const StepArgs* __restrict__ sa, uint64_t byte_base,
                  uint64_t* __restrict__ starts, uint64_t* __restrict__ ends, Plan* __restrict__ plan)
{
    pdl_wait();
    const uint32_t n_tiles = sa->n_tiles;
    const uint64_t cap_reads = sa->pa.cap_reads;
    const uint32_t lane = threadIdx.x & 31;
    // 32-bit unit arithmetic: the text is shorter than 2^40 bytes, so there are fewer than 2^28 units
    const uint32_t n_units = n_tiles * kEmitUnitsPerTile;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t per = (n_units + n_warps - 1) / n_warps;
    const uint32_t wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t g0_64 = (uint64_t)wid * per;
    const uint32_t g0 = g0_64 < n_units ? (uint32_t)g0_64 : n_units, g1 = n_units - g0 < per ? n_units : g0 + per;
    int64_t diff = 0;
    uint32_t overflow = 0;
    if (g0 < g1) {
        // line index of the first newline of unit g0: tile prefix + the K1a warps of the tile in front of the unit
        const uint32_t tile = g0 / kEmitUnitsPerTile;
        const uint32_t wit = (g0 % kEmitUnitsPerTile) * kEmitWarpsPerUnit;
        const uint32_t wc = lane < wit ? warp_count[(uint64_t)tile * kParseWarps + lane] : 0u;
        uint64_t line_base = tile_prefix[tile] + __reduce_add_sync(0xffffffffu, wc);
        // lane l owns masks kEmitMasks * l .. of the unit: 16-byte loads
        constexpr int kLd = kEmitMasks / 2;
        const ulonglong2* mp = reinterpret_cast<const ulonglong2*>(masks) + ((uint64_t)g0 * 32 + lane) * kLd;
        ulonglong2 m_nx[kLd];
#pragma unroll
        for (int i = 0; i < kLd; ++i) m_nx[i] = mp[i];
        const uint32_t rel0 = lane * (64u * kEmitMasks);
        // nsites pieces, relative to the warp's first unit: d_sum = sum of the in-unit parts, dn_sum = #ends - #starts,
        // dn_units = sum of (#ends - #starts) x (unit - g0); put together once after the loop
        int64_t d_sum = 0;
        int32_t dn_sum = 0;
        int64_t dn_units = 0;
        for (uint32_t g = g0; g < g1; ++g) {
            ulonglong2 m[kLd];
#pragma unroll
            for (int i = 0; i < kLd; ++i) m[i] = m_nx[i];
            mp += 32 * kLd;
            if (g + 1 < g1) {                                             // next unit's masks in flight while this one is decoded
#pragma unroll
                for (int i = 0; i < kLd; ++i) m_nx[i] = mp[i];
            }
            uint32_t w[2 * kEmitMasks];
#pragma unroll
            for (int i = 0; i < kLd; ++i) {
                w[4 * i] = (uint32_t)m[i].x; w[4 * i + 1] = (uint32_t)(m[i].x >> 32);
                w[4 * i + 2] = (uint32_t)m[i].y; w[4 * i + 3] = (uint32_t)(m[i].y >> 32);
            }
            uint32_t cnt = 0;
#pragma unroll
            for (int i = 0; i < 2 * kEmitMasks; ++i) cnt += __popc(w[i]);
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= (uint32_t)d) incl += t;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t lb = (uint32_t)line_base & 3u;
            const uint64_t r_base = line_base >> 2;
            const uint32_t q = lb + (incl - cnt);                          // (line index of this lane's first newline) - 4 r_base
            const uint64_t unit0 = byte_base + (uint64_t)g * kEmitUnitBytes;
            const bool fits = r_base + ((lb + total) >> 2) < cap_reads;    // warp-uniform
            int32_t d;
            if (fits) d = emit_unit<true>(w, q, rel0, unit0, starts + r_base, ends + r_base);
            else { d = emit_unit<false>(w, q, rel0, unit0, nullptr, nullptr); overflow = 1; }
            // the "+1" of every start is already in d
            const int32_t dn = (int32_t)count_phase(q, cnt, 1) - (int32_t)count_phase(q, cnt, 0);
            d_sum += d;
            dn_sum += dn;                                                  // |dn| <= 32 per unit and lane
            dn_units += (int32_t)(dn * (int32_t)(g - g0));
            line_base += total;
        }
        diff = d_sum + dn_units * (int64_t)kEmitUnitBytes
             + (int64_t)dn_sum * (int64_t)(byte_base + (uint64_t)g0 * kEmitUnitBytes);
    }
    uint64_t sum = (uint64_t)diff;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
        overflow |= __shfl_xor_sync(0xffffffffu, overflow, d);
    }
    if (lane == 0) {
        if (sum) atomicAdd((unsigned long long*)&plan->sum_ends, (unsigned long long)sum);
        if (overflow) atomicOr(&plan->table_overflow, 1u);
    }
}

// ---- K1L: finish the framing and build the ladder, single thread ---------------------------------------

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

// What the newline bookkeeping of K1 amounts to at the end of the buffer: Python's line iterator yields an unterminated
// last line too, and a header newline that is the very last byte opens no sequence line (image.py:662-667).
struct Framing {
    uint64_t n_lines, n_reads, nsites_true, nsites_ref;
    bool close_last;        // the last sequence line is closed by the end of the buffer, not by a newline
};
__device__ __forceinline__ Framing finish_framing(const uint8_t* __restrict__ text, uint64_t n, const Plan* __restrict__ plan)
{
    Framing f;
    const uint64_t T = plan->n_newlines;
    const bool last_nl = n > 0 && text[n - 1] == '\n';
    f.n_lines = T + ((n > 0 && !last_nl) ? 1 : 0);
    f.n_reads = (f.n_lines + 2) >> 2;                                  // line indices 1 mod 4
    uint64_t S = plan->sum_starts, E = plan->sum_ends;
    uint64_t ref_adjust = 0;
    f.close_last = false;
    if ((T & 3) == 1) {
        // a header newline opened a sequence line that no newline closed
        if (!last_nl) {                    // unterminated, non-empty sequence line: closes at EOF
            f.close_last = true;
            E += n;
            ref_adjust = 1;                // len(line) - 1 drops a real base here (image.py:666)
        } else {
            S -= n;                        // header newline was the last byte: no such line for Python
        }
    }
    f.nsites_true = E - S;
    f.nsites_ref = E - S - ref_adjust;
    return f;
}

// read-sharded samples: this shard's (records, bases) for the all-gather that precedes plan_kernel
__global__ void __launch_bounds__(32)
shard_stats_kernel(const StepArgs* __restrict__ sa, const Plan* __restrict__ plan, unsigned long long* __restrict__ out2)
{
    pdl_wait();
    if (threadIdx.x != 0) return;
    const Framing f = finish_framing(sa->text, sa->pa.n_bytes, plan);
    out2[0] = f.n_reads;
    out2[1] = f.nsites_ref;
}

__global__ void __launch_bounds__(64)
plan_kernel(const StepArgs* __restrict__ sa, uint64_t* __restrict__ starts, uint64_t* __restrict__ ends,
            Plan* __restrict__ plan)
{
    pdl_wait();
    const uint8_t* __restrict__ text = sa->text;
    const PlanArgs a = sa->pa;
    // one CTA of 64 threads: thread 0 finishes the framing and walks the ladder (sequential by nature, ~10 levels),
    // then thread l computes level l's priority threshold (a 64-step long division each) and clears its segment slots
    __shared__ uint64_t s_lv[kMaxLevels];
    __shared__ uint64_t s_nsites;
    __shared__ int s_nl;
    const int l = threadIdx.x;
    if (l == 0) {
        const uint64_t n = a.n_bytes;
        const Framing f = finish_framing(text, n, plan);
        const uint64_t n_reads = f.n_reads;
        if (f.close_last && n_reads >= 1 && n_reads - 1 < a.cap_reads) ends[n_reads - 1] = n;
        plan->n_bytes = n;
        plan->n_lines = f.n_lines;
        plan->n_reads = n_reads;
        plan->nsites_true = f.nsites_true;
        plan->nsites_ref = f.nsites_ref;
        if (n_reads > a.cap_reads) plan->table_overflow = 1;
        // sample-wide base count and global index of this buffer's first record
        uint64_t nsites_all = a.p.nsites_override ? a.p.nsites_override : f.nsites_ref;
        uint64_t index_base = a.p.read_index_base, reads_all = n_reads;
        if (a.shard_table) {
            nsites_all = 0; index_base = 0; reads_all = 0;
            for (uint32_t r = 0; r < a.shard_world; ++r) {
                if (r < a.shard_rank) index_base += a.shard_table[2 * r];
                reads_all += a.shard_table[2 * r];
                nsites_all += a.shard_table[2 * r + 1];
            }
        }
        plan->read_index_base = index_base;
        plan->total_reads = reads_all;
        plan->count_overflow = 0;
        plan->lanes_verdict = 0;

        // ---- ladder, image.py:669-695 in integers
        const uint64_t nsites = nsites_all;
        plan->nsites_ladder = nsites;
        int nl = 0;
        int status = VK_LADDER_OK;
        if (!a.p.has_max_bp) s_lv[nl++] = nsites;
        else if (a.p.is_query || nsites > a.p.min_bp) s_lv[nl++] = nsites < a.p.max_bp ? nsites : a.p.max_bp;
        else status = VK_LADDER_LESS_THAN_MIN;
        if (status == VK_LADDER_OK && !a.p.is_query) {
            while (s_lv[nl - 1] > a.p.min_bp && nl < kMaxLevels) {
                const uint64_t oneless = s_lv[nl - 1] - 1;
                if (oneless == 0) break;                      // the reference would raise in log10(0); min_bp = 0 only
                uint64_t p10 = 1;                             // 10^floor(log10(oneless)), without a division per digit
                while (p10 <= 1000000000000000000ull && p10 * 10 <= oneless) p10 *= 10;
                const uint64_t fd = oneless / p10;
                const uint64_t mult = fd >= 5 ? 5 : (fd >= 2 ? 2 : 1);   // largest of {1,2,5} <= first digit
                s_lv[nl++] = mult * p10;
            }
            if (s_lv[nl - 1] < a.p.min_bp) --nl;
        }
        if (status != VK_LADDER_OK) nl = 0;
        plan->status = status;
        plan->n_levels = nl;
        plan->long_reads = 0;
        plan->len_min = 0xFFFFFFFFu;
        plan->n_long = 0;
        plan->hist_ticket = 0;
        plan->len_max = 0;
        s_nl = nl;
        s_nsites = nsites;
    }
    __syncthreads();
    __shared__ double s_frac[kMaxLevels];            // share of the reads expected in segment l
    if (l < kMaxLevels) {
        const int nl = s_nl;
        const uint64_t nsites = s_nsites;
        const uint64_t bp = l < nl ? s_lv[l] : 0;
        const bool all = l < nl && (bp >= nsites || nsites == 0);
        const uint64_t thr = l < nl ? (all ? kThrAll : div_2p64(bp, nsites)) : 0;
        plan->level_bp[l] = bp;
        plan->level_all[l] = all ? 1u : 0u;
        plan->level_thr[l] = thr;
        plan->seg_reads[l] = 0;
        plan->seg_bases[l] = 0;
        plan->seg_extra[l] = 0;
        plan->seg_next[l] = 0;
        plan->seg_next2[l] = 0;
        plan->seg_chunks[l] = 0;
        s_frac[l] = all ? 1.0 : (double)thr * 5.421010862427522e-20;      // thr / 2^64
    }
    __syncthreads();
    // ---- layout of the segment-sorted read table and of the count kernel's CTAs, from EXPECTED segment sizes:
    // segment s holds the reads with thr[s+1] <= prio < thr[s], a binomial share of the n_reads reads.  A region
    // gets its expectation + 8 sigma + slack (an overflow is detected by the scatter kernel and retried with
    // exact_layout), so no counting pass over the read table is needed.  Thread s owns segment s.
    __shared__ double s_w[kMaxLevels];
    __shared__ uint64_t s_cap[kMaxLevels];
    __shared__ uint32_t s_ncta[kMaxLevels];
    __shared__ double s_wsum;
    {
        const int nl = s_nl;
        const uint64_t n_reads = plan->n_reads;
        const uint64_t stride = (n_reads + kUnitReads - 1) / kUnitReads * kUnitReads + kUnitReads;
        double f = 0.0;
        if (l < nl) f = s_frac[l] - (l + 1 < nl ? s_frac[l + 1] : 0.0);
        if (f < 0.0) f = 0.0;
        uint64_t cap = 0;
        if (l < nl) {
            const double e = (double)n_reads * f;
            cap = a.test_tight ? (uint64_t)(0.5 * e) + 1 : (uint64_t)(e + 8.0 * sqrt(e) + 1024.0);
            cap = (cap + kUnitReads - 1) / kUnitReads * kUnitReads;
            if (cap > stride || a.exact_layout) cap = stride;
        }
        s_w[l] = f;
        s_cap[l] = cap;
        plan->seg_cap[l] = cap;
        __syncthreads();
        __shared__ uint64_t s_off;
        if (l == 0) {
            double wsum = 0.0;
            uint64_t off = 0;
            for (int s = 0; s < nl; ++s) {                 // only the ladder's segments have a share / a region
                wsum += s_w[s];
                plan->seg_begin[s] = off;
                off += s_cap[s];
            }
            s_off = off;
            plan->bucket_overflow = (a.cap_chunks == 0 && off > a.cap_sorted) ? 1u : 0u;
            s_wsum = wsum;
            // chunk table (k <= 7): a read of len bases whose first base sits rlo bytes into its 16-byte word has
            // (rlo + len + 31) / 32 chunks, at most len / 32 + 2; regions from the same expected shares
            uint64_t coff = 0;
            if (a.cap_chunks) {
                const double B = (double)plan->nsites_true / 32.0 + 2.0 * (double)n_reads;      // upper bound on all chunks
                const double per_read = n_reads ? B / (double)n_reads : 0.0;
                for (int s2 = 0; s2 < nl; ++s2) {
                    const double e = (double)n_reads * s_w[s2];
                    double capd = a.test_tight ? 0.5 * e * per_read + 1.0 : (e + 8.0 * sqrt(e) + 1024.0) * per_read + 4096.0;
                    if (a.exact_layout || capd > B + 4096.0) capd = B + 4096.0;
                    const uint64_t cap = ((uint64_t)capd + 255u) & ~255ull;
                    plan->seg_cbegin[s2] = coff;
                    plan->seg_ccap[s2] = cap;
                    coff += cap;
                }
                plan->chunks_needed = coff;
                plan->chunk_table_small = coff > a.cap_chunks ? 1u : 0u;
            } else {
                plan->chunks_needed = 0;
                plan->chunk_table_small = 0;
            }
            for (int s2 = a.cap_chunks ? nl : 0; s2 < kMaxLevels; ++s2) { plan->seg_cbegin[s2] = coff; plan->seg_ccap[s2] = 0; }
            plan->seg_cbegin[kMaxLevels] = coff;
        }
        __syncthreads();
        if (l >= nl) plan->seg_begin[l] = s_off;           // empty regions at the end
        if (l == 0) plan->seg_begin[kMaxLevels] = s_off;
        // CTAs: one per segment, the rest in proportion to the expected bases; leftovers one by one to the segment with
        // the most expected bases per CTA (greedy = optimal for the slowest segment, which is what the kernel waits for)
        uint32_t n_ctas = a.n_count_ctas;
        if (a.reads_per_cta) {
            const uint64_t want = (n_reads + a.reads_per_cta - 1) / a.reads_per_cta;
            if (want < n_ctas) n_ctas = want > (uint64_t)nl ? (uint32_t)want : (uint32_t)nl;
            if (n_ctas > a.n_count_ctas) n_ctas = a.n_count_ctas;
        }
        const bool enough = (uint32_t)nl <= n_ctas;
        const uint32_t spare = enough ? n_ctas - (uint32_t)nl : 0u;
        uint32_t mine = (l < nl && enough) ? 1u : 0u;
        if (l < nl && enough && s_wsum > 0.0) mine += (uint32_t)((double)spare * (f / s_wsum));
        s_ncta[l] = mine;
        __syncthreads();
        if (l == 0 && enough && s_wsum > 0.0) {
            uint32_t given = 0;
            for (int s = 0; s < nl; ++s) given += s_ncta[s];
            for (uint32_t left = n_ctas > given ? n_ctas - given : 0u; left > 0; --left) {
                int best = 0;
                float load = -1.f;
                for (int s = 0; s < nl; ++s) {
                    const float ld = __fdividef((float)s_w[s], (float)s_ncta[s]);
                    if (ld > load) { load = ld; best = s; }
                }
                ++s_ncta[best];
            }
        }
        __syncthreads();
        __shared__ uint32_t s_crun;
        if (l == 0) {
            uint32_t crun = 0;
            for (int s = 0; s < nl; ++s) { plan->seg_cta_begin[s] = crun; crun += s_ncta[s]; }
            s_crun = crun;
            plan->seg_cta_begin[kMaxLevels] = crun;
        }
        __syncthreads();
        if (l >= nl) plan->seg_cta_begin[l] = s_crun;      // no CTAs beyond the ladder
        // A read table or a segment layout that does not fit its allocation: the host repeats the step with larger
        // buffers (with_table_retry).  Until then nothing downstream may touch the tables: starts / ends beyond the
        // capacity were never written, regions would lie outside `sorted`.  No CTA gets a segment, no region has room;
        // bucket_scatter_kernel and base_content_kernel return at once when they see either flag.
        __syncthreads();
        if (plan->table_overflow || plan->bucket_overflow || plan->chunk_table_small) {
            plan->seg_cta_begin[l] = 0;
            plan->seg_cap[l] = 0;
            plan->seg_begin[l] = 0;
            plan->seg_ccap[l] = 0;
            plan->seg_cbegin[l] = 0;
            if (l == 0) { plan->seg_cta_begin[kMaxLevels] = 0; plan->seg_begin[kMaxLevels] = 0; }
        }
    }
}

}  // namespace vk
