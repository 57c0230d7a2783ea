// vk_image.cuh -- K4: canonical fold, pixel mapping and rank scaling to uint8, all in integers.
//
// Stands in for dsk2ascii + the pandas/numpy half of make_image (varKoder/commands/image.py:875-919):
//   :900      left-join pixel table with counts, group by (x, y), mean  -> pixel shows the abundance of the
//             canonical class of the k-mers listed at it (fold_kernel + the LUT gather below)
//   :903-913  NaN -> 0, A[x, y] = count + 1, transpose, flip      -> LUT is already in final orientation,
//             unused pixels (-1) stay 0
//   :916      bins = np.quantile(A, i/256, linear)   -> with the n sorted values s[], (p, g) = divmod((n-1)i, 256):
//             256*bins[i] = 256*s[p] + (s[min(p+1,n-1)] - s[p])*g, an exact integer
//   :917-919  out = digitize(A, bins) - 1 = #{i : bins[i] <= v} - 1  -> upper bound over the 256 scaled bins
// The float64 route of the reference is exact for these dyadic rationals (values < 2^45), so the integer
// restatement is bit-identical (tests/golden, oracle/image.py); no +-1 tolerance is needed.
#pragma once
#include <cooperative_groups.h>
#include "vk_common.cuh"

namespace vk {
namespace cg = cooperative_groups;

// canon[l][x] (lex index x) = sum over segments s >= l of fwd[s][i] + fwd[s][rc i], i = internal index of x.
// live_plan (the fused step, where n_levels is an upper bound): only the rows of the ladder are folded -- nobody reads the others.
__global__ void __launch_bounds__(256)
fold_kernel(const unsigned long long* __restrict__ seg_hist, int k, int n_levels, unsigned long long* __restrict__ canon,
            const Plan* __restrict__ live_plan)
{
    pdl_wait();
    const uint32_t nk = 1u << (2 * k);
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nk) return;
    if (live_plan && live_plan->n_levels < n_levels) n_levels = live_plan->n_levels;
    const uint32_t i = lex_to_internal(x, k);
    const uint32_t rc = internal_revcomp(i, k);
    unsigned long long run = 0;
    int s = n_levels - 1;
    for (; s >= 3; s -= 4) {                    // four rows' loads in flight together (the loop was one round trip per row)
        unsigned long long f[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            f[j] = seg_hist[(size_t)(s - j) * nk + i];
            if (rc != i) f[j] += seg_hist[(size_t)(s - j) * nk + rc];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            run += f[j];
            canon[(size_t)(s - j) * nk + x] = run;
        }
    }
    for (; s >= 0; --s) {
        unsigned long long f = seg_hist[(size_t)s * nk + i];
        if (rc != i) f += seg_hist[(size_t)s * nk + rc];
        run += f;
        canon[(size_t)s * nk + x] = run;
    }
}

__device__ __forceinline__ unsigned long long pixel_value(const unsigned long long* __restrict__ canon_l,
                                                          const int32_t* __restrict__ lut, uint32_t p)
{
    const int32_t x = lut[p];
    VK_ASSERT(x >= -1);
    return x < 0 ? 0ull : canon_l[x] + 1ull;
}

// Rank scaling without bins.  For a pixel value v that occurs in the image, with c(v) = #{pixels <= v} and
// n pixels, the digitize-against-quantiles of image.py:916-919 collapses to
//     out(v) = min(255, floor(256 * (c(v) - 1) / (n - 1)))
// (bins[i] <= v  <=>  (n-1) * i <= 256 * (c(v) - 1), because v itself is one of the sorted values; derivation in
// DESIGN.md "K4", checked against the reference's PNGs in tests/golden).  So the kernel only needs c(v).
//
// One thread-block CLUSTER of 8 CTAs per level (images up to 128 x 128): every CTA gathers and sorts one eighth of
// the pixels in its own shared memory (keys = value << 16 | pixel index, bitonic network), the eight sorted slices
// are exchanged through distributed shared memory, and each CTA ranks its own pixels with one upper-bound search
// per slice: c(v) = sum of the eight positions.  9 levels x 8 CTAs keep 72 SMs busy instead of 9.
constexpr int kImgCluster = 8;
constexpr uint32_t kImgIdxBits = 16;

__global__ void __cluster_dims__(kImgCluster, 1, 1) __launch_bounds__(1024, 1)
image_kernel_cluster(const unsigned long long* __restrict__ canon, const int32_t* __restrict__ lut, uint32_t nk,
                     uint32_t n_pix, uint32_t S, uint8_t* __restrict__ pixels, const Plan* __restrict__ live_plan)
{
    pdl_wait();
    // the fused step renders max_levels_out levels (the ladder is only known on the device): the clusters of the levels
    // beyond it leave at once and give their SMs to the other samples in flight (a 30 Mbp sample has 7 of 16)
    if (live_plan && blockIdx.y >= (uint32_t)live_plan->n_levels) return;
    extern __shared__ unsigned long long s_keys[];        // [kImgCluster][S]: slice r at offset r * S, in every CTA
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    const uint32_t level = blockIdx.y;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    const unsigned long long* canon_l = canon + (size_t)level * nk;
    unsigned long long* const own = s_keys + (size_t)rank * S;

    for (uint32_t i = tid; i < S; i += nthr) {
        const uint32_t p = rank * S + i;
        own[i] = p < n_pix ? (pixel_value(canon_l, lut, p) << kImgIdxBits) | p : ~0ull;
    }
    __syncthreads();
    // ---- bitonic sort of the slice, ascending.  blockDim.x = S / 2: a thread owns elements e0 = 64 warp + lane and
    // e0 + 32, so every comparator of stride <= 32 stays inside a warp: stride 32 inside the thread, strides 16..1 by
    // shuffle, all in registers.  Only the strides >= 64 of the sizes >= 128 go through shared memory (15 steps with a
    // block barrier for S = 2048, instead of 66 shared-memory steps).
    {
        constexpr uint32_t FULL = 0xffffffffu;
        VK_ASSERT(2 * nthr == S);
        const uint32_t e0 = 64u * (tid >> 5) + (tid & 31u), e1 = e0 + 32u;
        unsigned long long a = own[e0], b = own[e1];
        auto low_strides = [&](const uint32_t size, uint32_t stride) {            // strides stride..1 (stride <= 32)
            if (stride == 32u) {
                const bool up = (e0 & size) == 0;
                if ((a > b) == up) { const unsigned long long t = a; a = b; b = t; }
                stride = 16u;
            }
            for (; stride > 0; stride >>= 1) {
                const unsigned long long oa = __shfl_xor_sync(FULL, a, stride), ob = __shfl_xor_sync(FULL, b, stride);
                // keep the smaller key iff (this is the lower partner) == (the run is ascending)
                const bool lower = (e0 & stride) == 0;                              // same for e1 (stride < 32)
                const bool keep_min_a = lower == ((e0 & size) == 0), keep_min_b = lower == ((e1 & size) == 0);
                a = (oa < a) == keep_min_a ? oa : a;
                b = (ob < b) == keep_min_b ? ob : b;
            }
        };
        for (uint32_t size = 2; size <= 64u && size <= S; size <<= 1) low_strides(size, size >> 1);
        for (uint32_t size = 128; size <= S; size <<= 1) {
            own[e0] = a; own[e1] = b;
            __syncthreads();
            for (uint32_t stride = size >> 1; stride >= 64u; stride >>= 1) {
                const uint32_t lo = 2 * tid - (tid & (stride - 1));
                const uint32_t hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long x = own[lo], y = own[hi];
                if ((x > y) == up) { own[lo] = y; own[hi] = x; }
                __syncthreads();
            }
            a = own[e0]; b = own[e1];
            low_strides(size, 32u);
        }
        own[e0] = a; own[e1] = b;         // (a thread's pair was last read by the thread itself)
    }
    __syncthreads();
    // ---- all-gather of the sorted slices over distributed shared memory (all remote loads of a thread in flight at once)
    cluster.sync();
    for (uint32_t i = tid; i < S; i += nthr) {
        unsigned long long v[kImgCluster];
#pragma unroll
        for (uint32_t r = 0; r < kImgCluster; ++r) {
            const unsigned long long* src = cluster.map_shared_rank(s_keys + (size_t)r * S, r);
            v[r] = src[i];
        }
#pragma unroll
        for (uint32_t r = 0; r < kImgCluster; ++r)
            if (r != rank) s_keys[(size_t)r * S + i] = v[r];
    }
    cluster.sync();                                       // nobody leaves (or is overwritten) while peers still read
    // ---- c(v) = sum over slices of the upper bound of v, then the closed form.  The eight searches of an element
    // advance in lock-step (S is a power of two: log2 S rounds of eight independent shared-memory loads).
    uint8_t* out = pixels + (size_t)level * n_pix;
    for (uint32_t e = tid; e < S; e += nthr) {
        const unsigned long long key = own[e];
        if (key == ~0ull) continue;
        const unsigned long long vmax = key | ((1ull << kImgIdxBits) - 1);      // every key of the same value is <= this
        uint32_t pos[kImgCluster];
#pragma unroll
        for (uint32_t r = 0; r < kImgCluster; ++r) pos[r] = 0;
        for (uint32_t step = S >> 1; step > 0; step >>= 1) {
#pragma unroll
            for (uint32_t r = 0; r < kImgCluster; ++r)
                if (s_keys[(size_t)r * S + pos[r] + step - 1] <= vmax) pos[r] += step;
        }
        uint32_t c = 0;
#pragma unroll
        for (uint32_t r = 0; r < kImgCluster; ++r)             // pos = # of the first S-1 elements <= vmax; the last one:
            c += pos[r] + (s_keys[(size_t)r * S + pos[r]] <= vmax ? 1u : 0u);
        const uint32_t g = n_pix > 1 ? (256u * (c - 1)) / (n_pix - 1) : 0u;     // c <= 2^14: fits 32 bits
        out[(uint32_t)key & ((1u << kImgIdxBits) - 1)] = (uint8_t)(g > 255u ? 255u : g);
    }
}

// ---- large images (k = 8, 9): values sorted in global memory ------------------------------------------
__global__ void __launch_bounds__(256)
image_gather_kernel(const unsigned long long* __restrict__ canon, const int32_t* __restrict__ lut, uint32_t nk,
                    uint32_t n_pix, uint32_t n_pad, unsigned long long* __restrict__ vals)
{
    pdl_wait();
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t l = blockIdx.y;
    if (p >= n_pad) return;
    vals[(size_t)l * n_pad + p] = p < n_pix ? pixel_value(canon + (size_t)l * nk, lut, p) : ~0ull;
}

// one compare-exchange step of the bitonic network on every level (blockIdx.y)
__global__ void __launch_bounds__(256)
bitonic_global_step(unsigned long long* __restrict__ vals, uint32_t n_pad, uint32_t size, uint32_t stride)
{
    pdl_wait();
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (n_pad >> 1)) return;
    unsigned long long* v = vals + (size_t)blockIdx.y * n_pad;
    const uint32_t lo = 2 * t - (t & (stride - 1));
    const uint32_t hi = lo + stride;
    const bool up = (lo & size) == 0;
    const unsigned long long a = v[lo], b = v[hi];
    if ((a > b) == up) { v[lo] = b; v[hi] = a; }
}

// two consecutive steps (strides `stride` and `stride / 2`) in one pass: a thread owns the four elements they connect
__global__ void __launch_bounds__(256)
bitonic_global_step2(unsigned long long* __restrict__ vals, uint32_t n_pad, uint32_t size, uint32_t stride)
{
    pdl_wait();
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (n_pad >> 2)) return;
    unsigned long long* v = vals + (size_t)blockIdx.y * n_pad;
    const uint32_t h = stride >> 1;
    const uint32_t low = q & (h - 1);
    const uint32_t i = 4 * (q - low) + low;                    // bits log2(h) and log2(stride) of i are clear
    const bool up = (i & size) == 0;                           // size > stride: the same for the four
    unsigned long long x0 = v[i], x1 = v[i + h], x2 = v[i + stride], x3 = v[i + stride + h];
    unsigned long long t;
    if ((x0 > x2) == up) { t = x0; x0 = x2; x2 = t; }
    if ((x1 > x3) == up) { t = x1; x1 = x3; x3 = t; }
    if ((x0 > x1) == up) { t = x0; x0 = x1; x1 = t; }
    if ((x2 > x3) == up) { t = x2; x2 = x3; x3 = t; }
    v[i] = x0; v[i + h] = x1; v[i + stride] = x2; v[i + stride + h] = x3;
}

// Steps with stride < kSortTile of the global bitonic sort, one tile of kSortTile elements per CTA of 1024 threads.
// A thread owns elements e0 = 128 warp + lane and e0 + 32, + 64, + 96, so every comparator of stride <= 64 stays inside
// a warp: strides 64 and 32 inside the thread, 16..1 by shuffle, all in registers (as in image_kernel_cluster); only
// the strides >= 128 go through shared memory with a block barrier (15 of the 78 steps of the first launch, 5 of the 12
// of the later ones).
constexpr uint32_t kSortTile = 4096;
__global__ void __launch_bounds__(1024)
bitonic_tile_kernel(unsigned long long* __restrict__ vals, uint32_t n_pad, uint32_t size_begin, uint32_t size_end,
                    uint32_t first_stride)
{
    pdl_wait();
    // runs, for size = size_begin .. size_end (doubling), the strides min(size/2, first_stride) .. 1 inside a tile
    constexpr uint32_t FULL = 0xffffffffu;
    __shared__ unsigned long long s[kSortTile];
    unsigned long long* v = vals + (size_t)blockIdx.y * n_pad + (size_t)blockIdx.x * kSortTile;
    const uint32_t base = blockIdx.x * kSortTile;
    const uint32_t tid = threadIdx.x;
    VK_ASSERT(blockDim.x == 1024);
    for (uint32_t i = tid; i < kSortTile; i += 1024) s[i] = v[i];
    __syncthreads();
    const uint32_t e0 = 128u * (tid >> 5) + (tid & 31u);
    unsigned long long a[4];
    auto exchange = [&](unsigned long long& x, unsigned long long& y, bool up) {
        if ((x > y) == up) { const unsigned long long t = x; x = y; y = t; }
    };
    // strides `stride` (<= 64) .. 1 of one size, on the registers
    auto low_strides = [&](const uint32_t size, uint32_t stride) {
        if (stride == 64u) {
            exchange(a[0], a[2], ((base + e0) & size) == 0);
            exchange(a[1], a[3], ((base + e0 + 32u) & size) == 0);
            stride = 32u;
        }
        if (stride == 32u) {
            exchange(a[0], a[1], ((base + e0) & size) == 0);
            exchange(a[2], a[3], ((base + e0 + 64u) & size) == 0);
            stride = 16u;
        }
        for (; stride > 0; stride >>= 1) {
            const bool lower = (e0 & stride) == 0;                  // the same for the four elements (stride < 32)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const unsigned long long o = __shfl_xor_sync(FULL, a[r], stride);
                const bool keep_min = lower == (((base + e0 + 32u * r) & size) == 0);
                a[r] = (o < a[r]) == keep_min ? o : a[r];
            }
        }
    };
    bool in_regs = false;
    for (uint32_t size = size_begin; size <= size_end; size <<= 1) {
        uint32_t stride = size >> 1;
        if (stride > first_stride) stride = first_stride;
        if (stride >= 128u) {
            if (in_regs) {
#pragma unroll
                for (int r = 0; r < 4; ++r) s[e0 + 32u * r] = a[r];
                in_regs = false;
                __syncthreads();
            }
            for (; stride >= 128u; stride >>= 1) {
                for (uint32_t t = tid; t < kSortTile / 2; t += 1024) {
                    const uint32_t lo = 2 * t - (t & (stride - 1));
                    const uint32_t hi = lo + stride;
                    const bool up = ((base + lo) & size) == 0;
                    const unsigned long long x = s[lo], y = s[hi];
                    if ((x > y) == up) { s[lo] = y; s[hi] = x; }
                }
                __syncthreads();
            }
        }
        if (!in_regs) {
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = s[e0 + 32u * r];
            in_regs = true;
        }
        low_strides(size, stride);
    }
    // every thread last touched only its own four elements of s (or none): write them out from the registers
#pragma unroll
    for (int r = 0; r < 4; ++r) v[e0 + 32u * r] = a[r];
}

__global__ void __launch_bounds__(256)
image_bins_kernel(const unsigned long long* __restrict__ vals, uint32_t n_pix, uint32_t n_pad,
                  unsigned long long* __restrict__ bins)
{
    pdl_wait();
    const uint32_t i = threadIdx.x, l = blockIdx.x;
    const unsigned long long* s = vals + (size_t)l * n_pad;
    const uint64_t t = (uint64_t)(n_pix - 1) * i;
    const uint32_t p = (uint32_t)(t >> 8), g = (uint32_t)(t & 255u);
    const uint32_t q = p + 1 < n_pix ? p + 1 : n_pix - 1;
    bins[(size_t)l * 256 + i] = 256ull * s[p] + (s[q] - s[p]) * g;
}

__global__ void __launch_bounds__(256)
image_digitize_kernel(const unsigned long long* __restrict__ canon, const int32_t* __restrict__ lut, uint32_t nk,
                      uint32_t n_pix, const unsigned long long* __restrict__ bins, uint8_t* __restrict__ pixels)
{
    pdl_wait();
    __shared__ unsigned long long s_bins[256];
    const uint32_t l = blockIdx.y;
    s_bins[threadIdx.x] = bins[(size_t)l * 256 + threadIdx.x];
    __syncthreads();
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pix) return;
    const unsigned long long v = 256ull * pixel_value(canon + (size_t)l * nk, lut, p);
    uint32_t lo = 0, hi = 256;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (s_bins[mid] <= v) lo = mid + 1; else hi = mid;
    }
    pixels[(size_t)l * n_pix + p] = (uint8_t)(lo - 1);
}

// ---- remap between pixel tables (the "next" row N3 of SURVEY.md section 8f) ---------------------------------------
// convert.remap (varKoder/commands/convert.py:34-77): new[pixel_out(K)] = old[pixel_in(K)] over the inner join of the two
// tables; with sum_rc the contributions are ADDED in a uint8 array (np.add.at wraps modulo 256) and the result is
// rescaled as np.uint8((a - a.min()) / a.max() * 255) in float64.  The join is precomputed on the host as, per output
// pixel, two source pixels and their multiplicities (varkoder_b200/mapping.py: remap_plan).  One CTA per image.
__global__ void __launch_bounds__(256)
remap_kernel(const uint8_t* __restrict__ in, uint32_t n_in, const int32_t* __restrict__ src0,
             const int32_t* __restrict__ src1, const uint8_t* __restrict__ mult, uint32_t n_out, int sum_rc,
             uint8_t* __restrict__ out)
{
    pdl_wait();
    __shared__ uint32_t s_min, s_max;
    const uint8_t* inb = in + (size_t)blockIdx.x * n_in;
    uint8_t* outb = out + (size_t)blockIdx.x * n_out;
    if (!sum_rc) {
        for (uint32_t p = threadIdx.x; p < n_out; p += blockDim.x) {
            const int32_t s = src0[p];
            outb[p] = s >= 0 ? inb[s] : (uint8_t)0;
        }
        return;
    }
    if (threadIdx.x == 0) { s_min = 255u; s_max = 0u; }
    __syncthreads();
    uint32_t mn = 255u, mx = 0u;
    for (uint32_t p = threadIdx.x; p < n_out; p += blockDim.x) {
        const int32_t a = src0[p], b = src1[p];
        uint32_t v = 0;
        if (a >= 0) v = (uint32_t)mult[2 * p] * inb[a] + (b >= 0 ? (uint32_t)mult[2 * p + 1] * inb[b] : 0u);
        v &= 0xFFu;
        outb[p] = (uint8_t)v;
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) { atomicMin(&s_min, mn); atomicMax(&s_max, mx); }
    __syncthreads();
    mn = s_min;
    mx = s_max;
    for (uint32_t p = threadIdx.x; p < n_out; p += blockDim.x) {
        const uint32_t v = outb[p];
        // float64, in numpy's order of operations; IEEE division and multiplication, no contraction possible
        outb[p] = mx ? (uint8_t)(int)(__dmul_rn(__ddiv_rn((double)(v - mn), (double)mx), 255.0)) : (uint8_t)0;
    }
}

}  // namespace vk
