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
#include "vk_common.cuh"

namespace vk {

// canon[l][x] (lex index x) = sum over segments s >= l of fwd[s][i] + fwd[s][rc i], i = internal index of x.
__global__ void __launch_bounds__(256)
fold_kernel(const unsigned long long* __restrict__ seg_hist, int k, int n_levels, unsigned long long* __restrict__ canon)
{
    const uint32_t nk = 1u << (2 * k);
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nk) return;
    const uint32_t i = lex_to_internal(x, k);
    const uint32_t rc = internal_revcomp(i, k);
    unsigned long long run = 0;
    for (int s = n_levels - 1; s >= 0; --s) {
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
    return x < 0 ? 0ull : canon_l[x] + 1ull;
}

// Rank scaling without bins.  For a pixel value v that occurs in the image, with c(v) = #{pixels <= v} and
// n pixels, the digitize-against-quantiles of image.py:916-919 collapses to
//     out(v) = min(255, floor(256 * (c(v) - 1) / (n - 1)))
// (bins[i] <= v  <=>  (n-1) * i <= 256 * (c(v) - 1), because v itself is one of the sorted values; derivation in
// DESIGN.md "K4", checked against the reference's PNGs in tests/golden).  So the kernel only needs c(v): sort the
// n values once, then an upper bound per pixel.
//
// One CTA per level, 16 values per thread: in-register bitonic network for the 16, then log2(n/16) merge-path
// levels through shared memory (each thread finds its diagonal by binary search and merges 16 outputs serially).
constexpr int kImageItems = 16;
// shared-memory index of logical element i: one pad word per 16, so that threads owning consecutive 16-element
// blocks (stride 17 x 8 B) do not all land on one bank (the unpadded layout was a 32-way conflict: 177 us -> see profiles)
__device__ __forceinline__ uint32_t sidx(uint32_t i) { return i + (i >> 4); }

__device__ __forceinline__ void cex(unsigned long long& a, unsigned long long& b, bool up)
{
    const bool sw = (a > b) == up;
    const unsigned long long lo = sw ? b : a, hi = sw ? a : b;
    a = lo;
    b = hi;
}

__global__ void __launch_bounds__(1024, 1)
image_kernel_smem(const unsigned long long* __restrict__ canon, const int32_t* __restrict__ lut, uint32_t nk,
                  uint32_t n_pix, uint32_t n_pad, uint8_t* __restrict__ pixels)
{
    extern __shared__ unsigned long long s_val[];        // n_pad values
    const unsigned long long* canon_l = canon + (size_t)blockIdx.x * nk;
    const uint32_t tid = threadIdx.x;                     // blockDim.x == n_pad / 16
    const uint32_t base = tid * kImageItems;

    unsigned long long v[kImageItems];
#pragma unroll
    for (int i = 0; i < kImageItems; ++i) v[i] = base + i < n_pix ? pixel_value(canon_l, lut, base + i) : ~0ull;

    // ---- 16 values in registers: bitonic sorting network, ascending
#pragma unroll
    for (int size = 2; size <= kImageItems; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int i = 0; i < kImageItems; ++i) {
                const int l = i ^ stride;
                if (l > i) cex(v[i], v[l], (i & size) == 0);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kImageItems; ++i) s_val[sidx(base) + i] = v[i];

    // ---- merge sorted runs pairwise: run = 16, 32, ..., n_pad / 2
    for (uint32_t run = kImageItems; run < n_pad; run <<= 1) {
        __syncthreads();
        const uint32_t pair_base = base & ~(2 * run - 1);
        const uint32_t A = pair_base, B = pair_base + run;      // logical offsets of the two runs
        const uint32_t diag = base - pair_base;           // this thread emits merged outputs diag .. diag+15
        uint32_t lo = diag > run ? diag - run : 0, hi = diag < run ? diag : run;
        while (lo < hi) {                                 // merge path: first a with A[a] > B[diag-1-a]
            const uint32_t mid = (lo + hi) >> 1;
            if (s_val[sidx(A + mid)] <= s_val[sidx(B + diag - 1 - mid)]) lo = mid + 1; else hi = mid;
        }
        uint32_t a = lo, b = diag - lo;
        unsigned long long ka = a < run ? s_val[sidx(A + a)] : ~0ull, kb = b < run ? s_val[sidx(B + b)] : ~0ull;
#pragma unroll
        for (int i = 0; i < kImageItems; ++i) {
            const bool take_a = b >= run || (a < run && ka <= kb);
            v[i] = take_a ? ka : kb;
            if (take_a) { ++a; ka = a < run ? s_val[sidx(A + a)] : ~0ull; }
            else { ++b; kb = b < run ? s_val[sidx(B + b)] : ~0ull; }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kImageItems; ++i) s_val[sidx(base) + i] = v[i];
    }
    __syncthreads();

    // ---- c(v) = upper bound over the n_pix real values (the padding sorts last), then the closed form
    uint8_t* out = pixels + (size_t)blockIdx.x * n_pix;
    for (uint32_t p = tid; p < n_pix; p += blockDim.x) {
        const unsigned long long val = pixel_value(canon_l, lut, p);
        uint32_t lo = 0, hi = n_pix;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_val[sidx(mid)] <= val) lo = mid + 1; else hi = mid;
        }
        const uint32_t g = n_pix > 1 ? (256u * (lo - 1)) / (n_pix - 1) : 0u;        // lo <= 2^14: fits 32 bits
        out[p] = (uint8_t)(g > 255u ? 255u : g);
    }
}

// ---- large images (k = 8, 9): values sorted in global memory ------------------------------------------
__global__ void __launch_bounds__(256)
image_gather_kernel(const unsigned long long* __restrict__ canon, const int32_t* __restrict__ lut, uint32_t nk,
                    uint32_t n_pix, uint32_t n_pad, unsigned long long* __restrict__ vals)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t l = blockIdx.y;
    if (p >= n_pad) return;
    vals[(size_t)l * n_pad + p] = p < n_pix ? pixel_value(canon + (size_t)l * nk, lut, p) : ~0ull;
}

// one compare-exchange step of the bitonic network on every level (blockIdx.y)
__global__ void __launch_bounds__(256)
bitonic_global_step(unsigned long long* __restrict__ vals, uint32_t n_pad, uint32_t size, uint32_t stride)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (n_pad >> 1)) return;
    unsigned long long* v = vals + (size_t)blockIdx.y * n_pad;
    const uint32_t lo = 2 * t - (t & (stride - 1));
    const uint32_t hi = lo + stride;
    const bool up = (lo & size) == 0;
    const unsigned long long a = v[lo], b = v[hi];
    if ((a > b) == up) { v[lo] = b; v[hi] = a; }
}

// all steps with stride < TILE/2... handled in shared memory: TILE elements per CTA
constexpr uint32_t kSortTile = 4096;
__global__ void __launch_bounds__(1024)
bitonic_tile_kernel(unsigned long long* __restrict__ vals, uint32_t n_pad, uint32_t size_begin, uint32_t size_end,
                    uint32_t first_stride)
{
    // runs, for size = size_begin .. size_end (doubling), the strides min(size/2, first_stride) .. 1 inside a tile
    __shared__ unsigned long long s[kSortTile];
    unsigned long long* v = vals + (size_t)blockIdx.y * n_pad + (size_t)blockIdx.x * kSortTile;
    const uint32_t base = blockIdx.x * kSortTile;
    for (uint32_t i = threadIdx.x; i < kSortTile; i += blockDim.x) s[i] = v[i];
    __syncthreads();
    for (uint32_t size = size_begin; size <= size_end; size <<= 1) {
        uint32_t stride = size >> 1;
        if (stride > first_stride) stride = first_stride;
        for (; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < kSortTile / 2; t += blockDim.x) {
                const uint32_t lo = 2 * t - (t & (stride - 1));
                const uint32_t hi = lo + stride;
                const bool up = ((base + lo) & size) == 0;
                const unsigned long long a = s[lo], b = s[hi];
                if ((a > b) == up) { s[lo] = b; s[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < kSortTile; i += blockDim.x) v[i] = s[i];
}

__global__ void __launch_bounds__(256)
image_bins_kernel(const unsigned long long* __restrict__ vals, uint32_t n_pix, uint32_t n_pad,
                  unsigned long long* __restrict__ bins)
{
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

}  // namespace vk
