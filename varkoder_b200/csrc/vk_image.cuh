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

// One CTA per level; the n = side^2 values (<= NPAD, a power of two) are sorted in shared memory (bitonic).
constexpr int kImageThreads = 1024;

__global__ void __launch_bounds__(kImageThreads, 1)
image_kernel_smem(const unsigned long long* __restrict__ canon, const int32_t* __restrict__ lut, uint32_t nk,
                  uint32_t n_pix, uint32_t n_pad, uint8_t* __restrict__ pixels)
{
    extern __shared__ unsigned long long s_val[];        // n_pad sorted values, then 256 bins
    unsigned long long* s_bins = s_val + n_pad;
    const unsigned long long* canon_l = canon + (size_t)blockIdx.x * nk;
    const uint32_t tid = threadIdx.x;

    for (uint32_t p = tid; p < n_pad; p += kImageThreads)
        s_val[p] = p < n_pix ? pixel_value(canon_l, lut, p) : ~0ull;
    __syncthreads();

    // bitonic sort, ascending
    for (uint32_t size = 2; size <= n_pad; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = tid; t < (n_pad >> 1); t += kImageThreads) {
                const uint32_t lo = 2 * t - (t & (stride - 1));      // index with the stride bit cleared
                const uint32_t hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = s_val[lo], b = s_val[hi];
                if ((a > b) == up) { s_val[lo] = b; s_val[hi] = a; }
            }
            __syncthreads();
        }
    }

    if (tid < 256) {
        const uint64_t t = (uint64_t)(n_pix - 1) * tid;
        const uint32_t p = (uint32_t)(t >> 8), g = (uint32_t)(t & 255u);
        const uint32_t q = p + 1 < n_pix ? p + 1 : n_pix - 1;
        s_bins[tid] = 256ull * s_val[p] + (s_val[q] - s_val[p]) * g;
    }
    __syncthreads();

    uint8_t* out = pixels + (size_t)blockIdx.x * n_pix;
    for (uint32_t p = tid; p < n_pix; p += kImageThreads) {
        const unsigned long long v = 256ull * pixel_value(canon_l, lut, p);
        // number of bins <= v (bins are non-decreasing, bins[0] <= v always)
        uint32_t lo = 0, hi = 256;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_bins[mid] <= v) lo = mid + 1; else hi = mid;
        }
        out[p] = (uint8_t)(lo - 1);
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
