// primitives.cu -- device-wide building blocks: exclusive scan, stable LSD radix sort of
// (u64 key, u32 value) pairs, finite bounding box.  Used by the map build (K1) and the voxel
// grid (K2).  Hand-written; no CUB/Thrust.
#include <cfloat>

#include "common.cuh"

namespace ssf {

// =========================================================================================
// Exclusive scan (u32).  256 threads x 16 items per block, block totals scanned recursively.
// =========================================================================================
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// block-wide exclusive scan of one value per thread (256 threads); returns exclusive prefix,
// *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t *total)
{
    __shared__ uint32_t warp_sums[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = warp_incl_scan(v);
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < 8 ? warp_sums[lane] : 0;
        uint32_t wi = warp_incl_scan(w);
        if (lane < 8) warp_sums[lane] = wi - w;  // exclusive
        if (lane == 7) *total = wi;
    }
    __syncthreads();
    uint32_t r = incl - v + warp_sums[warp];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_kernel(const uint32_t *__restrict__ in,
                                                                 uint32_t *__restrict__ out, size_t n,
                                                                 uint32_t *__restrict__ tile_sums)
{
    __shared__ uint32_t s_total;
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        sum += v[i];
    }
    uint32_t prefix = block_excl_scan_256(sum, &s_total);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = prefix;
        prefix += v[i];
    }
    if (threadIdx.x == 0 && tile_sums) tile_sums[blockIdx.x] = s_total;
}

__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(uint32_t *__restrict__ out, size_t n,
                                                                const uint32_t *__restrict__ tile_offsets)
{
    const uint32_t off = tile_offsets[blockIdx.x];
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) out[base + i] += off;
}

__global__ void scan_total_kernel(const uint32_t *__restrict__ in_last, const uint32_t *__restrict__ out_last,
                                  uint32_t *__restrict__ total)
{
    *total = *in_last + *out_last;
}

static int scan_rec(const uint32_t *in, uint32_t *out, size_t n, Scratch &s, int level, cudaStream_t st)
{
    const size_t tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles <= 1) {
        scan_tile_kernel<<<1, kScanThreads, 0, st>>>(in, out, n, nullptr);
        SSF_LAUNCHED();
        return SSF_OK;
    }
    if (level >= 3) {
        set_error("exclusive_scan_u32: input too large (%zu)", n);
        return SSF_ERR_INVALID;
    }
    SSF_TRY(s.scan_tmp[level].reserve(tiles));
    uint32_t *sums = s.scan_tmp[level].p;
    scan_tile_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(in, out, n, sums);
    SSF_LAUNCHED();
    SSF_TRY(scan_rec(sums, sums, tiles, s, level + 1, st));
    scan_add_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(out, n, sums);
    SSF_LAUNCHED();
    return SSF_OK;
}

int exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t n, uint32_t *total_dev, Scratch &s, cudaStream_t st)
{
    if (n == 0) {
        if (total_dev) SSF_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(uint32_t), st));
        return SSF_OK;
    }
    if (total_dev && in == out) {
        set_error("exclusive_scan_u32: total needs in != out");
        return SSF_ERR_INVALID;
    }
    SSF_TRY(scan_rec(in, out, n, s, 0, st));
    if (total_dev) {
        scan_total_kernel<<<1, 1, 0, st>>>(in + (n - 1), out + (n - 1), total_dev);
        SSF_LAUNCHED();
    }
    return SSF_OK;
}

// =========================================================================================
// Stable LSD radix sort, 8-bit digits.  Per pass: per-tile digit histogram -> exclusive scan
// over (digit, tile) -> stable scatter.  Stability inside a tile comes from a fixed
// (warp, round, lane) order with __match_any_sync ranking.
// =========================================================================================
constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;  // rounds per warp
constexpr int kSortTile = kSortThreads * kSortItems;
static_assert(kSortTile == kSortTileSize, "common.cuh announces the sort tile size");

__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const unsigned long long *__restrict__ keys, size_t n,
                                                                 int shift, uint32_t *__restrict__ ghist,
                                                                 uint32_t n_tiles)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * kSortTile;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        size_t idx = base + (size_t)i * kSortThreads + threadIdx.x;
        if (idx < n) atomicAdd(&h[(uint32_t)(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    ghist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kSortThreads)
    sort_scatter_kernel(const unsigned long long *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                        unsigned long long *__restrict__ keys_out, uint32_t *__restrict__ vals_out, size_t n, int shift,
                        const uint32_t *__restrict__ ghist_scanned, uint32_t n_tiles)
{
    __shared__ uint32_t wh[kSortThreads / 32][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (kSortThreads / 32) * 256; i += kSortThreads) (&wh[0][0])[i] = 0;
    __syncthreads();

    const size_t wbase = (size_t)blockIdx.x * kSortTile + (size_t)warp * (32 * kSortItems);
    unsigned long long k[kSortItems];
    uint32_t v[kSortItems];
    uint32_t rank[kSortItems];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const size_t idx = wbase + (size_t)r * 32 + lane;
        const bool ok = idx < n;
        k[r] = ok ? keys_in[idx] : 0ull;
        v[r] = ok ? vals_in[idx] : 0u;
        // inactive lanes get a digit outside 0..255 so they never match an active one
        const uint32_t digit = ok ? ((uint32_t)(k[r] >> shift) & 255u) : 256u + (uint32_t)lane;
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        uint32_t prev = 0;
        if (ok) prev = wh[warp][digit];
        __syncwarp();
        if (ok && (peers & lt_mask) == 0) wh[warp][digit] = prev + __popc(peers);
        __syncwarp();
        rank[r] = prev + __popc(peers & lt_mask);
    }
    __syncthreads();
    {
        // thread d owns digit d: turn per-warp counts into global output offsets
        const int d = threadIdx.x;
        uint32_t run = ghist_scanned[(size_t)d * n_tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; ++w) {
            uint32_t c = wh[w][d];
            wh[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const size_t idx = wbase + (size_t)r * 32 + lane;
        if (idx < n) {
            const uint32_t digit = (uint32_t)(k[r] >> shift) & 255u;
            const uint32_t pos = wh[warp][digit] + rank[r];
            keys_out[pos] = k[r];
            vals_out[pos] = v[r];
        }
    }
}

// The histogram kernel reads keys tile-strided by thread while the scatter kernel reads them
// warp-striped; both cover exactly the same tile, so the per-tile digit counts agree.

int radix_sort_pairs_u64(unsigned long long *keys, uint32_t *vals, size_t n, int bits, Scratch &s, cudaStream_t st)
{
    if (n <= 1 || bits <= 0) return SSF_OK;
    if (n >= (size_t)1 << 32) {
        set_error("radix_sort_pairs_u64: n too large");
        return SSF_ERR_INVALID;
    }
    const uint32_t n_tiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
    SSF_TRY(s.hist.reserve((size_t)256 * n_tiles));
    SSF_TRY(s.keys_alt.reserve(n));
    SSF_TRY(s.vals_alt.reserve(n));
    unsigned long long *kin = keys, *kout = s.keys_alt.p;
    uint32_t *vin = vals, *vout = s.vals_alt.p;
    const int passes = (bits + 7) / 8;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        sort_hist_kernel<<<n_tiles, kSortThreads, 0, st>>>(kin, n, shift, s.hist.p, n_tiles);
        SSF_LAUNCHED();
        SSF_TRY(exclusive_scan_u32(s.hist.p, s.hist.p, (size_t)256 * n_tiles, nullptr, s, st));
        sort_scatter_kernel<<<n_tiles, kSortThreads, 0, st>>>(kin, vin, kout, vout, n, shift, s.hist.p, n_tiles);
        SSF_LAUNCHED();
        unsigned long long *tk = kin; kin = kout; kout = tk;
        uint32_t *tv = vin; vin = vout; vout = tv;
    }
    if (kin != keys) {
        SSF_CUDA(cudaMemcpyAsync(keys, kin, n * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
        SSF_CUDA(cudaMemcpyAsync(vals, vin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
    return SSF_OK;
}

// =========================================================================================
// Segmented stable LSD radix sort of (u32 key, u32 value) pairs, 8-bit digits, 4 passes.
// The array is a sequence of segments, each a whole number of sort tiles; seg[t] = (first tile
// of t's segment, tiles in it).  The digit counts are laid out [segment][digit][tile in segment],
// so ONE global exclusive scan yields, for every (tile, digit), the output position that keeps
// each segment in its own slot range and orders it by digit, then tile, then rank -- i.e. an
// independent stable sort per segment, all segments in the same launches.
// =========================================================================================
__global__ void __launch_bounds__(kSortThreads)
    seg_hist_kernel(const uint32_t *__restrict__ keys, int shift, const uint2 *__restrict__ seg,
                    uint32_t *__restrict__ ghist)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * kSortTile;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i)
        atomicAdd(&h[(keys[base + (size_t)i * kSortThreads + threadIdx.x] >> shift) & 255u], 1u);
    __syncthreads();
    const uint2 sg = seg[blockIdx.x];
    ghist[(size_t)256 * sg.x + (size_t)threadIdx.x * sg.y + (blockIdx.x - sg.x)] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kSortThreads)
    seg_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                       uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int shift,
                       const uint2 *__restrict__ seg, const uint32_t *__restrict__ ghist_scanned)
{
    // The tile is first ordered by digit in shared memory (stable: warp slices in input order, ranks
    // within a warp in round and lane order), then written out in that order: the lanes of a store
    // cover a few digit runs -- a few sectors -- instead of 32 scattered 4-byte targets (the
    // low-digit passes, where every lane of a warp drew a different digit, took twice the time of
    // the high-digit ones).
    __shared__ uint32_t wh[kSortThreads / 32][256];
    __shared__ uint32_t s_k[kSortTile], s_v[kSortTile];
    __shared__ uint32_t s_gbase[256], s_dstart[256], s_wsum[kSortThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (kSortThreads / 32) * 256; i += kSortThreads) (&wh[0][0])[i] = 0;
    __syncthreads();
    const size_t wbase = (size_t)blockIdx.x * kSortTile + (size_t)warp * (32 * kSortItems);
    uint32_t k[kSortItems], v[kSortItems], rank[kSortItems];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const size_t idx = wbase + (size_t)r * 32 + lane;
        k[r] = keys_in[idx];
        v[r] = vals_in[idx];
    }
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t digit = (k[r] >> shift) & 255u;
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        const uint32_t prev = wh[warp][digit];
        __syncwarp();
        if ((peers & lt_mask) == 0) wh[warp][digit] = prev + __popc(peers);
        __syncwarp();
        rank[r] = prev + __popc(peers & lt_mask);
    }
    __syncthreads();
    {
        // thread d owns digit d: per-warp counts -> offsets of the warps inside the digit's run; the
        // tile's digit totals -> start of each run in the tile (block exclusive scan)
        static_assert(kSortThreads == 256, "one thread per digit");
        const uint2 sg = seg[blockIdx.x];
        const int d = threadIdx.x;
        s_gbase[d] = ghist_scanned[(size_t)256 * sg.x + (size_t)d * sg.y + (blockIdx.x - sg.x)];
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; ++w) {
            const uint32_t c = wh[w][d];
            wh[w][d] = run;
            run += c;
        }
        const uint32_t incl = warp_incl_scan(run);
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; ++w) before += w < warp ? s_wsum[w] : 0u;
        s_dstart[d] = before + incl - run;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t digit = (k[r] >> shift) & 255u;
        const uint32_t lp = s_dstart[digit] + wh[warp][digit] + rank[r];
        SSF_CHECK(lp < (uint32_t)kSortTile);
        s_k[lp] = k[r];
        s_v[lp] = v[r];
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < (uint32_t)kSortTile; i += kSortThreads) {
        const uint32_t key = s_k[i], digit = (key >> shift) & 255u;
        const uint32_t pos = s_gbase[digit] + (i - s_dstart[digit]);
        SSF_CHECK(pos >= seg[blockIdx.x].x * kSortTile && pos < (seg[blockIdx.x].x + seg[blockIdx.x].y) * kSortTile);
        keys_out[pos] = key;
        vals_out[pos] = s_v[i];
    }
}

// in-place exclusive scan of every segment's [digit][tile] counts, starting at the segment's base
// offset (a segment keeps its slot range): one 1024-thread block per segment, 16 counts per thread and trip
constexpr int kSegScanThreads = 1024;
constexpr int kSegScanItems = 16;
__global__ void __launch_bounds__(kSegScanThreads)
    seg_scan_kernel(uint32_t *__restrict__ ghist, const uint4 *__restrict__ segs, uint32_t entries_per_tile)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_total;
    const uint4 sg = segs[blockIdx.x];  // first tile, tiles, base offset
    uint32_t *h = ghist + (size_t)entries_per_tile * sg.x;
    const uint32_t n = entries_per_tile * sg.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t carry = sg.z;  // the same in every thread
    for (uint32_t off = 0; off < n; off += kSegScanThreads * kSegScanItems) {
        const uint32_t base = off + threadIdx.x * kSegScanItems;
        uint32_t v[kSegScanItems], sum = 0;
#pragma unroll
        for (int i = 0; i < kSegScanItems; ++i) {
            v[i] = base + i < n ? h[base + i] : 0u;
            sum += v[i];
        }
        const uint32_t incl = warp_incl_scan(sum);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = s_warp[lane];
            const uint32_t wi = warp_incl_scan(w);
            s_warp[lane] = wi - w;  // exclusive prefix of the warp totals
            if (lane == 31) s_total = wi;
        }
        __syncthreads();
        uint32_t prefix = incl - sum + s_warp[warp] + carry;
#pragma unroll
        for (int i = 0; i < kSegScanItems; ++i) {
            if (base + i < n) h[base + i] = prefix;
            prefix += v[i];
        }
        carry += s_total;
        __syncthreads();
    }
}

int seg_scan_u32(uint32_t *counts, const uint4 *segs, uint32_t n_segs, uint32_t entries_per_tile, cudaStream_t st)
{
    if (n_segs == 0) return SSF_OK;
    seg_scan_kernel<<<n_segs, kSegScanThreads, 0, st>>>(counts, segs, entries_per_tile);
    SSF_LAUNCHED();
    return SSF_OK;
}

int seg_radix_sort_pairs_u32(uint32_t *keys, uint32_t *vals, size_t n, const uint2 *seg_of_tile, const uint4 *segs,
                             uint32_t n_segs, Scratch &s, cudaStream_t st)
{
    if (n == 0) return SSF_OK;
    if (n % kSortTile != 0 || n >= (size_t)1 << 32) {
        set_error("seg_radix_sort_pairs_u32: n must be a multiple of %d and below 2^32", kSortTile);
        return SSF_ERR_INVALID;
    }
    const uint32_t n_tiles = (uint32_t)(n / kSortTile);
    SSF_TRY(s.hist.reserve((size_t)256 * n_tiles));
    SSF_TRY(s.keys_alt.reserve(n / 2 + 1));
    SSF_TRY(s.vals_alt.reserve(n));
    uint32_t *kin = keys, *kout = reinterpret_cast<uint32_t *>(s.keys_alt.p);
    uint32_t *vin = vals, *vout = s.vals_alt.p;
    for (int p = 0; p < 4; ++p) {  // an even number of passes: the result ends up in keys / vals
        const int shift = 8 * p;
        seg_hist_kernel<<<n_tiles, kSortThreads, 0, st>>>(kin, shift, seg_of_tile, s.hist.p);
        SSF_LAUNCHED();
        SSF_TRY(seg_scan_u32(s.hist.p, segs, n_segs, 256, st));
        seg_scatter_kernel<<<n_tiles, kSortThreads, 0, st>>>(kin, vin, kout, vout, shift, seg_of_tile, s.hist.p);
        SSF_LAUNCHED();
        uint32_t *tk = kin; kin = kout; kout = tk;
        uint32_t *tv = vin; vin = vout; vout = tv;
    }
    return SSF_OK;
}

// =========================================================================================
// Finite bounding box
// =========================================================================================
__global__ void bbox_init_kernel(int *bbox_ord, uint32_t *n_finite)
{
    if (threadIdx.x < 3) bbox_ord[threadIdx.x] = float_ordered(FLT_MAX);
    else if (threadIdx.x < 6) bbox_ord[threadIdx.x] = float_ordered(-FLT_MAX);
    if (threadIdx.x == 0) *n_finite = 0;
}

__global__ void __launch_bounds__(256) bbox_kernel(const float4 *__restrict__ pts, size_t n, int *bbox_ord,
                                                   uint32_t *n_finite)
{
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    uint32_t cnt = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
            mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
            mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
            ++cnt;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], d));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], d));
        }
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&bbox_ord[k], float_ordered(mn[k]));
            atomicMax(&bbox_ord[3 + k], float_ordered(mx[k]));
        }
        atomicAdd(n_finite, cnt);
    }
}

__global__ void bbox_decode_kernel(int *bbox_ord)
{
    if (threadIdx.x < 6) {
        float f = ordered_float(bbox_ord[threadIdx.x]);
        reinterpret_cast<float *>(bbox_ord)[threadIdx.x] = f;
    }
}

int bbox_finite(const float4 *pts, size_t n, float *bbox_dev, uint32_t *n_finite_dev, cudaStream_t st)
{
    int *ord = reinterpret_cast<int *>(bbox_dev);
    bbox_init_kernel<<<1, 32, 0, st>>>(ord, n_finite_dev);
    SSF_LAUNCHED();
    if (n > 0) {
        unsigned blocks = (unsigned)((n + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        bbox_kernel<<<blocks, 256, 0, st>>>(pts, n, ord, n_finite_dev);
        SSF_LAUNCHED();
    }
    bbox_decode_kernel<<<1, 32, 0, st>>>(ord);
    SSF_LAUNCHED();
    return SSF_OK;
}

}  // namespace ssf
