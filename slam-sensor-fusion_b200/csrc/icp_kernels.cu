// icp_kernels.cu -- K3/K4/K5: the registration loop on the device.
//
// Replaces ICPPointToPoint::calculateAlignment and its helpers (reference
// localization/src/icp_point_to_point.cpp:57-84, 99-110, 112-159, 161-170, 185-254) for a
// batch of independent scans against one HBM-resident map.  No host round trip inside an
// alignment: the data-dependent control flow (abort, early break, lazy re-search) lives in
// per-scan device state, and every kernel of the fixed launch sequence exits early for
// scans that are done.
//
//   search_accum_kernel<KIND, THREADS>   persistent blocks over the tiles that hold points: TMA tile load,
//                                         transform, certificate check or exact walk (near part, then the
//                                         far rings packed densely; the warps of a block claim chunks of 32
//                                         queued walks as they finish), residual / Jacobian partial sums
//                                         (Gauss-Newton: an 8x8 Gram matrix on the FP64 tensor core, gn_gram)
//   rowsum_solve_kernel                   ordered sum of a scan's partial rows + 6x6 Cholesky or 3x3 SVD +
//                                         pose update + stop rules (map-sharded: rowsum_xchg_solve_kernel
//                                         exchanges the rows across ranks over peer memory inside the same
//                                         launch; or rowsum_kernel + ncclAllReduce / the caller's hook +
//                                         solve_kernel)
//   ref_search_kernel / ref_reduce_kernel<RT, CT> / ref_step_kernel   the reference's own state machine,
//                                         STRICT (sequential float chains) or FAST; two block shapes
//   run_batch                             the launch sequence, captured once per shape into a CUDA graph
#include <climits>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "icp.cuh"
#include "nn_device.cuh"
#include "small_math.cuh"

namespace ssf {

#ifdef SSF_NN_STATS
extern "C" int ssf_debug_nn_stats(unsigned long long *out, int reset)
{
    if (cudaMemcpyFromSymbol(out, g_nn_stats, sizeof(g_nn_stats)) != cudaSuccess) return -2;
    if (reset) {
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(g_nn_stats, z, sizeof(z));
    }
    return 0;
}
#endif

// =========================================================================================
// state init / results
// =========================================================================================
__global__ void init_states_kernel(ScanState *st, const float *T_init, uint32_t n_scans, float *trace_err,
                                   int32_t *trace_search, int trace_len, float *pose_hist)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_scans) return;
    ScanState &z = st[s];
    for (int i = 0; i < 16; ++i) {
        const float v = T_init[16 * s + i];
        z.T[i] = v;
        z.T_init[i] = v;
        pose_hist[((size_t)s * kCertHist) * 16 + i] = v;  // pose of search launch 0
        z.T_step[i] = (i % 5 == 0) ? 1.f : 0.f;
    }
    z.last_error = FLT_MAX;
    z.error = 1e6f;
    z.pend_error = 0.f;
    z.iterations = 0;
    z.done = 0;
    z.converged = 0;
    z.aborted = 0;
    z.n_searches = 0;
    z.k_last = 0;
    z.need_search = 0;
    z.have_step = 0;
    z.fitness = 0.0;
    z.rmse = 0.0;
    z.last_step = 3.0e38f;
    z.comm_error = 0;
    for (int i = 0; i < trace_len; ++i) {
        trace_err[(size_t)s * trace_len + i] = nanf("");
        trace_search[(size_t)s * trace_len + i] = 0;
    }
}

int init_states(BatchBuffers &b, const float *T_init_dev, cudaStream_t st)
{
    if (b.n_scans == 0) return SSF_OK;
    init_states_kernel<<<(unsigned)((b.n_scans + 127) / 128), 128, 0, st>>>(b.state.p, T_init_dev, (uint32_t)b.n_scans,
                                                                          b.trace_err.p, b.trace_search.p, b.trace_len,
                                                                          b.pose_hist.p);
    SSF_LAUNCHED();
    return SSF_OK;
}

__global__ void no_points_kernel(ScanState *st, uint32_t n_scans, int mode)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_scans) return;
    ScanState &z = st[s];
    z.done = 1;
    z.n_searches = 1;
    if (mode == SSF_MODE_O3D_P2P) z.error = 0.f;
    else z.aborted = 1;
}

__global__ void results_kernel(ScanState *st, ssf_icp_result *out, uint32_t n_scans, int mode, float acc_err)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_scans) return;
    ScanState &z = st[s];
    ssf_icp_result r;
    if (mode == SSF_MODE_REFERENCE && !z.aborted) {
        // cpp:249-253: error = last_error_, has_converged = last_error_ < acceptable_mean_error_
        z.error = z.last_error;
        z.converged = z.last_error < acc_err ? 1 : 0;
    }
    const bool ab = z.aborted || z.comm_error;
    for (int i = 0; i < 16; ++i) r.transformation[i] = ab ? z.T_init[i] : z.T[i];
    r.error = ab ? 1e6f : z.error;
    r.iterations = ab ? 0 : z.iterations;
    r.has_converged = ab ? 0 : z.converged;
    r.n_searches = z.n_searches;
    r.k_final = z.k_last;
    r.aborted = z.comm_error ? 2 : z.aborted;  // 2: a peer's sums never arrived (ssf_batch_results -> SSF_ERR_COMM)
    r.n_source = (int32_t)z.n_pts;
    r.fitness = z.n_pts ? (float)((double)z.k_last / (double)z.n_pts) : 0.f;
    r.device_ms = 0.f;
    out[s] = r;
}

// =========================================================================================
// K3 + K4 fused: transform, exact NN, rejection, partial sums (GN and Open3D-flow modes)
// =========================================================================================
enum AccumKind { ACC_GN_P2P = 0, ACC_GN_P2PLANE = 1, ACC_KABSCH = 2 };

// layout of a partial row (kAccum doubles)
//   GN:     [0..20] upper triangle of J^T J (row-major), [21..26] J^T r, [27] sum r^2, [28] K
//   KABSCH: [0] K, [1..3] sum (p-c), [4..6] sum (q-c), [7..15] sum (p-c)(q-c)^T (row r, col c),
//           [16] sum |p-q|^2, c = translation of the initial transform (pivot against cancellation)

// ---- 1-D TMA bulk copy of a scan tile into shared memory --------------------------------------
// One thread arms the mbarrier with the byte count and issues cp.async.bulk (SASS UBLKCP); the
// copy engine moves the tile while the block loads the pose, and every thread then waits on the
// barrier phase.  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tile_load_issue(void *smem_dst, unsigned long long *bar, const void *gmem_src,
                                                uint32_t bytes)
{
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    // the tile is read once per launch: evict-first in L2 (like the certificate loads and the correspondence /
    // certificate stores, __ldcs / __stcs), so that it does not push out the map the gathers hit -- converged
    // launches 131 -> 127 us
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(gmem_src), "r"(bytes), "r"(b), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void tile_bar_init(unsigned long long *bar)
{
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tile_bar_wait(unsigned long long *bar, uint32_t phase)
{
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(b), "r"(phase)
                     : "memory");
    }
}

// sharded maps: a tile left out of this rank's work lists is never written, so every row starts as
// "owned by another rank, no certificate"
__global__ void __launch_bounds__(256) reset_rows_kernel(int32_t *__restrict__ corr, uint2 *__restrict__ cert, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    corr[i] = -2;
    cert[i] = make_uint2(0u, 0xFFFFFFFFu);
}

// ... and no tile carries the stamp of an earlier run (or of whatever the allocation held before)
__global__ void __launch_bounds__(256) reset_stamps_kernel(double *__restrict__ partials, uint32_t n_tiles)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_tiles) partials[(size_t)t * kAccum + (kAccum - 1)] = 0.0;
}

__global__ void zero_u32_kernel(uint32_t *p, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

// ---- map sharding: which tiles can hold a query this rank owns ------------------------------------
// A rank owns the queries whose shard column floor((x' - ox) * inv_h) lies in [own_lo, own_hi), x' being
// the x coordinate of the TRANSFORMED source point.  tile_box holds the bounding box of every tile's
// source points in the sensor frame; moved by the scan's current pose it gives an interval of x' that
// contains every query of the tile (interval arithmetic, widened by the rounding of the per-point
// transform), hence a column interval (cell_coord is monotone).  A tile whose interval misses the
// rank's columns is left out of the search kernel's work list: nothing is loaded or transformed for it.
__device__ __forceinline__ bool tile_may_own(const ShardView &sv, const float *T, uint32_t tile)
{
    if (!sv.enabled) return true;
    const float4 lo = sv.tile_box[2 * (size_t)tile], hi = sv.tile_box[2 * (size_t)tile + 1];
    const float c[3] = {T[0], T[4], T[8]}, l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
    float xmin = T[12], xmax = T[12], mag = fabsf(T[12]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float a = c[k] * l[k], b = c[k] * h[k];
        xmin += fminf(a, b);
        xmax += fmaxf(a, b);
        mag += fmaxf(fabsf(a), fabsf(b));
    }
    if (!(mag < 3.0e38f)) return true;  // non-finite points in the tile (their rows are answered by the first rank)
    const float slack = 4e-6f * mag + 1e-5f;
    const int clo = cell_coord(xmin - slack, sv.ox, sv.inv_h, 1 << 24), chi = cell_coord(xmax + slack, sv.ox, sv.inv_h, 1 << 24);
    return chi >= sv.own_lo && clo < sv.own_hi;
}

// bounding box of every tile's source points (sensor frame); a tile with a non-finite point gets an
// infinite box.  One 128-thread block per tile.
__global__ void __launch_bounds__(128)
    tile_box_kernel(const float4 *__restrict__ src, const uint32_t *__restrict__ tile_scan,
                    const ScanState *__restrict__ states, float4 *__restrict__ tile_box)
{
    __shared__ float smin[4][3], smax[4][3];
    const uint32_t tile = blockIdx.x;
    const ScanState &z = states[tile_scan[tile]];
    const uint32_t k = tile - z.tile_begin;
    if ((size_t)k * kTile >= z.n_pts) return;
    const uint32_t n_here = min((uint32_t)kTile, z.n_pts - k * kTile);
    const float4 *p = src + (size_t)z.pt_begin + (size_t)k * kTile;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    bool bad = false;
    for (uint32_t i = threadIdx.x; i < n_here; i += 128) {
        const float4 q = p[i];
        bad |= !(isfinite(q.x) && isfinite(q.y) && isfinite(q.z));
        mn[0] = fminf(mn[0], q.x); mn[1] = fminf(mn[1], q.y); mn[2] = fminf(mn[2], q.z);
        mx[0] = fmaxf(mx[0], q.x); mx[1] = fmaxf(mx[1], q.y); mx[2] = fmaxf(mx[2], q.z);
    }
    bad = __syncthreads_or(bad);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], d));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], d));
        }
    if ((threadIdx.x & 31) == 0)
        for (int c = 0; c < 3; ++c) { smin[threadIdx.x >> 5][c] = mn[c]; smax[threadIdx.x >> 5][c] = mx[c]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float inf = __int_as_float(0x7f800000);
        for (int c = 0; c < 3; ++c) {
            mn[c] = fminf(fminf(smin[0][c], smin[1][c]), fminf(smin[2][c], smin[3][c]));
            mx[c] = fmaxf(fmaxf(smax[0][c], smax[1][c]), fmaxf(smax[2][c], smax[3][c]));
            if (bad) { mn[c] = -inf; mx[c] = inf; }
        }
        tile_box[2 * (size_t)tile] = make_float4(mn[0], mn[1], mn[2], 0.f);
        tile_box[2 * (size_t)tile + 1] = make_float4(mx[0], mx[1], mx[2], 0.f);
    }
}

// The work list of a search launch: one 16-byte record per tile in use -- (tile, scan, first slot,
// points) -- so the search kernel reaches its data in one hop.  A scan's slots are sized for its RAW
// points; after the voxel stage only the first ceil(n_pts / kTile) tiles are in use.  Called by one
// warp per scan; sharded maps list only the tiles that can hold an owned query at pose T.
__device__ __forceinline__ void list_scan_tiles(const ScanState &z, uint32_t scan, const float *T, const ShardView &sv,
                                                uint4 *__restrict__ active, uint32_t *__restrict__ n_active)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n = (z.n_pts + kTile - 1) / kTile;
    for (uint32_t k0 = 0; k0 < n; k0 += 32) {
        const uint32_t k = k0 + lane;
        const bool on = k < n && tile_may_own(sv, T, z.tile_begin + k);
        const uint32_t bal = __ballot_sync(0xffffffffu, on);
        uint32_t base = 0;
        if (lane == 0 && bal) base = atomicAdd(n_active, (uint32_t)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (on)
            active[base + __popc(bal & ((1u << lane) - 1u))] =
                make_uint4(z.tile_begin + k, scan, z.pt_begin + k * kTile, min((uint32_t)kTile, z.n_pts - k * kTile));
    }
}

__global__ void __launch_bounds__(128)
    active_tiles_kernel(const ScanState *__restrict__ states, uint32_t n_scans, ShardView sv, uint4 *__restrict__ active,
                        uint32_t *__restrict__ n_active)
{
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per scan
    if (s >= n_scans) return;
    list_scan_tiles(states[s], s, states[s].T, sv, active, n_active);
}

// One half (values 16*HALF .. 16*HALF+15 of the partial row) of the Gauss-Newton sums over this
// thread's queries of the tile, reduced over the warp: lane l returns the warp-wide sum of value
// 16*HALF + (l & 15).  Row layout: [0..20] J^T J upper triangle, [21..26] J^T r, [27] sum r^2, [28] K.
template <int KIND, int HALF, int THREADS>
__device__ __forceinline__ double gn_half(const MapView &map, const float4 *s_q, const uint32_t *s_pos, uint32_t n_here)
{
    constexpr int kQ_ = kTile / THREADS;
    constexpr int lo = 16 * HALF;
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.0;
    for (int k = 0; k < kQ_; ++k) {
        const uint32_t r = (uint32_t)k * THREADS + threadIdx.x;
        if (r >= n_here) break;
        const uint32_t pos = s_pos[r];
        if (pos == kNoPos) continue;
        const float4 p = s_q[r];
        float4 q, nf;
        if (KIND == ACC_GN_P2PLANE) ld_point_normal(map.pn + 2 * (size_t)pos, q, nf);
        else q = __ldg(&map.pts[pos]);
        const double px = p.x, py = p.y, pz = p.z;
        const double e[3] = {px - (double)q.x, py - (double)q.y, pz - (double)q.z};
        if (KIND == ACC_GN_P2PLANE) {
            const double n[3] = {nf.x, nf.y, nf.z};
            const double a[6] = {fma(py, n[2], -(pz * n[1])), fma(pz, n[0], -(px * n[2])),
                                 fma(px, n[1], -(py * n[0])), n[0], n[1], n[2]};
            int t = 0;
#pragma unroll
            for (int u = 0; u < 6; ++u)
#pragma unroll
                for (int w = u; w < 6; ++w, ++t)
                    if (t >= lo && t < lo + 16) v[t - lo] = fma(a[u], a[w], v[t - lo]);
            if (HALF == 1) {
                const double rs = fma(n[0], e[0], fma(n[1], e[1], n[2] * e[2]));
#pragma unroll
                for (int u = 0; u < 6; ++u) v[21 - lo + u] = fma(a[u], rs, v[21 - lo + u]);
                v[27 - lo] = fma(rs, rs, v[27 - lo]);
            }
        } else {
            // J = [-[p]x | I]; J^T J = [[ -[p]x^T -[p]x , [p]x ], [ -[p]x , I ]]
            const double J[3][6] = {{0, pz, -py, 1, 0, 0}, {-pz, 0, px, 0, 1, 0}, {py, -px, 0, 0, 0, 1}};
            int t = 0;
#pragma unroll
            for (int u = 0; u < 6; ++u)
#pragma unroll
                for (int w = u; w < 6; ++w, ++t)
                    if (t >= lo && t < lo + 16) v[t - lo] += J[0][u] * J[0][w] + J[1][u] * J[1][w] + J[2][u] * J[2][w];
            if (HALF == 1) {
#pragma unroll
                for (int u = 0; u < 6; ++u) v[21 - lo + u] += J[0][u] * e[0] + J[1][u] * e[1] + J[2][u] * e[2];
                v[27 - lo] += e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
            }
        }
        if (HALF == 1) v[28 - lo] += 1.0;
    }
    return warp_transpose_reduce16(v);
}

// ---- Gauss-Newton sums as a Gram matrix on the FP64 tensor core ------------------------------------
// The 29 sums of a partial row are the upper triangle of F F^T, F = one 8-vector of features per
// residual: (a0..a5, r, 1) with a = J^T row and r the residual -- J^T J = sum a a^T, J^T r = sum a r,
// sum r^2 and the count K all fall out of one 8x8 product.  mma.sync m8n8k4 (f64) adds four
// residuals per instruction into an accumulator of two doubles per lane, so a thread keeps 2
// accumulators instead of 29 and the warp needs no shuffle reduction at all.
// Fragment layout (PTX ISA, mma.m8n8k4 .f64): lane l holds A[l >> 2][l & 3], B[l & 3][l >> 2] and
// C[l >> 2][2 (l & 3) + {0, 1}]; with A = F (feature x residual) and B = F^T both operands of a
// lane are the same value F[l >> 2][4 j + (l & 3)].
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// one feature vector per lane -> F F^T added to (c0, c1).  F is staged in the warp's 2 KB tile,
// [feature][lane ^ (feature << 2)]: the swizzle makes the stores and the fragment loads free of
// bank conflicts (a fragment load touches columns 4 (j ^ g) + t: all 32 of a row pair).
__device__ __forceinline__ void gram_round(double *F, const double (&f)[8], double &c0, double &c1)
{
    const uint32_t lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    __syncwarp();  // the previous round's fragment loads are done
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i) F[i * 32 + (lane ^ (i << 2))] = f[i];
    __syncwarp();
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j) {
        const double x = F[g * 32 + (((j ^ g) << 2) | t)];
        dmma_m8n8k4(c0, c1, x, x);
    }
}

// partial-row index of Gram entry (R, col), R <= col; -1 for entries the row does not keep.
// Row layout: [0..20] J^T J upper triangle (row-major), [21..26] J^T r, [27] sum r^2, [28] K.
__device__ __forceinline__ int gram_slot(int R, int col)
{
    if (R > col) return -1;
    if (col <= 5) return R * 6 - R * (R - 1) / 2 + (col - R);
    if (col == 6) return R <= 5 ? 21 + R : 27;
    return R == 7 ? 28 : -1;
}

// the Gauss-Newton sums of this warp's queries of the tile -> row[0..31] (shared memory)
template <int KIND, int THREADS>
__device__ __forceinline__ void gn_gram(const MapView &map, const float4 *s_q, const uint32_t *s_pos, uint32_t n_here,
                                        double *F, double *row)
{
    constexpr int kQ_ = kTile / THREADS;
    const uint32_t lane = threadIdx.x & 31;
    double c0 = 0.0, c1 = 0.0;
    // (loading the gathers of all four queries up front was measured slower: 132 -> 143 us per
    // converged launch)
    for (int k = 0; k < kQ_; ++k) {
        const uint32_t r = (uint32_t)k * THREADS + threadIdx.x;
        const uint32_t pos = r < n_here ? s_pos[r] : kNoPos;
        const bool hit = pos != kNoPos;
        if (!__any_sync(0xffffffffu, hit)) continue;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f), q = p, nf = p;
        if (hit) {
            p = s_q[r];
            if (KIND == ACC_GN_P2PLANE) ld_point_normal(map.pn + 2 * (size_t)pos, q, nf);
            else q = __ldg(&map.pts[pos]);
        }
        const double px = p.x, py = p.y, pz = p.z;
        const double e[3] = {px - (double)q.x, py - (double)q.y, pz - (double)q.z};
        const double one = hit ? 1.0 : 0.0;
        if (KIND == ACC_GN_P2PLANE) {
            const double n[3] = {nf.x, nf.y, nf.z};
            const double f[8] = {fma(py, n[2], -(pz * n[1])), fma(pz, n[0], -(px * n[2])), fma(px, n[1], -(py * n[0])),
                                 n[0], n[1], n[2], fma(n[0], e[0], fma(n[1], e[1], n[2] * e[2])), one};
            gram_round(F, f, c0, c1);
        } else {
            // J = [-[p]x | I]: three residual rows per query; the count rides on the first
            const double f0[8] = {0.0, pz, -py, one, 0.0, 0.0, e[0], one};
            const double f1[8] = {-pz, 0.0, px, 0.0, one, 0.0, e[1], 0.0};
            const double f2[8] = {py, -px, 0.0, 0.0, 0.0, one, e[2], 0.0};
            gram_round(F, f0, c0, c1);
            gram_round(F, f1, c0, c1);
            gram_round(F, f2, c0, c1);
        }
    }
    const int R = (int)(lane >> 2), col = 2 * (int)(lane & 3);
    if (lane < 3) row[29 + lane] = 0.0;
    const int s0 = gram_slot(R, col), s1 = gram_slot(R, col + 1);
    if (s0 >= 0) row[s0] = c0;
    if (s1 >= 0) row[s1] = c1;
}

// first index of the next chunk of 32 items of a list the warps of a block consume together (counter in
// shared memory, one atomic per warp and chunk).  Collective: every lane of the warp calls it.
__device__ __forceinline__ uint32_t warp_claim(uint32_t *counter)
{
    uint32_t base = 0;
    if ((threadIdx.x & 31u) == 0) base = atomicAdd(counter, 32u);
    return __shfl_sync(0xffffffffu, base, 0);
}

// Persistent blocks fetch tiles (kTile consecutive queries of one scan) from a shared counter;
// every thread takes kQ queries of a tile.
//   V  transform; with use_cert, try to confirm last iteration's neighbour from its certificate
//      (nn_verify) -- a handful of flops and one gather instead of a search;
//   S  the queries that could not be confirmed, packed densely over the threads: full exact walk,
//      new certificate;
//   K4 residual / Jacobian terms of the matched queries: gn_gram (128-thread blocks, tensor core),
//      gn_half x 2 (512-thread blocks) or the Kabsch moments with a 32-value warp transpose.
// resident 128-thread blocks per SM (64 registers each).  Measured on the 256-scan step: 6 -> 40.3 k scans/s,
// 7 -> 42.6 k, 8 -> 42.8 k, 9 (56 registers) -> 40.4 k, 10 (48 registers) -> 40.7 k: more blocks take shared
// memory away from the L1 the gathers live in (converged launches 135 -> 169 us at 9)
#ifndef SSF_MINB
#define SSF_MINB 8
#endif
template <int KIND, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 128 ? SSF_MINB : 2)
    search_accum_kernel(MapView map, const float4 *__restrict__ src, const uint32_t *__restrict__ tile_scan,
                        const ScanState *__restrict__ states, float limit, int32_t *__restrict__ corr,
                        double *__restrict__ partials, uint2 *__restrict__ cert, const float *__restrict__ pose_hist,
                        int use_cert, int pass, const uint4 *__restrict__ active,
                        const uint32_t *__restrict__ n_active, uint32_t *__restrict__ fetch,
                        unsigned long long *__restrict__ stats)
{
    __shared__ __align__(128) float4 s_q[kTile];  // TMA destination; transformed in place, w = owned by this rank
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_pos[kTile];
    // walk state per query (running best, second-smallest d2) and the two work lists; all of it is dead
    // once the correspondences are written, when the same 8 KB hold the feature tiles of the Gauss-Newton
    // Gram matrix (gram_round).  Shared memory per block decides the L1 the walks and gathers live in:
    // 8 blocks of < 19.5 KB fit the 164 KB carve-out (92 KB of L1 per SM); at 21.6 KB the driver took 196 KB
    __shared__ __align__(16) unsigned char s_raw[kTile * 16];
    unsigned long long *const s_key = reinterpret_cast<unsigned long long *>(s_raw);
    float *const s_b2 = reinterpret_cast<float *>(s_raw + kTile * 8);
    unsigned short *const s_queue = reinterpret_cast<unsigned short *>(s_raw + kTile * 12);
    unsigned short *const s_far = reinterpret_cast<unsigned short *>(s_raw + kTile * 14);
    __shared__ uint32_t s_nq, s_nfar, s_nfar_none, s_next;
    __shared__ uint32_t s_take[2];  // next unclaimed chunk of 32 queued queries: near walks, far walks
    __shared__ float sT[16];
    constexpr int kQ_ = kTile / THREADS;  // queries per thread
    __shared__ double sred[THREADS / 32][kAccum];
    const uint32_t none_hi = __float_as_uint(limit);
    const bool searchable = limit > 0.f && map.n_pts > 0;
    const uint32_t n_tiles = *n_active;
    // the next tile is claimed while this tile's sums are taken (late, so that a block in a long
    // walk does not sit on a tile another block could start): the atomic's round trip is hidden
    uint32_t nxt = 0;
    if (threadIdx.x == 0) {
        tile_bar_init(&s_bar);
        nxt = atomicAdd(fetch, 1u);
    }
    uint32_t phase = 0;
    while (true) {
        if (threadIdx.x == 0) {
            s_next = nxt;
            s_nq = 0;
            s_nfar = 0;
            s_nfar_none = 0;
            s_take[0] = 0;
            s_take[1] = 0;
        }
        __syncthreads();
        const uint32_t ti = s_next;
        if (ti >= n_tiles) break;
        const uint4 rec = active[ti];  // (tile, scan, first slot, points)
        const uint32_t tile = rec.x, scan = rec.y, n_here = rec.w;
        const size_t slot0 = rec.z;
        if (threadIdx.x == 0) tile_load_issue(s_q, &s_bar, src + slot0, n_here * (uint32_t)sizeof(float4));
        const ScanState &z = states[scan];
        SSF_CHECK(n_here > 0 && n_here <= kTile && slot0 == (size_t)z.pt_begin + (size_t)(tile - z.tile_begin) * kTile);
        const bool done = z.done != 0;
        // a certificate only pays off if the next pose update is small: write them once the last
        // update was below a few margins (the updates shrink fast)
        const bool make_cert = z.last_step < map.cert_step && pass < kCertHist;
        const float *hist = pose_hist + (size_t)scan * kCertHist * 16;
        if (threadIdx.x < 16) sT[threadIdx.x] = z.T[threadIdx.x];
        // certificates of this thread's queries: issue the loads before waiting for the tile
        uint2 crt[kQ_];
#pragma unroll
        for (int k = 0; k < kQ_; ++k) {
            const uint32_t r = (uint32_t)k * THREADS + threadIdx.x;
            crt[k] = make_uint2(0u, kNoPos);
            if (use_cert && r < n_here) crt[k] = __ldcs(&cert[slot0 + r]);
        }
        __syncthreads();
        tile_bar_wait(&s_bar, phase);
        phase ^= 1u;
        if (done) {
            if (threadIdx.x == 0) nxt = atomicAdd(fetch, 1u);
            __syncthreads();  // s_next is rewritten at the top
            continue;
        }
        // ---- V ----
#pragma unroll
        for (int k = 0; k < kQ_; ++k) {
            const uint32_t r = (uint32_t)k * THREADS + threadIdx.x;
            if (r >= n_here) continue;
            const float4 s4 = s_q[r];
            const float3 p = transform_point(sT, s4.x, s4.y, s4.z);
            // map sharding: only the rank that owns the query's column searches it, so every
            // correspondence is counted exactly once across ranks (the halo covers the radius)
            const int col = cell_coord(p.x, map.shard_ox, map.shard_inv_h, 1 << 24);
            const bool mine = col >= map.own_lo && col < map.own_hi;
            const bool ok = searchable && isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
            s_q[r] = make_float4(p.x, p.y, p.z, mine ? 1.f : 0.f);
            s_pos[r] = kNoPos;
            if (!mine || !ok) {
                corr[slot0 + r] = mine ? -1 : -2;
                cert[slot0 + r] = make_uint2(0u, kNoPos);  // no certificate (corr[] no longer holds a neighbour)
                continue;
            }
            const float L = __uint_as_float(crt[k].x & ~31u);
            if (L > 0.f) {
                // where this query was when the certificate was issued: same source point, the pose
                // of that search launch (bit-identical to the position searched then)
                const float3 pc = transform_point(hist + (crt[k].x & 31u) * 16, s4.x, s4.y, s4.z);
                unsigned long long key;
                if (nn_verify(map, pc.x, pc.y, pc.z, L, crt[k].y, p.x, p.y, p.z, limit, key)) {
                    // same neighbour as before: corr[] already holds its index
                    if ((uint32_t)(key >> 32) < none_hi) s_pos[r] = crt[k].y;
                    continue;
                }
            }
            s_pos[r] = use_cert ? crt[k].y : kNoPos;  // last iteration's neighbour seeds the walk
            s_queue[atomicAdd(&s_nq, 1u)] = (unsigned short)r;
        }
        __syncthreads();
        // search statistics of this launch: [0] queries answered, [1] queries that needed a walk
        if (threadIdx.x == 0) {
            atomicAdd(&stats[0], (unsigned long long)n_here);
            atomicAdd(&stats[1], (unsigned long long)s_nq);
        }
        // ---- S: near part of the walk for every queued query; the few that must go on to rings 2..
        // are queued again and finished afterwards, packed densely, so that a warp is not held up
        // by the lanes that drew a far query ----
        // (a warp claims the next 32 queued queries when it is through with its last ones: the walks differ
        // in length, and a fixed share per warp left the others waiting at the barrier -- first launch of a
        // 256-scan step 1.90 -> 1.46 ms.  Dropping the barrier between the two parts as well, with warps
        // starting on a far list that is still being filled, was measured slower: 1.54 ms)
        const uint32_t nq = s_nq;
        for (uint32_t base = warp_claim(&s_take[0]); base < nq; base = warp_claim(&s_take[0])) {
            const uint32_t i = base + (threadIdx.x & 31u);
            if (i >= nq) continue;
            const uint32_t r = s_queue[i];
            SSF_CHECK(r < n_here);
            const float4 p = s_q[r];
            unsigned long long key;
            uint32_t pos;
            float b2 = 0.f;
            bool far;
            if (make_cert) {
                NNBest<true> B;
                far = nn_walk_near<true>(map, p.x, p.y, p.z, limit, map.cert_mu, B, s_pos[r]);
                key = B.key; pos = B.pos; b2 = B.b2;
            } else {
                NNBest<false> B;
                far = nn_walk_near<false>(map, p.x, p.y, p.z, limit, 0.f, B, s_pos[r]);
                key = B.key; pos = B.pos;
            }
            s_key[r] = key;
            s_pos[r] = pos;
            s_b2[r] = b2;
            // two classes, filled from the two ends of the list: queries that have found nothing yet
            // walk every row within the threshold radius and cost several times the others, so
            // they are packed together
            if (far) {
                if ((uint32_t)(key >> 32) < none_hi) s_far[atomicAdd(&s_nfar, 1u)] = (unsigned short)r;
                else s_far[kTile - 1 - atomicAdd(&s_nfar_none, 1u)] = (unsigned short)r;
            }
        }
        __syncthreads();
        const uint32_t nfar = s_nfar, nfar_none = s_nfar_none;
        for (uint32_t base = warp_claim(&s_take[1]); base < nfar + nfar_none; base = warp_claim(&s_take[1])) {
            const uint32_t i = base + (threadIdx.x & 31u);
            if (i >= nfar + nfar_none) continue;
            const uint32_t r = i < nfar_none ? s_far[kTile - 1 - i] : s_far[i - nfar_none];
            const float4 p = s_q[r];
            if (make_cert) {
                NNBest<true> B;
                B.key = s_key[r]; B.pos = s_pos[r]; B.b2 = s_b2[r]; B.mu = map.cert_mu;
                // the seed of the near part: last iteration's neighbour, still in the certificate array
                B.skip = use_cert ? cert[slot0 + r].y : kNoPos;
                B.refresh();
                nn_walk_far_flat<true>(map, p.x, p.y, p.z, B);
                s_key[r] = B.key; s_pos[r] = B.pos; s_b2[r] = B.b2;
            } else {
                NNBest<false> B;
                B.key = s_key[r]; B.pos = s_pos[r];
                B.skip = kNoPos;
                nn_walk_far_flat<false>(map, p.x, p.y, p.z, B);
                s_key[r] = B.key; s_pos[r] = B.pos;
            }
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nq; i += THREADS) {
            const uint32_t r = s_queue[i];
            const float4 p = s_q[r];
            const unsigned long long key = s_key[r];
            const bool hit = (uint32_t)(key >> 32) < none_hi;
            __stcs(&corr[slot0 + r], hit ? (int)(uint32_t)key : -1);
            float radius = 0.f;
            if (make_cert) {
                NNBest<true> B;
                B.key = key; B.b2 = s_b2[r]; B.mu = map.cert_mu;
                radius = cert_radius(B);
            }
            __stcs(&cert[slot0 + r], make_uint2(cert_pack(radius, pass), hit ? s_pos[r] : kNoPos));
            if (!hit) s_pos[r] = kNoPos;
            else NN_STAT(5, 1);
        }
        __syncthreads();
        // ---- K4: residual / Jacobian terms of the matched queries ----
        if (threadIdx.x == 0) nxt = atomicAdd(fetch, 1u);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if constexpr (KIND == ACC_KABSCH) {
            double v[kAccum];
#pragma unroll
            for (int i = 0; i < kAccum; ++i) v[i] = 0.0;
            for (int k = 0; k < kQ_; ++k) {
                const uint32_t r = (uint32_t)k * THREADS + threadIdx.x;
                if (r >= n_here) break;
                const uint32_t pos = s_pos[r];
                if (pos == kNoPos) continue;
                const float4 p = s_q[r];
                const float4 q = __ldg(&map.pts[pos]);
                const double cx = z.T_init[12], cy = z.T_init[13], cz = z.T_init[14];
                const double a[3] = {(double)p.x - cx, (double)p.y - cy, (double)p.z - cz};
                const double b[3] = {(double)q.x - cx, (double)q.y - cy, (double)q.z - cz};
                v[0] += 1.0;
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    v[1 + t] += a[t];
                    v[4 + t] += b[t];
                }
#pragma unroll
                for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[7 + 3 * rr + c] = fma(a[rr], b[c], v[7 + 3 * rr + c]);
                const double ex = (double)p.x - q.x, ey = (double)p.y - q.y, ez = (double)p.z - q.z;
                v[16] += fma(ex, ex, fma(ey, ey, ez * ez));
            }
            sred[warp][lane] = warp_transpose_reduce32(v);
        } else {
            // the 29 sums are taken in two halves of 16 accumulators (32 registers each instead of
            // 64): the second half re-reads the neighbour (an L1 hit) and recomputes the six-vector
            if constexpr (THREADS == kThreads) {
                gn_gram<KIND, THREADS>(map, s_q, s_pos, n_here, reinterpret_cast<double *>(s_raw) + warp * 256, sred[warp]);
            } else {
                const double lo = gn_half<KIND, 0, THREADS>(map, s_q, s_pos, n_here);
                const double hi = gn_half<KIND, 1, THREADS>(map, s_q, s_pos, n_here);
                if (lane < 16) {
                    sred[warp][lane] = lo;
                    sred[warp][16 + lane] = hi;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < kAccum) {
            double sum = 0.0;
#pragma unroll
            for (int k = 0; k < THREADS / 32; ++k) sum += sred[k][threadIdx.x];
            // slot 31 (unused by every mode) carries the pass that wrote the row: a tile left out of a
            // sharded rank's work list keeps an older stamp and is not added (scan_rowsum)
            partials[(size_t)tile * kAccum + threadIdx.x] = threadIdx.x == kAccum - 1 ? (double)(pass + 1) : sum;
        }
        // (the barrier at the top of the loop orders these reads before the next tile's writes)
    }
}

// ordered sum of the partial rows of one scan (rows stamped with this pass only).  8 warps take interleaved
// rows, then the 8 partial totals are added in warp order: fixed order, deterministic.  Returns the total of
// value `lane` in warp 0 (other warps: undefined).  256 threads.
__device__ __forceinline__ double scan_rowsum(const ScanState &z, const double *__restrict__ partials, int pass,
                                              double (*sw)[kAccum])
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double s = 0.0;
    if (!z.done) {
        const uint32_t n_tiles = (z.n_pts + kTile - 1) / kTile;
        const double stamp = (double)(pass + 1);
        for (uint32_t t = warp; t < n_tiles; t += 8) {
            const double v = partials[(size_t)(z.tile_begin + t) * kAccum + lane];
            if (__shfl_sync(0xffffffffu, v, kAccum - 1) == stamp) s += v;
        }
    }
    sw[warp][lane] = s;
    __syncthreads();
    double tot = 0.0;
    if (warp == 0) {
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += sw[w][lane];
    }
    return tot;
}

__global__ void __launch_bounds__(256) rowsum_kernel(const ScanState *__restrict__ states,
                                                     const double *__restrict__ partials, double *__restrict__ sums,
                                                     int pass)
{
    __shared__ double sw[8][kAccum];
    const double tot = scan_rowsum(states[blockIdx.x], partials, pass, sw);
    if (threadIdx.x < 32) sums[(size_t)blockIdx.x * kAccum + threadIdx.x] = tot;
}

// one thread: normal equations -> Cholesky -> pose update -> stop rules (sv = the scan's kAccum totals)
__device__ void gn_solve(ScanState &z, const double *sv, int pass, float acc_err, float eps, float *hist, float *trace)
{
    const long long K = (long long)(sv[28] + 0.5);
    z.n_searches += 1;
    z.k_last = (int)K;
    if (pass == 0 && K < 10) {  // same guard as the reference's first search (cpp:196-200)
        z.aborted = 1;
        z.done = 1;
        return;
    }
    if (K < 6) { z.done = 1; return; }
    const float err = (float)sqrt(sv[27] / (double)K);
    z.error = err;
    if (trace) trace[pass] = err;  // debug trace: error measured by pass `pass`
    if (err < acc_err) { z.converged = 1; z.done = 1; return; }
    double A[36], nb[6], x[6];
    int t = 0;
    for (int u = 0; u < 6; ++u)
        for (int w = u; w < 6; ++w) {
            A[u * 6 + w] = sv[t];
            A[w * 6 + u] = sv[t];
            ++t;
        }
    for (int u = 0; u < 6; ++u) nb[u] = -sv[21 + u];
    if (!cholesky_solve6(A, nb, x)) { z.done = 1; return; }
    double Ts[16];
    se3_from_twist(x, Ts);
    compose_round(Ts, z.T);
    if (pass + 1 < kCertHist)
        for (int i = 0; i < 16; ++i) hist[(pass + 1) * 16 + i] = z.T[i];  // pose of search launch pass + 1
    z.iterations += 1;
    double mx = 0.0;
    for (int u = 0; u < 6; ++u) mx = fmax(mx, fabs(x[u]));
    z.last_step = (float)fmax(fmax(fabs(x[3]), fmax(fabs(x[4]), fabs(x[5]))),
                              30.0 * fmax(fabs(x[0]), fmax(fabs(x[1]), fabs(x[2]))));
    if (mx < (double)eps) { z.converged = 1; z.done = 1; }
}

__device__ void o3d_solve(ScanState &z, const double *sv, int pass, int max_iteration, float *hist, float *trace)
{
    const long long K = (long long)(sv[0] + 0.5);
    z.n_searches += 1;
    z.k_last = (int)K;
    const double prev_fit = z.fitness, prev_rmse = z.rmse;
    z.fitness = z.n_pts ? (double)K / (double)z.n_pts : 0.0;
    z.rmse = K > 0 ? sqrt(sv[16] / (double)K) : 0.0;
    z.error = (float)z.rmse;
    if (trace) trace[pass] = z.error;
    if (pass > 0 && fabs(prev_fit - z.fitness) < 1e-6 && fabs(prev_rmse - z.rmse) < 1e-6) {
        z.converged = 1;
        z.done = 1;
        return;
    }
    if (pass == max_iteration || K == 0) { z.done = 1; return; }
    // un-pivot: centroids and centred cross-covariance
    const double c[3] = {z.T_init[12], z.T_init[13], z.T_init[14]};
    double ma[3], mb[3], sp[3], sq[3], H[9];
    for (int k = 0; k < 3; ++k) {
        ma[k] = sv[1 + k] / (double)K;
        mb[k] = sv[4 + k] / (double)K;
        sp[k] = ma[k] + c[k];
        sq[k] = mb[k] + c[k];
    }
    for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) H[cc * 3 + r] = sv[7 + 3 * r + cc] - (double)K * ma[r] * mb[cc];
    double Ts[16];
    kabsch_from_moments_d(sp, sq, H, Ts);
    {   // step applied to a point 30 m from the pivot: |Ts * y - y| with y = pivot + (30, 30, 30) / sqrt(3)
        double mv = 0.0;
        for (int r = 0; r < 3; ++r) {
            double acc = Ts[12 + r];
            for (int cc = 0; cc < 3; ++cc) acc += (Ts[cc * 4 + r] - (r == cc ? 1.0 : 0.0)) * (c[cc] + 17.32);
            mv = fmax(mv, fabs(acc));
        }
        z.last_step = (float)mv;
    }
    compose_round(Ts, z.T);
    if (pass + 1 < kCertHist)
        for (int i = 0; i < 16; ++i) hist[(pass + 1) * 16 + i] = z.T[i];  // pose of search launch pass + 1
    z.iterations += 1;
}

// what the solve step needs besides the sums
struct SolveArgs {
    int pass, o3d;
    float acc_err, eps;
    int max_iteration;
    float *pose_hist;
    float *trace_err;
    int trace_len;
    ShardView shard;              // sharded maps: rebuild the scan's tile list for the next pass at the new pose
    uint4 *active;
    uint32_t *n_active_next;
};

// warp 0 of a scan's block: sv[] holds the scan's totals; lane 0 solves, then the warp lists the tiles
// of the next search launch (sharded maps only: the owned columns move with the pose)
__device__ __forceinline__ void solve_and_list(ScanState &z, uint32_t scan, const double *sv, const SolveArgs &a)
{
    const int lane = threadIdx.x & 31;
    if (lane == 0) {
        float *hist = a.pose_hist + (size_t)scan * kCertHist * 16;
        float *trace = a.pass < a.trace_len ? a.trace_err + (size_t)scan * a.trace_len : nullptr;
        if (a.o3d) o3d_solve(z, sv, a.pass, a.max_iteration, hist, trace);
        else gn_solve(z, sv, a.pass, a.acc_err, a.eps, hist, trace);
    }
    __syncwarp();
    if (a.shard.enabled && !z.done) list_scan_tiles(z, scan, z.T, a.shard, a.active, a.n_active_next);
}

// the sums arrive from outside (map sharding: all-reduced across ranks by NCCL or the caller's hook)
__global__ void __launch_bounds__(32) solve_kernel(ScanState *states, const double *__restrict__ sums, SolveArgs a)
{
    __shared__ double sv[kAccum];
    ScanState &z = states[blockIdx.x];
    if (z.done) return;
    sv[threadIdx.x] = sums[(size_t)blockIdx.x * kAccum + threadIdx.x];
    __syncwarp();
    solve_and_list(z, blockIdx.x, sv, a);
}

// single GPU: ordered sum of the scan's partial rows and the solve in one launch
__global__ void __launch_bounds__(256)
    rowsum_solve_kernel(ScanState *states, const double *__restrict__ partials, double *__restrict__ sums, SolveArgs a)
{
    __shared__ double sw[8][kAccum];
    __shared__ double sv[kAccum];
    ScanState &z = states[blockIdx.x];
    if (z.done) return;
    const double tot = scan_rowsum(z, partials, a.pass, sw);
    if (threadIdx.x >= 32) return;
    sums[(size_t)blockIdx.x * kAccum + threadIdx.x] = tot;
    sv[threadIdx.x] = tot;
    __syncwarp();
    solve_and_list(z, blockIdx.x, sv, a);
}

// =========================================================================================
// REFERENCE mode (icp_point_to_point.cpp:185-254)
// =========================================================================================

// sourceTargetCorrespondences (cpp:57-84).  first != 0: P = T_init * src (cpp:191-192) and
// every row is searched; later searches only touch rows that still have a correspondence
// (the source shrinks, cpp:77-83).  Rows are not physically compacted: corr < 0 marks a row
// as dropped, and dropped rows add exact zeros to the ordered sums.
__global__ void __launch_bounds__(kTile)
    ref_search_kernel(MapView map, const float4 *__restrict__ src, float4 *__restrict__ P, float4 *__restrict__ Q,
                      int32_t *__restrict__ corr, uint32_t *__restrict__ pos_of, const uint32_t *__restrict__ tile_scan,
                      const ScanState *__restrict__ states, float limit, int first)
{
    __shared__ float sT[16];
    const uint32_t scan = tile_scan[blockIdx.x];
    const ScanState &z = states[scan];
    if (z.done || (!first && !z.need_search)) return;
    const uint32_t row0 = (blockIdx.x - z.tile_begin) * kTile;
    if (row0 >= z.n_pts) return;
    if (first) {
        if (threadIdx.x < 16) sT[threadIdx.x] = z.T_init[threadIdx.x];
        __syncthreads();
    }
    const uint32_t row = row0 + threadIdx.x;
    const size_t slot = (size_t)z.pt_begin + row;
    bool active = row < z.n_pts;
    float3 p = make_float3(0.f, 0.f, 0.f);
    uint32_t seed = kNoPos;
    if (active) {
        if (first) {
            const float4 s4 = src[slot];
            p = transform_point(sT, s4.x, s4.y, s4.z);
            P[slot] = make_float4(p.x, p.y, p.z, 1.f);
        } else if (corr[slot] < 0) {
            active = false;  // dropped rows stay dropped
        } else {
            const float4 p4 = P[slot];
            p = make_float3(p4.x, p4.y, p4.z);
            seed = pos_of[slot];  // the neighbour of the previous search: a bound to start the walk from
        }
    }
    if (!active) return;
    // (a seed only tightens the bound the walk starts with; the result is that of the unseeded walk)
    NNHit h;
    h.d2 = limit; h.idx = -1; h.pos = 0;
    if (limit > 0.f && map.n_pts > 0 && isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        NNBest<false> B;
        if (nn_walk_near<false>(map, p.x, p.y, p.z, limit, 0.f, B, seed)) {
            B.skip = kNoPos;
            nn_walk_far<false>(map, p.x, p.y, p.z, B);
        }
        if ((uint32_t)(B.key >> 32) < __float_as_uint(limit)) {
            h.d2 = B.bd();
            h.idx = (int)(uint32_t)(B.key & 0xFFFFFFFFull);
            h.pos = B.pos;
        }
    }
    corr[slot] = h.idx;
    if (h.idx >= 0) {
        float4 q = __ldg(&map.pts[h.pos]);
        q.w = 1.f;
        Q[slot] = q;
        pos_of[slot] = h.pos;
    }
}

// applyTransformation(T_step, P) (cpp:230)
__global__ void __launch_bounds__(kTile)
    ref_step_kernel(float4 *__restrict__ P, const int32_t *__restrict__ corr, const uint32_t *__restrict__ tile_scan,
                    ScanState *states)
{
    __shared__ float sT[16];
    const uint32_t scan = tile_scan[blockIdx.x];
    const ScanState &z = states[scan];
    if (z.done || !z.have_step) return;
    const uint32_t row0 = (blockIdx.x - z.tile_begin) * kTile;
    if (row0 >= z.n_pts) return;
    if (threadIdx.x < 16) sT[threadIdx.x] = z.T_step[threadIdx.x];
    __syncthreads();
    const uint32_t row = row0 + threadIdx.x;
    if (row >= z.n_pts) return;
    const size_t slot = (size_t)z.pt_begin + row;
    if (corr[slot] < 0) return;
    const float4 p4 = P[slot];
    const float3 p = transform_point(sT, p4.x, p4.y, p4.z);
    P[slot] = make_float4(p.x, p.y, p.z, 1.f);
}

// ---- ordered float chains (STRICT) ------------------------------------------------------------
// Two shapes of the one-block-per-scan kernel (RT threads, CT rows staged per buffer).  A block is bound by the
// latency of its chain warp, so a batch's throughput is the number of RESIDENT blocks:
//   512 threads, 1024 rows (74 KB): warp 0 walks the chains, 15 warps stage the next tile's values -- the
//       staging of a tile hides behind its chain even for one scan alone; 2 blocks per SM (62 registers)
//   256 threads, 704 rows (51 KB): 4 blocks per SM, for batches of more scans than 2 x SMs (the offline sequence)
constexpr int kRefThreads = 512, kChainTile = 1024;
constexpr int kRefThreadsNarrow = 256, kChainTileNarrow = 704;
constexpr int kChainMax = 9;

template <int CT>
struct ChainBuf {
    float v[2][kChainMax][CT + 4];  // 16-byte aligned columns (the chain warp reads them with 128-bit loads)
};

// per-row values of the first pass: |p - q| (cpp:166), p (cpp:119), q (cpp:120)
__device__ __forceinline__ void rowvals_pass1(const float4 &p, const float4 &q, bool alive, float *o)
{
    if (!alive) {
#pragma unroll
        for (int i = 0; i < 7; ++i) o[i] = 0.f;
        return;
    }
    const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
    const float yz = dy * dy + dz * dz;  // Eigen fixed-size-3 reduction: a0 + (a1 + a2)
    o[0] = sqrtf(dx * dx + yz);
    o[1] = p.x; o[2] = p.y; o[3] = p.z;
    o[4] = q.x; o[5] = q.y; o[6] = q.z;
}

// per-row values of the second pass: (p - cs)(q - ct)^T, H(r,c) at index c*3 + r (cpp:126-134)
__device__ __forceinline__ void rowvals_pass2(const float4 &p, const float4 &q, bool alive, const float *cs,
                                              const float *ct, float *o)
{
    if (!alive) {
#pragma unroll
        for (int i = 0; i < 9; ++i) o[i] = 0.f;
        return;
    }
    const float a[3] = {p.x - cs[0], p.y - cs[1], p.z - cs[2]};
    const float b[3] = {q.x - ct[0], q.y - ct[1], q.z - ct[2]};
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) o[c * 3 + r] = a[r] * b[c];
}

// Sequential float sums of NCH per-row value streams over the rows of one scan, in row order.
// Warp 0 walks the chains (lane = chain) while warps 1.. compute the next tile's values.
template <int NCH, int PASS, int kRefThreads, int kChainTile>
__device__ void strict_chains(const float4 *P, const float4 *Q, const int32_t *corr, size_t base, uint32_t n,
                              const float *cs, const float *ct, ChainBuf<kChainTile> &buf, float *out /*[NCH] in smem*/)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_tiles = (n + kChainTile - 1) / kChainTile;
    // staging by warps 1..: every thread issues the loads of ALL its rows of the tile first (three
    // independent loads per row; P / Q of a dropped row are stale or unset and never used), then
    // computes and stores -- one memory latency per tile instead of one per row
    // the staging warps are the ones that do NOT share warp 0's scheduler (warp id mod 4 != 0): nothing else is
    // issued on the sub-partition the dependent adds run on
    constexpr int kFillThreads = kRefThreads - kRefThreads / 4;
    constexpr int kFillRows = (kChainTile + kFillThreads - 1) / kFillThreads;
    const bool filler = (warp & 3) != 0;
    const uint32_t fidx = (uint32_t)(warp - 1 - (warp >> 2)) * 32u + (uint32_t)lane;  // 0 .. kFillThreads-1 over the staging warps
    auto fill = [&](uint32_t t, int which) {
        int32_t c[kFillRows];
        float4 p[kFillRows], q[kFillRows];
#pragma unroll
        for (int k = 0; k < kFillRows; ++k) {
            const uint32_t j = fidx + (uint32_t)k * kFillThreads, row = t * kChainTile + j;
            c[k] = -1;
            p[k] = q[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < (uint32_t)kChainTile && row < n) {
                c[k] = corr[base + row];
                p[k] = P[base + row];
                q[k] = Q[base + row];
            }
        }
#pragma unroll
        for (int k = 0; k < kFillRows; ++k) {
            const uint32_t j = fidx + (uint32_t)k * kFillThreads;
            if (j >= (uint32_t)kChainTile) continue;
            float o[kChainMax];
            const bool alive = c[k] >= 0;  // rows past the end of the scan: c = -1 -> exact zeros
            if (PASS == 1) rowvals_pass1(p[k], q[k], alive, o);
            else rowvals_pass2(p[k], q[k], alive, cs, ct, o);
#pragma unroll
            for (int i = 0; i < NCH; ++i) buf.v[which][i][j] = o[i];
        }
    };
    float acc = 0.f;
    if (filler && n_tiles > 0) fill(0, 0);
    __syncthreads();
    for (uint32_t t = 0; t < n_tiles; ++t) {
        if (warp == 0) {
            if (lane < NCH) {
                // the adds are one dependent chain (4 cycles each); keep the next 16 values in
                // registers so the shared-memory loads never sit on the chain
                // two register sets in alternation: the four 128-bit loads of one are issued before the
                // 16 dependent adds of the other, so their latency hides behind the chain (measured
                // 5.3 cycles per row against 4.1 for the adds alone, profiles/exp/mb/chain.cu)
                const float4 *col = reinterpret_cast<const float4 *>(buf.v[t & 1][lane]);
                float4 a[4], b[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = col[u];
                for (int j = 0; j < kChainTile / 4; j += 8) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) b[u] = col[j + 4 + u];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        acc = __fadd_rn(acc, a[u].x); acc = __fadd_rn(acc, a[u].y);
                        acc = __fadd_rn(acc, a[u].z); acc = __fadd_rn(acc, a[u].w);
                    }
                    if (j + 8 < kChainTile / 4) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) a[u] = col[j + 8 + u];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        acc = __fadd_rn(acc, b[u].x); acc = __fadd_rn(acc, b[u].y);
                        acc = __fadd_rn(acc, b[u].z); acc = __fadd_rn(acc, b[u].w);
                    }
                }
            }
        } else if (filler && t + 1 < n_tiles) {
            fill(t + 1, (t + 1) & 1);
        }
        __syncthreads();
    }
    if (warp == 0 && lane < NCH) out[lane] = acc;
    __syncthreads();
}

// ---- parallel double sums (FAST) ----------------------------------------------------------------
template <int NV, int PASS, int kRefThreads>
__device__ void fast_sums(const float4 *P, const float4 *Q, const int32_t *corr, size_t base, uint32_t n,
                          const float *cs, const float *ct, double *scratch /*[kRefThreads]*/, float *out)
{
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0;
    for (uint32_t row = threadIdx.x; row < n; row += kRefThreads) {
        if (corr[base + row] < 0) continue;
        const float4 p = P[base + row], q = Q[base + row];
        float o[kChainMax];
        if (PASS == 1) rowvals_pass1(p, q, true, o);
        else rowvals_pass2(p, q, true, cs, ct, o);
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] += (double)o[i];
    }
    for (int i = 0; i < NV; ++i) {
        scratch[threadIdx.x] = acc[i];
        __syncthreads();
        for (int s = kRefThreads / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) scratch[threadIdx.x] += scratch[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[i] = (float)scratch[0];
        __syncthreads();
    }
}

template <int kRefThreads>
__device__ uint32_t block_count_alive(const int32_t *corr, size_t base, uint32_t n, uint32_t *scratch)
{
    uint32_t c = 0;
    for (uint32_t row = threadIdx.x; row < n; row += kRefThreads) c += corr[base + row] >= 0 ? 1u : 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = c;
    __syncthreads();
    uint32_t tot = 0;
    for (int w = 0; w < kRefThreads / 32; ++w) tot += scratch[w];
    __syncthreads();
    return tot;
}

// The per-pass control of calculateAlignment, one block per scan.
//   phase 0: after the pre-loop search (cpp:195-205), then falls through to phase 1 of pass 0
//   phase 1: top of pass i  -- error, break test, re-search test, else Kabsch step (cpp:209-234)
//   phase 2: after a re-search -- Kabsch step on the new correspondences (cpp:226-234)
template <int kRefThreads, int kChainTile>
__global__ void __launch_bounds__(kRefThreads, kRefThreads == kRefThreadsNarrow ? 4 : 1)
    ref_reduce_kernel(ScanState *states, const float4 *__restrict__ P, const float4 *__restrict__ Q,
                      const int32_t *__restrict__ corr, int phase, int pass, int reduce, float acc_err, float eps,
                      float *trace_err, int32_t *trace_search, int trace_len)
{
    extern __shared__ __align__(16) unsigned char ref_dyn_smem[];  // 2 x 9 x (rows + 4) floats: above the static limit
    ChainBuf<kChainTile> &buf = *reinterpret_cast<ChainBuf<kChainTile> *>(ref_dyn_smem);
    __shared__ double dscratch[kRefThreads];
    __shared__ uint32_t uscratch[kRefThreads / 32];
    __shared__ float sums1[7], sums2[9], cs[3], ct[3];
    __shared__ int s_action;  // 0 stop, 1 kabsch
    ScanState &z = states[blockIdx.x];
    if (z.done) return;
    if (phase == 2 && !z.need_search) return;
    const size_t base = z.pt_begin;
    const uint32_t n = z.n_pts;
    uint32_t K = 0;
    if (phase != 2 && threadIdx.x == 0) z.have_step = 0;  // the previous pass's step has been applied
    if (phase != 1) K = block_count_alive<kRefThreads>(corr, base, n, uscratch);
    if (phase == 0) {
        if (threadIdx.x == 0) {
            z.n_searches = 1;
            z.k_last = (int)K;
            if (K < 10) {  // cpp:196-200
                z.aborted = 1;
                z.done = 1;
            }
        }
        if (K < 10) return;
        phase = 1;
    } else if (phase == 2) {
        if (threadIdx.x == 0) {
            z.n_searches += 1;
            z.k_last = (int)K;
            z.need_search = 0;
            if (pass < trace_len) trace_search[(size_t)blockIdx.x * trace_len + pass] = 1;
            if (K == 0) {  // contract: the reference divides by zero here (cpp:122-123)
                z.last_error = z.pend_error;
                z.done = 1;
            }
        }
        if (K == 0) return;
    }
    // centroid / error chains (phase 1 needs the error; both need the centroids)
    if (reduce == SSF_REDUCE_STRICT) strict_chains<7, 1, kRefThreads, kChainTile>(P, Q, corr, base, n, nullptr, nullptr, buf, sums1);
    else fast_sums<7, 1, kRefThreads>(P, Q, corr, base, n, nullptr, nullptr, dscratch, sums1);
    if (threadIdx.x == 0) {
        s_action = 1;
        const float kf = (float)z.k_last;
        if (phase == 1) {
            const float error = sums1[0] / kf;  // cpp:169
            if (pass < trace_len) trace_err[(size_t)blockIdx.x * trace_len + pass] = error;
            if (error < acc_err) {  // cpp:215-219
                z.last_error = error;
                z.done = 1;
                s_action = 0;
            } else if (fabsf(z.last_error - error) < eps) {  // cpp:221-224
                z.need_search = 1;
                z.pend_error = error;
                s_action = 0;
            } else {
                z.pend_error = error;
            }
        }
        for (int k = 0; k < 3; ++k) {  // cpp:122-123
            cs[k] = sums1[1 + k] / kf;
            ct[k] = sums1[4 + k] / kf;
        }
    }
    __syncthreads();
    if (!s_action) return;
    if (reduce == SSF_REDUCE_STRICT) strict_chains<9, 2, kRefThreads, kChainTile>(P, Q, corr, base, n, cs, ct, buf, sums2);
    else fast_sums<9, 2, kRefThreads>(P, Q, corr, base, n, cs, ct, dscratch, sums2);
    if (threadIdx.x == 0) {
        float H[9], T_step[16];
        for (int i = 0; i < 9; ++i) H[i] = sums2[i];
        kabsch_from_moments_f(cs, ct, H, T_step);  // cpp:137-158
        mat4_mul_f(T_step, z.T, z.T);              // cpp:228
        for (int i = 0; i < 16; ++i) z.T_step[i] = T_step[i];
        z.have_step = 1;
        z.last_error = z.pend_error;  // cpp:232
        z.iterations += 1;            // cpp:234
    }
}

// =========================================================================================
// standalone search (parity / benchmark entry)
// =========================================================================================
__global__ void __launch_bounds__(kThreads)
    nn_only_kernel(MapView map, const float4 *__restrict__ q, uint32_t n, float limit, int32_t *__restrict__ idx,
                   float *__restrict__ d2)
{
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const float4 p = q[i];
    const NNHit h = nn_query(map, p.x, p.y, p.z, limit);
    idx[i] = h.idx;
    d2[i] = h.idx >= 0 ? h.d2 : FLT_MAX;
}

int nn_search_device(const MapView &map, const float4 *queries, size_t n, float limit, int32_t *idx, float *d2,
                     cudaStream_t st)
{
    if (n == 0) return SSF_OK;
    nn_only_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, st>>>(map, queries, (uint32_t)n, limit, idx,
                                                                                  d2);
    SSF_LAUNCHED();
    g_queries.fetch_add(n, std::memory_order_relaxed);
    return SSF_OK;
}

// =========================================================================================
// host: the fixed launch sequence of one batch alignment
// =========================================================================================
int SearchTimer::begin(cudaStream_t st)
{
    if (!enabled) return SSF_OK;
    while (pool.size() < used + 2) {
        cudaEvent_t e;
        SSF_CUDA(cudaEventCreate(&e));
        pool.push_back(e);
    }
    SSF_CUDA(cudaEventRecord(pool[used], st));
    return SSF_OK;
}
int SearchTimer::end(cudaStream_t st)
{
    if (!enabled) return SSF_OK;
    SSF_CUDA(cudaEventRecord(pool[used + 1], st));
    used += 2;
    return SSF_OK;
}

// ---- in-kernel exchange of the per-scan rows across ranks (map sharding) ---------------------------
// Buffer of one rank: rows[2][world][max_scans][kAccum] doubles, then flags[2][world][max_scans][2] u64
// ((epoch, shape) per parity, source rank and scan).
__device__ __forceinline__ double *xch_rows(void *base, const XchView &x, int par, int r)
{
    return reinterpret_cast<double *>(base) + ((size_t)(par * x.world + r) * x.max_scans) * kAccum;
}
__device__ __forceinline__ unsigned long long *xch_flag(void *base, const XchView &x, int par, int r, uint32_t scan)
{
    return reinterpret_cast<unsigned long long *>(reinterpret_cast<double *>(base) +
                                                  (size_t)2 * x.world * x.max_scans * kAccum) +
           2 * ((size_t)(par * x.world + r) * x.max_scans + scan);
}

// One launch per iteration, one block per scan: ordered sum of the scan's partial rows, STORED into every
// rank's buffer (peer stores over NVLink) and published with a per-scan release flag; then the block
// waits for the same scan's rows of every other rank, adds them in rank order (bit-identical on all
// ranks) and solves.  The compute (row sum, solve) and the collective (all-gather of 32 doubles per scan
// + ordered reduction) are one kernel; no host call, no NCCL launch, and ranks synchronise per scan, not
// per batch.  All blocks of the launch are resident at once (n_scans <= a few hundred), and a block only
// waits for stores its peers issue BEFORE they wait themselves, so there is no cyclic wait.
// The epoch of the run's first pass lives in device memory (x.epoch): the launch is graph-capturable.
__global__ void __launch_bounds__(256)
    rowsum_xchg_solve_kernel(ScanState *states, const double *__restrict__ partials, double *__restrict__ sums, XchView x,
                             unsigned long long shape, SolveArgs a)
{
    __shared__ double sw[8][kAccum];
    __shared__ double sv[kAccum];
    ScanState &z = states[blockIdx.x];
    if (z.done) return;  // identical on every rank: the states are replicas
    const double tot = scan_rowsum(z, partials, a.pass, sw);
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const unsigned long long epoch = *x.epoch + (unsigned long long)a.pass;
    const int par = (int)(epoch & 1ull);
    for (int r = 0; r < x.world; ++r) xch_rows(x.peers[r], x, par, x.rank)[(size_t)blockIdx.x * kAccum + lane] = tot;
    __threadfence_system();
    __syncwarp();
    if (lane < x.world) {
        unsigned long long *f = xch_flag(x.peers[lane], x, par, x.rank, blockIdx.x);
        f[1] = shape;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
    }
    void *mine = x.peers[x.rank];
    bool late = false;
    if (lane < x.world) {
        // bounded wait: a peer that failed before its row-sum, or whose epochs drifted (different batch
        // size / iteration count / mode), must not leave this GPU spinning for ever
        const unsigned long long *f = xch_flag(mine, x, par, lane, blockIdx.x);
        unsigned long long v, t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= epoch) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            late = t1 - t0 > x.timeout_ns;
        } while (!late);
        // epochs drifted (v > epoch) or the peer ran another batch shape: its rows are not this pass's
        if (!late && (v != epoch || __ldcg(f + 1) != shape)) late = true;
    }
    if (__any_sync(0xffffffffu, late)) {  // give up on this scan: flagged in the result, ssf_batch_results -> SSF_ERR_COMM
        if (lane == 0) {
            z.comm_error = 1;
            z.done = 1;
        }
        return;
    }
    double all = 0.0;
    for (int r = 0; r < x.world; ++r) all += __ldcg(xch_rows(mine, x, par, r) + (size_t)blockIdx.x * kAccum + lane);
    sums[(size_t)blockIdx.x * kAccum + lane] = all;
    sv[lane] = all;
    __syncwarp();
    solve_and_list(z, blockIdx.x, sv, a);
}

// the run is over: the next one starts at a fresh epoch (one per pass; O3D runs one more than num_iterations)
__global__ void xch_advance_kernel(unsigned long long *epoch, unsigned long long by) { *epoch += by; }

// per-scan totals of the partial rows and the solve.  Map sharding: the totals are summed across ranks
// inside the kernel (peer exchange), by ncclAllReduce on the library's stream, or through the caller's
// hook; otherwise row sum and solve are one launch.
static int reduce_and_solve(const IcpConfig &cfg, BatchBuffers &b, int pass, int o3d, const ShardView &shard,
                            cudaStream_t st)
{
    const int n_pass = cfg.num_iterations + 1;
    SolveArgs a{pass, o3d, cfg.acc_err, cfg.eps, cfg.num_iterations, b.pose_hist.p, b.trace_err.p, b.trace_len,
                shard, b.active.p, b.counters.p + (pass + 1 < n_pass ? pass + 1 : n_pass)};
    if (cfg.xch.world > 0) {  // in-kernel exchange over peer memory
        const unsigned long long shape = ((unsigned long long)b.n_scans << 32) |
                                         ((unsigned long long)(cfg.num_iterations & 0xFFFFF) << 8) |
                                         (unsigned long long)(cfg.mode & 0xFF);
        rowsum_xchg_solve_kernel<<<(unsigned)b.n_scans, 256, 0, st>>>(b.state.p, b.partials.p, b.sums.p, cfg.xch, shape, a);
        SSF_LAUNCHED();
        return SSF_OK;
    }
    if (!cfg.allreduce && !cfg.nccl_allreduce) {
        rowsum_solve_kernel<<<(unsigned)b.n_scans, 256, 0, st>>>(b.state.p, b.partials.p, b.sums.p, a);
        SSF_LAUNCHED();
        return SSF_OK;
    }
    rowsum_kernel<<<(unsigned)b.n_scans, 256, 0, st>>>(b.state.p, b.partials.p, b.sums.p, pass);
    SSF_LAUNCHED();
    if (cfg.nccl_allreduce) {
        SSF_TRY(cfg.nccl_allreduce(cfg.nccl_user, b.sums.p, b.n_scans * kAccum, st));
    } else if (cfg.allreduce(cfg.allreduce_user, b.sums.p, b.n_scans * kAccum, (void *)st) != 0) {
        set_error("all-reduce hook failed");
        return SSF_ERR_COMM;
    }
    solve_kernel<<<(unsigned)b.n_scans, 32, 0, st>>>(b.state.p, b.sums.p, a);
    SSF_LAUNCHED();
    return SSF_OK;
}

#define TIMED_SEARCH(launch)                       \
    do {                                           \
        if (timer) SSF_TRY(timer->begin(st));      \
        launch;                                    \
        SSF_LAUNCHED();                            \
        if (timer) SSF_TRY(timer->end(st));        \
    } while (0)

// launch the fused search: small batches (a few scans) take one query per thread on 512-thread
// blocks -- four times the parallelism per tile, for latency --, large ones four queries per thread
template <int KIND>
static void launch_search(bool wide, unsigned grid, cudaStream_t st, const MapView &map, const BatchBuffers &b,
                          float limit, int use_cert, int pass, const uint32_t *n_active, uint32_t *fetch)
{
    if (wide)
        search_accum_kernel<KIND, kWideThreads><<<grid, kWideThreads, 0, st>>>(map, b.src.p, b.tile_scan.p, b.state.p, limit, b.corr.p,
                                                                 b.partials.p, b.cert.p, b.pose_hist.p, use_cert, pass,
                                                                 b.active.p, n_active, fetch, b.search_stats.p + 2 * pass);
    else
        search_accum_kernel<KIND, kThreads><<<grid, kThreads, 0, st>>>(map, b.src.p, b.tile_scan.p, b.state.p, limit,
                                                                       b.corr.p, b.partials.p, b.cert.p, b.pose_hist.p,
                                                                       use_cert, pass, b.active.p, n_active, fetch,
                                                                       b.search_stats.p + 2 * pass);
}

static int search_grid(const BatchBuffers &b, unsigned *grid, bool *wide, int *sms)
{
    // queried per call: contexts on different devices share this code (a process-wide cache would
    // pin the first device's count), and the attribute read costs well under a microsecond
    int n_sm = 0, dev = 0;
    SSF_CUDA(cudaGetDevice(&dev));
    SSF_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    *sms = n_sm;
    // persistent blocks; sized by the batch CAPACITY so that the launch shape (and a captured graph)
    // does not change with the number of points of an upload -- surplus blocks fetch once and exit
    size_t g = (size_t)n_sm * (size_t)SSF_MINB;
    // few tiles: latency matters.  Crossover measured with 32x1024 scans against the 5M-point map (profiles/exp/
    // exp_shape.py): 8 scans (496 tiles) 0.38 ms wide / 0.44 narrow, 16 scans (992 tiles) 0.58 / 0.54,
    // 64 scans 1.63 / 1.14
    *wide = b.max_tiles < 5u * (size_t)n_sm;
    if (const char *fw = getenv("SSF_SEARCH_WIDE")) *wide = atoi(fw) != 0;  // (experiments: force one shape)
    if (*wide) g = (size_t)n_sm * 2u;  // 512-thread blocks, two per SM
    if (g > b.max_tiles) g = b.max_tiles;
    *grid = (unsigned)(g ? g : 1);
    return SSF_OK;
}

// the launches of one batch alignment (everything after the voxel stage), in stream order
static int enqueue_batch(const MapView &map, const IcpConfig &cfg, BatchBuffers &b, const float *T_init, bool certs,
                         unsigned grid, bool wide, int n_sm, cudaStream_t st, SearchTimer *timer)
{
    SSF_TRY(init_states(b, T_init, st));
    const unsigned tiles = (unsigned)b.n_tiles, scans = (unsigned)b.n_scans;
    const float limit = cfg.max_corr;
    ScanState *S = b.state.p;
    if (tiles == 0) {
        // no source point at all: the first search finds 0 (< 10) correspondences -- the reference's
        // abort sentinel (cpp:196-200); the Open3D flow ends with nothing matched (rmse 0)
        no_points_kernel<<<(scans + 127) / 128, 128, 0, st>>>(S, scans, cfg.mode);
        SSF_LAUNCHED();
        results_kernel<<<(scans + 127) / 128, 128, 0, st>>>(S, b.results.p, scans, cfg.mode, cfg.acc_err);
        SSF_LAUNCHED();
        return SSF_OK;
    }
    const int n_pass = cfg.num_iterations + 1;  // search launches of the longest flow (O3D runs one more than GN)
    const bool sharded = map.own_lo != INT32_MIN || map.own_hi != INT32_MAX;
    ShardView shard;
    if (sharded) {
        shard.enabled = 1;
        shard.ox = map.shard_ox;
        shard.inv_h = map.shard_inv_h;
        shard.own_lo = map.own_lo;
        shard.own_hi = map.own_hi;
        shard.tile_box = b.tile_box.p;
    }
    // sharded: launch i works on the list built for it (counters[i]); else one list serves every launch
    auto n_active_of = [&](int i) { return b.counters.p + (sharded ? i : 0); };
    auto fetch_of = [&](int i) { return b.counters.p + n_pass + 1 + i; };
    if (cfg.mode != SSF_MODE_REFERENCE) {
        // (zeroed by a kernel: memsets and copies on the compute stream can queue behind the other
        // batch's H2D copy on a copy engine and stall the pipeline)
        const uint32_t n_cnt = 2u * (uint32_t)n_pass + 2u;
        zero_u32_kernel<<<(n_cnt + 255) / 256, 256, 0, st>>>(b.counters.p, n_cnt);
        SSF_LAUNCHED();
        zero_u32_kernel<<<(unsigned)((4 * n_pass + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<uint32_t *>(b.search_stats.p), 4u * (uint32_t)n_pass);
        SSF_LAUNCHED();
        if (sharded) {
            tile_box_kernel<<<tiles, 128, 0, st>>>(b.src.p, b.tile_scan.p, S, b.tile_box.p);
            SSF_LAUNCHED();
            reset_rows_kernel<<<(unsigned)((b.n_slots + 255) / 256), 256, 0, st>>>(b.corr.p, b.cert.p, (uint32_t)b.n_slots);
            SSF_LAUNCHED();
            reset_stamps_kernel<<<(tiles + 255) / 256, 256, 0, st>>>(b.partials.p, tiles);
            SSF_LAUNCHED();
        }
        active_tiles_kernel<<<(scans * 32 + 127) / 128, 128, 0, st>>>(S, scans, shard, b.active.p, b.counters.p);
        SSF_LAUNCHED();
    }
    if (cfg.mode == SSF_MODE_GN_P2P || cfg.mode == SSF_MODE_GN_P2PLANE) {
        if (cfg.mode == SSF_MODE_GN_P2PLANE && !map.pn) {
            set_error("SSF_MODE_GN_P2PLANE needs target normals (ssf_icp_set_target normals == NULL)");
            return SSF_ERR_STATE;
        }
        for (int i = 0; i < cfg.num_iterations; ++i) {
            if (cfg.mode == SSF_MODE_GN_P2PLANE)
                TIMED_SEARCH(launch_search<ACC_GN_P2PLANE>(wide, grid, st, map, b, limit, certs && i > 0, i, n_active_of(i), fetch_of(i)));
            else
                TIMED_SEARCH(launch_search<ACC_GN_P2P>(wide, grid, st, map, b, limit, certs && i > 0, i, n_active_of(i), fetch_of(i)));
            SSF_TRY(reduce_and_solve(cfg, b, i, 0, shard, st));
        }
    } else if (cfg.mode == SSF_MODE_O3D_P2P) {
        for (int i = 0; i <= cfg.num_iterations; ++i) {
            TIMED_SEARCH(launch_search<ACC_KABSCH>(wide, grid, st, map, b, limit, certs && i > 0, i, n_active_of(i), fetch_of(i)));
            SSF_TRY(reduce_and_solve(cfg, b, i, 1, shard, st));
        }
    } else if (cfg.mode == SSF_MODE_REFERENCE) {
        if (map.own_lo != INT32_MIN || map.own_hi != INT32_MAX) {
            set_error("SSF_MODE_REFERENCE does not run on a sharded map (its ordered float sums do not split "
                      "across ranks); use a GN or O3D mode");
            return SSF_ERR_STATE;
        }
        uint32_t *pos_of = reinterpret_cast<uint32_t *>(b.cert.p);  // the certificate array is free in this mode
        TIMED_SEARCH((ref_search_kernel<<<tiles, kTile, 0, st>>>(map, b.src.p, b.P.p, b.Q.p, b.corr.p, pos_of, b.tile_scan.p,
                                                                S, limit, 1)));
        // ordered chains are the same sums in either shape; the double tree of FAST depends on the thread count,
        // so only STRICT batches switch (a scan must give the same result alone and in a batch)
        const bool narrow = cfg.reduce == SSF_REDUCE_STRICT && scans > 2u * (unsigned)n_sm;
        auto reduce = [&](int phase, int pass) {
            if (narrow)
                ref_reduce_kernel<kRefThreadsNarrow, kChainTileNarrow><<<scans, kRefThreadsNarrow, sizeof(ChainBuf<kChainTileNarrow>), st>>>(
                    S, b.P.p, b.Q.p, b.corr.p, phase, pass, cfg.reduce, cfg.acc_err, cfg.eps, b.trace_err.p, b.trace_search.p, b.trace_len);
            else
                ref_reduce_kernel<kRefThreads, kChainTile><<<scans, kRefThreads, sizeof(ChainBuf<kChainTile>), st>>>(
                    S, b.P.p, b.Q.p, b.corr.p, phase, pass, cfg.reduce, cfg.acc_err, cfg.eps, b.trace_err.p, b.trace_search.p, b.trace_len);
        };
        for (int i = 0; i < cfg.num_iterations; ++i) {
            reduce(i == 0 ? 0 : 1, i);
            SSF_LAUNCHED();
            ref_search_kernel<<<tiles, kTile, 0, st>>>(map, b.src.p, b.P.p, b.Q.p, b.corr.p, pos_of, b.tile_scan.p, S, limit, 0);
            SSF_LAUNCHED();
            reduce(2, i);
            SSF_LAUNCHED();
            if (i + 1 < cfg.num_iterations) {
                ref_step_kernel<<<tiles, kTile, 0, st>>>(b.P.p, b.corr.p, b.tile_scan.p, S);
                SSF_LAUNCHED();
            }
        }
    } else {
        set_error("unknown mode %d", cfg.mode);
        return SSF_ERR_INVALID;
    }
    if (cfg.xch.world > 0 && cfg.mode != SSF_MODE_REFERENCE) {
        xch_advance_kernel<<<1, 1, 0, st>>>(cfg.xch.epoch, (unsigned long long)cfg.num_iterations + 2);
        SSF_LAUNCHED();
    }
    results_kernel<<<(scans + 127) / 128, 128, 0, st>>>(S, b.results.p, scans, cfg.mode, cfg.acc_err);
    SSF_LAUNCHED();
    return SSF_OK;
}

// The Gauss-Newton / Open3D-flow loop is a fixed launch sequence whose shape depends only on the
// batch capacity, the scan count and the parameters -- all data-dependent control lives in device
// state -- so it is captured once into a CUDA graph and replayed: one graph launch per alignment,
// no host round trip, no per-kernel launch latency.  REFERENCE mode launches per-tile grids, so its
// graph is re-captured when the upload's tile count changes.  Runs with the search timer or the caller's
// all-reduce hook are host-side actions per launch and are enqueued directly.
int run_batch(const MapView &map, const IcpConfig &cfg, BatchBuffers &b, const float *T_init, cudaStream_t st,
              SearchTimer *timer)
{
    if (b.n_scans == 0) return SSF_OK;
    // SSF_NO_CERT=1 turns the search certificates off (every query walks every iteration): the
    // results must not change by a single bit (tests/test_gpu_parity.py::test_certificates_change_nothing)
    const char *nc = getenv("SSF_NO_CERT");
    const bool certs = !(nc && atoi(nc) != 0);
    unsigned grid = 1;
    bool wide = false;
    int n_sm = 0;
    SSF_TRY(search_grid(b, &grid, &wide, &n_sm));
    if (cfg.mode != SSF_MODE_REFERENCE) {  // allocations happen here, never inside a capture
        SSF_TRY(b.active.reserve(b.max_tiles ? b.max_tiles : 1));
        SSF_TRY(b.counters.reserve(2 * ((size_t)cfg.num_iterations + 1) + 2));
        if (map.own_lo != INT32_MIN || map.own_hi != INT32_MAX) SSF_TRY(b.tile_box.reserve(2 * (b.max_tiles ? b.max_tiles : 1)));
        SSF_TRY(b.search_stats.reserve(2 * ((size_t)cfg.num_iterations + 1)));
        b.search_stats_len = cfg.num_iterations + (cfg.mode == SSF_MODE_O3D_P2P ? 1 : 0);
    } else {
        b.search_stats_len = 0;
    }
    if (cfg.mode == SSF_MODE_REFERENCE) {
        // function attributes are per device: set on every run (cheap) rather than once per process
        SSF_CUDA(cudaFuncSetAttribute(ref_reduce_kernel<kRefThreads, kChainTile>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(ChainBuf<kChainTile>)));
        SSF_CUDA(cudaFuncSetAttribute(ref_reduce_kernel<kRefThreadsNarrow, kChainTileNarrow>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(ChainBuf<kChainTileNarrow>)));
        g_queries.fetch_add(b.n_slots, std::memory_order_relaxed);
    }
    const char *ng = getenv("SSF_NO_GRAPH");
    // (the caller's all-reduce hook is a host action per iteration: not capturable; the in-kernel exchange and
    // ncclAllReduce are)
    const bool graphable = b.n_tiles > 0 && !(cfg.allreduce && !cfg.nccl_allreduce && cfg.xch.world == 0) &&
                           !(timer && timer->enabled) && !(ng && atoi(ng) != 0);
    if (!graphable) {
        if (cfg.mode != SSF_MODE_REFERENCE)
            g_queries.fetch_add((uint64_t)b.n_slots * (uint64_t)cfg.num_iterations, std::memory_order_relaxed);
        return enqueue_batch(map, cfg, b, T_init, certs, grid, wide, n_sm, st, timer);
    }

    std::vector<unsigned long long> key;
    auto put = [&](const void *p, size_t n) {
        const unsigned char *c = static_cast<const unsigned char *>(p);
        for (size_t i = 0; i < n; i += 8) {
            unsigned long long w = 0;
            memcpy(&w, c + i, n - i < 8 ? n - i : 8);
            key.push_back(w);
        }
    };
    put(&map, sizeof(map));
    put(&cfg.max_corr, sizeof(float)); put(&cfg.acc_err, sizeof(float)); put(&cfg.eps, sizeof(float));
    const unsigned long long scalars[] = {(unsigned long long)cfg.num_iterations, (unsigned long long)cfg.mode,
                                          (unsigned long long)certs, (unsigned long long)grid, (unsigned long long)b.n_scans,
                                          (unsigned long long)wide, (unsigned long long)b.trace_len, (unsigned long long)cfg.xch.world,
                                          (unsigned long long)cfg.xch.rank, (unsigned long long)b.n_tiles, (unsigned long long)cfg.reduce};
    put(scalars, sizeof(scalars));
    const void *ptrs[] = {b.src.p, b.corr.p, b.cert.p, b.pose_hist.p, b.tile_scan.p, b.active.p, b.counters.p,
                          b.partials.p, b.sums.p, b.state.p, b.results.p, b.trace_err.p, b.trace_search.p, T_init,
                          b.search_stats.p, b.tile_box.p, b.P.p, b.Q.p, (const void *)cfg.xch.peers, (const void *)cfg.xch.epoch,
                          (const void *)cfg.nccl_allreduce, cfg.nccl_user};
    put(ptrs, sizeof(ptrs));
    if (!b.graph_exec || key != b.graph_key) {
        if (b.graph_exec) {
            cudaGraphExecDestroy(b.graph_exec);
            b.graph_exec = nullptr;
        }
        cudaGraph_t graph = nullptr;
        SSF_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const uint64_t before = g_launches.load();
        const int rc = enqueue_batch(map, cfg, b, T_init, certs, grid, wide, n_sm, st, nullptr);
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        b.graph_kernels = g_launches.load() - before;
        g_launches.fetch_sub(b.graph_kernels);  // counted when the graph is launched
        if (rc != SSF_OK || ce != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            if (rc != SSF_OK) return rc;
            set_error("stream capture failed: %s", cudaGetErrorString(ce));
            return SSF_ERR_CUDA;
        }
        const cudaError_t ie = cudaGraphInstantiate(&b.graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {
            b.graph_exec = nullptr;
            set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
            return SSF_ERR_CUDA;
        }
        b.graph_key = key;
    }
    SSF_CUDA(cudaGraphLaunch(b.graph_exec, st));
    g_launches.fetch_add(b.graph_kernels, std::memory_order_relaxed);
    if (cfg.mode != SSF_MODE_REFERENCE)
        g_queries.fetch_add((uint64_t)b.n_slots * (uint64_t)cfg.num_iterations, std::memory_order_relaxed);
    return SSF_OK;
}

}  // namespace ssf
