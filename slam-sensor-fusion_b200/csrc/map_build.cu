// map_build.cu -- K1: GPU voxel-grid build of the target map.
//
// Replaces kdtree_.setInputCloud(target_cloud) at reference
// localization/src/icp_point_to_point.cpp:54 (a FLANN KD-tree build on the CPU).
//
//   bbox -> directory key per point -> stable radix sort (key, original index) -> gather the
//   cloud into key order (float4, w = original index) -> unique cells + start offsets ->
//   directory fill (occupancy mask + first cell id per 32-cell block of a row).
//
// The cell edge is chosen from the data: a first pass at a volumetric guess measures the
// points per occupied cell, and the edge is shrunk (surface-like clouds) until a cell holds a
// few points, within a memory budget for the dense directory.  The rejection threshold does
// not enter: the search (nn_device.cuh) is exact for any radius at any cell edge.
//
// One-time cost per map (or per re-crop); algorithmic bytes 2*16*M + 8*n_dir per pass.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "map_index.cuh"

namespace ssf {

__global__ void __launch_bounds__(256)
    map_keys_kernel(const float4 *__restrict__ raw, uint32_t n, float ox, float oy, float oz, float inv_h, int nx, int ny,
                    int nz, int nbx, int nty, unsigned long long sentinel, unsigned long long *__restrict__ keys,
                    uint32_t *__restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = raw[i];
    unsigned long long k = sentinel;
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        int cx = cell_coord(p.x, ox, inv_h, nx), cy = cell_coord(p.y, oy, inv_h, ny), cz = cell_coord(p.z, oz, inv_h, nz);
        cx = min(max(cx, 0), nx - 1);
        cy = min(max(cy, 0), ny - 1);
        cz = min(max(cz, 0), nz - 1);
        k = ((unsigned long long)dir_index(nbx, nty, cx >> 5, cy, cz) << 5) | (unsigned long long)(cx & 31);
    }
    keys[i] = k;
    vals[i] = i;
}

__global__ void __launch_bounds__(256)
    map_flags_kernel(const unsigned long long *__restrict__ keys, uint32_t n_finite, uint32_t *__restrict__ flags)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_finite) return;
    flags[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
    map_gather_kernel(const float4 *__restrict__ raw, const float4 *__restrict__ raw_nrm,
                      const int32_t *__restrict__ global_index, const uint32_t *__restrict__ vals, uint32_t n_finite,
                      float4 *__restrict__ pts, float4 *__restrict__ nrm)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_finite) return;
    const uint32_t src = vals[j];
    float4 p = raw[src];
    p.w = __int_as_float(global_index ? global_index[src] : (int)src);
    pts[j] = p;
    if (raw_nrm) {  // (point, normal) side by side: one 32-byte record per point for the residual gathers
        nrm[2 * (size_t)j] = p;
        nrm[2 * (size_t)j + 1] = raw_nrm[src];
    }
}

// cell c (= exclusive scan of the flags at its first point) starts at point j; the first cell of
// a directory block records its id there, every cell sets its bit in the block's mask
__global__ void __launch_bounds__(256)
    map_cells_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ flags,
                     const uint32_t *__restrict__ scan, uint32_t n_finite, uint32_t *__restrict__ cell_start,
                     uint32_t n_cells, uint2 *__restrict__ dir)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) cell_start[n_cells] = n_finite;
    if (j >= n_finite) return;
    if (flags[j]) {
        const uint32_t c = scan[j];
        cell_start[c] = j;
        const unsigned long long k = keys[j];
        const uint32_t e = (uint32_t)(k >> 5);
        atomicOr(&dir[e].x, 1u << (uint32_t)(k & 31));
        if (j == 0 || (keys[j - 1] >> 5) != (k >> 5)) dir[e].y = c;
    }
}

static int bit_width_u64(unsigned long long v)
{
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

struct GridDims {
    int nx, ny, nz, nbx, nty, ntz;
    unsigned long long n_dir;
};

// same float expressions as cell_coord() on the device (no FMA possible: sub then mul)
static bool grid_dims(const float *bb, const float *org, float cell, GridDims &g)
{
    const float inv_h = 1.0f / cell;
    const float ux = (bb[3] - org[0]) * inv_h, uy = (bb[4] - org[1]) * inv_h, uz = (bb[5] - org[2]) * inv_h;
    if (!(ux < 1048000.f) || !(uy < 1048000.f) || !(uz < 1048000.f)) return false;
    g.nx = (int)floorf(ux) + 1; g.ny = (int)floorf(uy) + 1; g.nz = (int)floorf(uz) + 1;
    g.nbx = (g.nx + 31) / 32; g.nty = (g.ny + 3) / 4; g.ntz = (g.nz + 3) / 4;
    g.n_dir = (unsigned long long)g.ntz * g.nty * g.nbx * 16ull;
    return true;
}

// keys -> sort -> flags -> scan; leaves the sorted (key, original index) pairs in m.keys/m.vals
static int index_pass(MapIndex &m, const GridDims &g, float cell, const float *org, uint32_t n_finite, Scratch &s,
                      cudaStream_t st, uint32_t *cnt_dev, uint32_t *n_cells_out)
{
    const size_t n = m.n_raw;
    const unsigned long long sentinel = g.n_dir << 5;
    const int key_bits = bit_width_u64(sentinel);
    const unsigned blocks_n = (unsigned)((n + 255) / 256);
    map_keys_kernel<<<blocks_n, 256, 0, st>>>(m.raw.p, (uint32_t)n, org[0], org[1], org[2], 1.0f / cell, g.nx, g.ny, g.nz,
                                              g.nbx, g.nty, sentinel, m.keys.p, m.vals.p);
    SSF_LAUNCHED();
    SSF_TRY(radix_sort_pairs_u64(m.keys.p, m.vals.p, n, key_bits, s, st));
    const unsigned blocks_f = (n_finite + 255) / 256;
    map_flags_kernel<<<blocks_f, 256, 0, st>>>(m.keys.p, n_finite, m.flags.p);
    SSF_LAUNCHED();
    SSF_TRY(exclusive_scan_u32(m.flags.p, m.cell_id.p, n_finite, cnt_dev + 1, s, st));
    SSF_CUDA(cudaMemcpyAsync(n_cells_out, cnt_dev + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SSF_CUDA(cudaStreamSynchronize(st));
    return SSF_OK;
}

int build_map_index(MapIndex &m, float cell_size, Scratch &s, cudaStream_t st)
{
    const size_t n = m.n_raw;
    if (n >= (size_t)1 << 31) {
        set_error("target cloud too large for 31-bit indices (%zu)", n);
        return SSF_ERR_INVALID;
    }
    if (cell_size != cell_size || std::isinf(cell_size)) {
        set_error("invalid cell size %g", cell_size);
        return SSF_ERR_INVALID;
    }
    m.view = MapView{};
    m.reach_limit = -1.f;
    m.cell_size = 0.f;
    m.n_cells = m.n_dir = 0;
    m.build_passes = 0;
    SSF_TRY(m.small.reserve(16));
    float *bbox_dev = m.small.p;
    uint32_t *cnt_dev = reinterpret_cast<uint32_t *>(m.small.p + 8);  // [0] n_finite, [1] n_cells
    SSF_TRY(bbox_finite(m.raw.p, n, bbox_dev, cnt_dev, st));
    float hb[6];
    uint32_t n_finite = 0;
    SSF_CUDA(cudaMemcpyAsync(hb, bbox_dev, sizeof(hb), cudaMemcpyDeviceToHost, st));
    SSF_CUDA(cudaMemcpyAsync(&n_finite, cnt_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SSF_CUDA(cudaStreamSynchronize(st));
    memcpy(m.bbox, hb, sizeof(hb));

    MapView v{};
    v.n_pts = n_finite;
    v.own_lo = m.own_lo;
    v.own_hi = m.own_hi;
    v.shard_ox = m.sharded ? m.shard_origin[0] : 0.f;
    v.shard_inv_h = m.sharded ? 1.0f / m.shard_cell : 1.f;
    if (n_finite == 0) {  // empty map: every search misses (nn_query returns before touching the index)
        v.ox = v.oy = v.oz = 0.f;
        v.inv_h = 1.f;
        v.hq = 0.999999f;
        v.nx = v.ny = v.nz = 1;
        v.nbx = v.nty = 1;
        SSF_TRY(m.dir.reserve(16));
        SSF_TRY(m.cell_start.reserve(1));
        SSF_TRY(m.pts.reserve(1));
        SSF_CUDA(cudaMemsetAsync(m.dir.p, 0, 16 * sizeof(uint2), st));
        v.dir = m.dir.p; v.cell_start = m.cell_start.p; v.pts = m.pts.p; v.pn = nullptr;
        m.n_dir = 16;
        m.cell_size = 1.f;
        m.view = v;
        return SSF_OK;
    }
    const float org[3] = {hb[0], hb[1], hb[2]};
    const float ext[3] = {hb[3] - hb[0], hb[4] - hb[1], hb[5] - hb[2]};
    const float ext_max = std::max(ext[0], std::max(ext[1], ext[2]));
    // directory budget: at most 64 bytes of directory per point (>= 8 MB), and < 2 GB
    const unsigned long long budget =
        std::min<unsigned long long>(1ull << 28, std::max<unsigned long long>(1ull << 20, 8ull * n_finite));
    // smallest usable edge: keeps every cell coordinate below 2^20 (float cell arithmetic) and
    // the edge itself far from the float spacing of the coordinates
    const float abs_max = std::max(std::max(fabsf(hb[0]), fabsf(hb[3])),
                                   std::max(std::max(fabsf(hb[1]), fabsf(hb[4])), std::max(fabsf(hb[2]), fabsf(hb[5]))));
    const float h_floor = std::max(std::max(ext_max / 1.0e6f, abs_max * 1.0e-5f), 1.0e-30f);
    auto fit = [&](float h, GridDims &g) -> float {  // grow h until the grid is addressable and within budget
        if (!(h > h_floor)) h = h_floor;
        for (int it = 0; it < 400; ++it) {
            if (grid_dims(hb, org, h, g) && g.n_dir <= budget) return h;
            h *= 1.1f;
        }
        return -1.f;
    };

    SSF_TRY(m.keys.reserve(n));
    SSF_TRY(m.vals.reserve(n));
    SSF_TRY(m.flags.reserve(n_finite));
    SSF_TRY(m.cell_id.reserve(n_finite));

    GridDims g{};
    float h;
    uint32_t n_cells = 0;
    const char *env = getenv("SSF_CELL_SIZE");
    if (env && atof(env) > 0.0) cell_size = (float)atof(env);
    if (cell_size > 0.f) {
        h = fit(cell_size, g);
        if (h < 0.f) { set_error("map extent does not fit the directory at any cell size"); return SSF_ERR_INVALID; }
        SSF_TRY(index_pass(m, g, h, org, n_finite, s, st, cnt_dev, &n_cells));
        m.build_passes = 1;
    } else {
        // volumetric guess (4 points per cell if the cloud filled its box), then shrink while the
        // occupied cells are crowded: for a surface-like cloud points/cell scales with h^2
        const float eps = std::max(ext_max * 1.0e-3f, 1.0e-20f);
        const double vol = (double)std::max(ext[0], eps) * std::max(ext[1], eps) * std::max(ext[2], eps);
        h = ext_max > 0.f ? (float)cbrt(4.0 * vol / (double)n_finite) : 1.0f;
        const float target = getenv("SSF_CELL_POINTS") ? (float)atof(getenv("SSF_CELL_POINTS")) : 4.5f;
        for (int pass = 0; pass < 4; ++pass) {
            h = fit(h, g);
            if (h < 0.f) { set_error("map extent does not fit the directory at any cell size"); return SSF_ERR_INVALID; }
            SSF_TRY(index_pass(m, g, h, org, n_finite, s, st, cnt_dev, &n_cells));
            m.build_passes = pass + 1;
            const float rho = (float)n_finite / (float)(n_cells ? n_cells : 1);
            // h and g must stay the edge / grid the index was built with: only move on to a new
            // edge when another index_pass follows
            if (rho <= 2.0f * target || ext_max == 0.f || pass + 1 == 4) break;
            float h_new = h * sqrtf(target / rho);
            if (h_new < h / 16.f) h_new = h / 16.f;
            GridDims g2{};
            h_new = fit(h_new, g2);
            if (h_new < 0.f || h_new > 0.9f * h) break;
            h = h_new;
        }
    }
    v.ox = org[0]; v.oy = org[1]; v.oz = org[2];
    v.inv_h = 1.0f / h;
    v.hq = nextafterf((1.0f / v.inv_h) * 0.999999f, 0.f);
    v.nx = g.nx; v.ny = g.ny; v.nz = g.nz;
    v.nbx = g.nbx; v.nty = g.nty;
    for (int k = 0; k < 3; ++k) { v.bmin[k] = hb[k]; v.bmax[k] = hb[3 + k]; }
    v.cert_mu = (getenv("SSF_CERT_MU") ? (float)atof(getenv("SSF_CERT_MU")) : 0.1f) * h;
    v.cert_step = (getenv("SSF_CERT_STEP") ? (float)atof(getenv("SSF_CERT_STEP")) : 6.0f) * v.cert_mu;

    SSF_TRY(m.pts.reserve(n_finite));
    if (m.has_normals) SSF_TRY(m.nrm.reserve(2 * (size_t)n_finite));
    SSF_TRY(m.cell_start.reserve((size_t)n_cells + 1));
    SSF_TRY(m.dir.reserve((size_t)g.n_dir));
    SSF_CUDA(cudaMemsetAsync(m.dir.p, 0, (size_t)g.n_dir * sizeof(uint2), st));
    const unsigned blocks_f = (n_finite + 255) / 256;
    map_gather_kernel<<<blocks_f, 256, 0, st>>>(m.raw.p, m.has_normals ? m.raw_nrm.p : nullptr,
                                                m.has_global_index ? m.global_index.p : nullptr, m.vals.p, n_finite,
                                                m.pts.p, m.has_normals ? m.nrm.p : nullptr);
    SSF_LAUNCHED();
    map_cells_kernel<<<blocks_f, 256, 0, st>>>(m.keys.p, m.flags.p, m.cell_id.p, n_finite, m.cell_start.p, n_cells,
                                               m.dir.p);
    SSF_LAUNCHED();
    v.pts = m.pts.p;
    v.pn = m.has_normals ? m.nrm.p : nullptr;
    v.dir = m.dir.p;
    v.cell_start = m.cell_start.p;
    v.n_cells = n_cells;
    m.view = v;
    m.n_cells = n_cells;
    m.n_dir = (uint32_t)g.n_dir;
    m.cell_size = h;
    return SSF_OK;
}

// ---- reach mask ------------------------------------------------------------------------------------
// bit (cx & 31) of out[dir_index(cx >> 5, cy, cz)] is set iff some occupied cell c' can hold a point
// within `reach` of a position binned into cell c.  Lower bound of that distance along one axis for
// a cell offset D: (|D| - 1 - slack) cells of edge hq (slack covers the rounding of the binning
// expression for both points, hq under-estimates the edge); the three axes add in squares.  A set
// bit promises nothing (the walk runs as usual); a clear bit is a proof of emptiness.
__global__ void __launch_bounds__(256)
    reach_mask_kernel(const uint2 *__restrict__ dir, uint32_t n_dir, int ny, int nz, int nbx, int nty, int K, float slack,
                      float hq, float reach2, uint32_t *__restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_dir) return;
    const uint32_t t = i >> 4;
    const int bx = (int)(t % (uint32_t)nbx);
    const int cy = (int)((t / (uint32_t)nbx) % (uint32_t)nty) * 4 + (int)(i & 3u);
    const int cz = (int)(t / ((uint32_t)nbx * (uint32_t)nty)) * 4 + (int)((i >> 2) & 3u);
    uint32_t res = 0;
    if (cy < ny && cz < nz) {
        for (int dz = -K; dz <= K; ++dz) {
            const int rz = cz + dz;
            if (rz < 0 || rz >= nz) continue;
            const float gz = fmaxf((float)abs(dz) - 1.0f - slack, 0.f) * hq;
            for (int dy = -K; dy <= K; ++dy) {
                const int ry = cy + dy;
                if (ry < 0 || ry >= ny) continue;
                const float gy = fmaxf((float)abs(dy) - 1.0f - slack, 0.f) * hq;
                const float rem = __fmul_ru(__fsub_ru(reach2, __fadd_rd(__fmul_rd(gy, gy), __fmul_rd(gz, gz))), 1.00001f);
                if (!(rem > 0.f)) continue;
                const int dxmax = min((int)floorf(1.0f + slack + __fdiv_ru(__fsqrt_ru(rem), hq) + 1e-3f), 31);
                const uint32_t cur = dir[dir_index(nbx, nty, bx, ry, rz)].x;
                const uint32_t prev = bx > 0 ? dir[dir_index(nbx, nty, bx - 1, ry, rz)].x : 0u;
                const uint32_t next = bx + 1 < nbx ? dir[dir_index(nbx, nty, bx + 1, ry, rz)].x : 0u;
                if (!(cur | prev | next)) continue;
                const unsigned long long lo = ((unsigned long long)cur << 32) | prev, hi = ((unsigned long long)next << 32) | cur;
                res |= cur;
                for (int sft = 1; sft <= dxmax; ++sft) res |= (uint32_t)(lo >> (32 - sft)) | (uint32_t)(hi >> sft);
            }
        }
    }
    out[i] = res;
}

int ensure_reach_mask(MapIndex &m, float limit, cudaStream_t st)
{
    const char *off = getenv("SSF_NO_REACH");
    if (m.view.n_pts == 0 || !(limit > 0.f) || !std::isfinite(limit) || m.n_dir == 0 || (off && atoi(off) != 0)) {
        m.view.reach = nullptr;
        m.reach_limit = -1.f;
        return SSF_OK;
    }
    if (m.reach_limit == limit) return SSF_OK;
    // what a walk prunes with: (sqrt(limit) + mu)^2 in certificate mode, limit otherwise
    const double reach = (sqrt((double)limit) * (1.0 + 2e-6) + (double)m.view.cert_mu) * (1.0 + 1e-5);
    const double slack = 8e-7 * (double)(std::max(m.view.nx, std::max(m.view.ny, m.view.nz)) + 4);
    const int K = (int)ceil(1.0 + slack + reach / (double)m.view.hq);
    m.reach_limit = limit;
    if (K > 6) {  // (2K+1)^2 rows per cell: the mask would clear too few bits to pay for itself
        m.view.reach = nullptr;
        return SSF_OK;
    }
    SSF_TRY(m.reach.reserve(m.n_dir));
    reach_mask_kernel<<<(m.n_dir + 255) / 256, 256, 0, st>>>(m.dir.p, m.n_dir, m.view.ny, m.view.nz, m.view.nbx, m.view.nty, K,
                                                              (float)slack, m.view.hq, (float)(reach * reach), m.reach.p);
    SSF_LAUNCHED();
    m.view.reach = m.reach.p;
    m.view.reach2 = nextafterf((float)(reach * reach * (1.0 - 1e-5)), 0.f);
    return SSF_OK;
}

}  // namespace ssf
