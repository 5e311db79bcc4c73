// map_build.cu -- K1: GPU voxel-hash build of the target map.
//
// Replaces kdtree_.setInputCloud(target_cloud) at reference
// localization/src/icp_point_to_point.cpp:54 (a FLANN KD-tree build on the CPU).
//
//   bbox -> cell key per point -> stable radix sort (key, original index) -> gather the
//   cloud into key order (float4, w = original index) -> unique cells + start offsets ->
//   open-addressing hash of every cell that is occupied or x-adjacent to an occupied cell,
//   value = offsets of the three cells {cx-1, cx, cx+1} (contiguous in the sorted cloud).
//
// One-time cost per map (or per re-crop); algorithmic bytes 2*16*M + 8*n_cells.
#include <cmath>

#include "map_index.cuh"

namespace ssf {

__global__ void __launch_bounds__(256) map_keys_kernel(const float4 *__restrict__ raw, uint32_t n, float ox, float oy,
                                                       float oz, float inv_h, int nx, int ny, int nz,
                                                       unsigned long long sentinel,
                                                       unsigned long long *__restrict__ keys,
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
        k = cell_key(cx, cy, cz, nx);
    }
    keys[i] = k;
    vals[i] = i;
}

__global__ void __launch_bounds__(256)
    map_gather_kernel(const float4 *__restrict__ raw, const float4 *__restrict__ raw_nrm,
                      const int32_t *__restrict__ global_index, const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t n_finite,
                      float4 *__restrict__ pts, float4 *__restrict__ nrm, uint32_t *__restrict__ flags)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_finite) return;
    const uint32_t src = vals[j];
    float4 p = raw[src];
    p.w = __int_as_float(global_index ? global_index[src] : (int)src);
    pts[j] = p;
    if (raw_nrm) nrm[j] = raw_nrm[src];
    flags[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1u : 0u;
}

// cell_id[j] = exclusive scan of flags (+flag - 1 = id of the cell point j belongs to)
__global__ void __launch_bounds__(256)
    map_cells_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ flags,
                     const uint32_t *__restrict__ scan, uint32_t n_finite, unsigned long long *__restrict__ cell_keys,
                     uint32_t *__restrict__ cell_start, uint32_t n_cells)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) cell_start[n_cells] = n_finite;
    if (j >= n_finite) return;
    if (flags[j]) {
        const uint32_t c = scan[j];
        cell_keys[c] = keys[j];
        cell_start[c] = j;
    }
}

// entries contributed by occupied cell c: itself, its left neighbour unless an earlier cell
// already covers it, its right neighbour unless that cell is occupied
__device__ __forceinline__ void cell_entries(const unsigned long long *cell_keys, uint32_t n_cells, uint32_t c,
                                             bool &left, bool &right)
{
    const unsigned long long k = cell_keys[c];
    left = (c == 0) || (cell_keys[c - 1] + 2 < k);
    right = (c + 1 == n_cells) || (cell_keys[c + 1] > k + 1);
}

__global__ void __launch_bounds__(256)
    map_count_entries_kernel(const unsigned long long *__restrict__ cell_keys, uint32_t n_cells, uint32_t *n_entries)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t cnt = 0;
    if (c < n_cells) {
        bool l, r;
        cell_entries(cell_keys, n_cells, c, l, r);
        cnt = 1u + (l ? 1u : 0u) + (r ? 1u : 0u);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_entries, cnt);
}

__device__ __forceinline__ void hash_insert(unsigned long long *hkeys, uint32_t hmask, unsigned long long k)
{
    uint32_t slot = hash_key(k) & hmask;
    while (true) {
        unsigned long long old = atomicCAS(&hkeys[slot], kEmptyKey, k);
        if (old == kEmptyKey || old == k) return;
        slot = (slot + 1) & hmask;
    }
}

__global__ void __launch_bounds__(256) map_insert_kernel(const unsigned long long *__restrict__ cell_keys,
                                                         uint32_t n_cells, unsigned long long *hkeys, uint32_t hmask)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    bool l, r;
    cell_entries(cell_keys, n_cells, c, l, r);
    const unsigned long long k = cell_keys[c];
    hash_insert(hkeys, hmask, k);
    if (l) hash_insert(hkeys, hmask, k - 1);
    if (r) hash_insert(hkeys, hmask, k + 1);
}

__global__ void __launch_bounds__(256)
    map_fill_kernel(const unsigned long long *__restrict__ hkeys, uint4 *__restrict__ hvals, uint32_t table_size,
                    const unsigned long long *__restrict__ cell_keys, const uint32_t *__restrict__ cell_start,
                    uint32_t n_cells)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= table_size) return;
    const unsigned long long k = hkeys[s];
    if (k == kEmptyKey) return;
    // first cell with key >= k - 1
    uint32_t lo = 0, hi = n_cells;
    const unsigned long long want = k ? k - 1 : 0;  // key 0 has no left neighbour
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (cell_keys[mid] < want) lo = mid + 1;
        else hi = mid;
    }
    uint32_t c = lo;
    uint4 v;
    v.x = cell_start[c];
    if (c < n_cells && cell_keys[c] == k - 1) ++c;
    v.y = cell_start[c];
    if (c < n_cells && cell_keys[c] == k) ++c;
    v.z = cell_start[c];
    if (c < n_cells && cell_keys[c] == k + 1) ++c;
    v.w = cell_start[c];
    hvals[s] = v;
}

static int bit_width_u64(unsigned long long v)
{
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

int build_map_index(MapIndex &m, float cell_size, Scratch &s, cudaStream_t st)
{
    const size_t n = m.n_raw;
    if (n >= (size_t)1 << 31) {
        set_error("target cloud too large for 31-bit indices (%zu)", n);
        return SSF_ERR_INVALID;
    }
    if (!(cell_size > 0.f) || !std::isfinite(cell_size)) {
        set_error("invalid cell size %g", cell_size);
        return SSF_ERR_INVALID;
    }
    m.view = MapView{};
    m.cell_size = cell_size;
    m.n_cells = m.n_entries = m.table_size = 0;
    SSF_TRY(m.small.reserve(16));
    float *bbox_dev = m.small.p;
    uint32_t *cnt_dev = reinterpret_cast<uint32_t *>(m.small.p + 8);  // [0] n_finite, [1] n_cells, [2] n_entries
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
    v.inv_h = 1.0f / cell_size;
    if (n_finite == 0) {  // empty map: every search misses
        v.ox = v.oy = v.oz = 0.f;
        v.nx = v.ny = v.nz = 1;
        SSF_TRY(m.hkeys.reserve(2));
        SSF_TRY(m.hvals.reserve(2));
        SSF_TRY(m.pts.reserve(1));
        SSF_CUDA(cudaMemsetAsync(m.hkeys.p, 0xFF, 2 * sizeof(unsigned long long), st));
        v.hkeys = m.hkeys.p; v.hvals = m.hvals.p; v.hmask = 1; v.pts = m.pts.p; v.nrm = nullptr;
        m.table_size = 2;
        m.view = v;
        return SSF_OK;
    }
    v.ox = hb[0]; v.oy = hb[1]; v.oz = hb[2];
    if (m.sharded) {  // global grid: every rank uses the same origin so cell columns agree
        if (m.shard_origin[0] > hb[0] || m.shard_origin[1] > hb[1] || m.shard_origin[2] > hb[2]) {
            set_error("shard origin must not exceed the shard's own minimum");
            return SSF_ERR_INVALID;
        }
        v.ox = m.shard_origin[0]; v.oy = m.shard_origin[1]; v.oz = m.shard_origin[2];
    }
    // same float expression as cell_coord() on the device (no FMA possible: sub then mul)
    const float ux = (hb[3] - v.ox) * v.inv_h, uy = (hb[4] - v.oy) * v.inv_h, uz = (hb[5] - v.oz) * v.inv_h;
    if (!(ux < 1.0e6f) || !(uy < 65000.f) || !(uz < 65000.f)) {
        set_error("map extent (%g x %g x %g cells of %g m) exceeds the voxel-hash key space", ux, uy, uz, cell_size);
        return SSF_ERR_INVALID;
    }
    v.nx = (int)floorf(ux) + 1; v.ny = (int)floorf(uy) + 1; v.nz = (int)floorf(uz) + 1;
    const int yz_bits = bit_width_u64((unsigned long long)((v.ny > v.nz ? v.ny : v.nz) - 1));
    const unsigned long long sentinel = ((unsigned long long)1 << (2 * yz_bits)) * (unsigned long long)(v.nx + 2);
    const int key_bits = bit_width_u64(sentinel);

    SSF_TRY(m.keys.reserve(n));
    SSF_TRY(m.vals.reserve(n));
    const unsigned blocks_n = (unsigned)((n + 255) / 256);
    map_keys_kernel<<<blocks_n, 256, 0, st>>>(m.raw.p, (uint32_t)n, v.ox, v.oy, v.oz, v.inv_h, v.nx, v.ny, v.nz, sentinel,
                                              m.keys.p, m.vals.p);
    SSF_LAUNCHED();
    SSF_TRY(radix_sort_pairs_u64(m.keys.p, m.vals.p, n, key_bits, s, st));

    SSF_TRY(m.pts.reserve(n_finite));
    if (m.has_normals) SSF_TRY(m.nrm.reserve(n_finite));
    SSF_TRY(m.flags.reserve(n_finite));
    SSF_TRY(m.cell_id.reserve(n_finite));
    const unsigned blocks_f = (n_finite + 255) / 256;
    map_gather_kernel<<<blocks_f, 256, 0, st>>>(m.raw.p, m.has_normals ? m.raw_nrm.p : nullptr,
                                                m.has_global_index ? m.global_index.p : nullptr, m.keys.p, m.vals.p,
                                                n_finite, m.pts.p, m.has_normals ? m.nrm.p : nullptr, m.flags.p);
    SSF_LAUNCHED();
    SSF_TRY(exclusive_scan_u32(m.flags.p, m.cell_id.p, n_finite, cnt_dev + 1, s, st));
    uint32_t n_cells = 0;
    SSF_CUDA(cudaMemcpyAsync(&n_cells, cnt_dev + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SSF_CUDA(cudaStreamSynchronize(st));
    SSF_TRY(m.cell_keys.reserve(n_cells));
    SSF_TRY(m.cell_start.reserve((size_t)n_cells + 1));
    map_cells_kernel<<<blocks_f, 256, 0, st>>>(m.keys.p, m.flags.p, m.cell_id.p, n_finite, m.cell_keys.p, m.cell_start.p,
                                               n_cells);
    SSF_LAUNCHED();
    SSF_CUDA(cudaMemsetAsync(cnt_dev + 2, 0, sizeof(uint32_t), st));
    const unsigned blocks_c = (n_cells + 255) / 256;
    map_count_entries_kernel<<<blocks_c, 256, 0, st>>>(m.cell_keys.p, n_cells, cnt_dev + 2);
    SSF_LAUNCHED();
    uint32_t n_entries = 0;
    SSF_CUDA(cudaMemcpyAsync(&n_entries, cnt_dev + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SSF_CUDA(cudaStreamSynchronize(st));
    uint32_t table = 2;
    while (table < 2u * n_entries && table < (1u << 31)) table <<= 1;
    SSF_TRY(m.hkeys.reserve(table));
    SSF_TRY(m.hvals.reserve(table));
    SSF_CUDA(cudaMemsetAsync(m.hkeys.p, 0xFF, (size_t)table * sizeof(unsigned long long), st));
    map_insert_kernel<<<blocks_c, 256, 0, st>>>(m.cell_keys.p, n_cells, m.hkeys.p, table - 1);
    SSF_LAUNCHED();
    map_fill_kernel<<<(table + 255) / 256, 256, 0, st>>>(m.hkeys.p, m.hvals.p, table, m.cell_keys.p, m.cell_start.p,
                                                        n_cells);
    SSF_LAUNCHED();
    v.pts = m.pts.p;
    v.nrm = m.has_normals ? m.nrm.p : nullptr;
    v.hkeys = m.hkeys.p;
    v.hvals = m.hvals.p;
    v.hmask = table - 1;
    m.view = v;
    m.n_cells = n_cells;
    m.n_entries = n_entries;
    m.table_size = table;
    return SSF_OK;
}

}  // namespace ssf
