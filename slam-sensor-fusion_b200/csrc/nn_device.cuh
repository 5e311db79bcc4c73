// nn_device.cuh -- exact 1-NN of one query against the voxel-hash map (device side of K3).
//
// Replaces kdtree_.nearestKSearch(p, 1, idx, d2) + the `d2 < max_correspondence_dist_` test
// of reference localization/src/icp_point_to_point.cpp:64-70.
//
// Exactness contract (tests/test_gpu_parity.py): for every query the result equals an O(M)
// scan of the whole map that keeps the smallest
//     d2 = fl(fl(fl(dx*dx) + fl(dy*dy)) + fl(dz*dz))      (flann::L2_Simple, no FMA)
// with d2 < limit, ties broken by the lowest ORIGINAL map index.
//
// Why the cell walk is exact: a candidate q can only win if d2(q) <= best, which implies
// |p.k - q.k| <= rb on every axis k for rb = sqrtf(best) * (1 + 1e-6).  cell_coord() is
// monotone, so q's cell lies in [cell(p.k - rb), cell(p.k + rb)] when the interval ends are
// rounded outwards (__fsub_rd / __fadd_ru).  Every cell of that box is visited unless the box
// has shrunk past it, and the box is re-derived whenever best shrinks.
//
// (d2, index) pairs are compared as one 64-bit key: d2 >= 0, so its bit pattern orders like
// the float, and the index in the low word breaks ties towards the lowest index.
//
// Divergence: a lane's work is a sequence of contiguous candidate runs (own cell, its two
// x-neighbours, then up to eight neighbouring rows).  The walk is a FLAT loop -- each trip
// either evaluates four candidates of the current run or sets up the next run -- so a warp
// runs for max-over-lanes of the TOTAL work, not the sum over rows of per-row maxima.
#pragma once
#include <cfloat>

#include "common.cuh"

namespace ssf {

#ifdef SSF_NN_STATS
// debug build only (make stats): [0] eval4 calls, [1] probes, [2] probe slots read, [3] queries,
// [4] wide-path queries, [5] matched queries, [6] outer loop trips, [7] sum over warps of max trips
static __device__ unsigned long long g_nn_stats[8];  // one copy per translation unit (icp_kernels.cu reports its own)
#define NN_STAT(i, v) atomicAdd(&g_nn_stats[i], (unsigned long long)(v))
#else
#define NN_STAT(i, v) ((void)0)
#endif

#ifdef SSF_BOUNDS
// debug build only (make bounds): trap on any index outside its array
#define SSF_CHECK(cond) do { if (!(cond)) { printf("SSF_CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); __trap(); } } while (0)
#else
#define SSF_CHECK(cond) ((void)0)
#endif

struct NNHit {
    float d2;      // best squared distance (== limit when nothing was found)
    int idx;       // original index of the best target point, -1 if none
    uint32_t pos;  // its position in the sorted cloud (valid when idx >= 0)
};

struct CellBox {
    int x0, x1, y0, y1, z0, z1;
};

__device__ __forceinline__ CellBox cell_box(const MapView &m, float px, float py, float pz, float best)
{
    // 1e-6 relative slack covers the rounding of dx, dx*dx and the sums (< 2^-22 in total);
    // the absolute slack covers products that underflow to zero
    const float rb = __fadd_ru(__fmul_ru(__fsqrt_ru(best), 1.000001f), 1e-18f);
    CellBox b;
    b.x0 = max(cell_coord(__fsub_rd(px, rb), m.ox, m.inv_h, m.nx), 0);
    b.x1 = min(cell_coord(__fadd_ru(px, rb), m.ox, m.inv_h, m.nx), m.nx - 1);
    b.y0 = max(cell_coord(__fsub_rd(py, rb), m.oy, m.inv_h, m.ny), 0);
    b.y1 = min(cell_coord(__fadd_ru(py, rb), m.oy, m.inv_h, m.ny), m.ny - 1);
    b.z0 = max(cell_coord(__fsub_rd(pz, rb), m.oz, m.inv_h, m.nz), 0);
    b.z1 = min(cell_coord(__fadd_ru(pz, rb), m.oz, m.inv_h, m.nz), m.nz - 1);
    return b;
}

__device__ __forceinline__ void shrink_box(CellBox &b, const CellBox &nb)
{
    b.x0 = max(b.x0, nb.x0); b.x1 = min(b.x1, nb.x1);
    b.y0 = max(b.y0, nb.y0); b.y1 = min(b.y1, nb.y1);
    b.z0 = max(b.z0, nb.z0); b.z1 = min(b.z1, nb.z1);
}

// probe the table for the entry centred on cell (cx, cy, cz); false when none of
// cx-1, cx, cx+1 holds a point
__device__ __forceinline__ bool probe(const MapView &m, int cx, int cy, int cz, uint4 &v)
{
    const unsigned long long k = cell_key(cx, cy, cz, m.nx);
    uint32_t slot = hash_key(k) & m.hmask;
    NN_STAT(1, 1);
    while (true) {
        NN_STAT(2, 1);
        SSF_CHECK(slot <= m.hmask);
        const unsigned long long t = __ldg(&m.hkeys[slot]);
        if (t == k) {
            SSF_CHECK(m.hvals[slot].x <= m.hvals[slot].y && m.hvals[slot].y <= m.hvals[slot].z &&
                      m.hvals[slot].z <= m.hvals[slot].w && m.hvals[slot].w <= m.n_pts);
            v = __ldg(&m.hvals[slot]);
            return true;
        }
        if (t == kEmptyKey) return false;
        slot = (slot + 1) & m.hmask;
    }
}

__device__ __forceinline__ unsigned long long cand_key(const float4 &q, float px, float py, float pz)
{
    const float dx = __fsub_rn(px, q.x), dy = __fsub_rn(py, q.y), dz = __fsub_rn(pz, q.z);
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)__float_as_uint(q.w);
}

// four candidates of run [j, e) per call; indices past the end re-read the last one (harmless)
__device__ __forceinline__ void eval4(const MapView &m, uint32_t j, uint32_t e, float px, float py, float pz,
                                      unsigned long long &best, uint32_t &pos)
{
    NN_STAT(0, 1);
    SSF_CHECK(j < e && e <= m.n_pts);
    const uint32_t last = e - 1;
    const uint32_t j0 = j, j1 = min(j + 1, last), j2 = min(j + 2, last), j3 = min(j + 3, last);
    const float4 q0 = __ldg(&m.pts[j0]), q1 = __ldg(&m.pts[j1]), q2 = __ldg(&m.pts[j2]), q3 = __ldg(&m.pts[j3]);
    const unsigned long long k0 = cand_key(q0, px, py, pz), k1 = cand_key(q1, px, py, pz),
                             k2 = cand_key(q2, px, py, pz), k3 = cand_key(q3, px, py, pz);
    if (k0 < best) { best = k0; pos = j0; }
    if (k1 < best) { best = k1; pos = j1; }
    if (k2 < best) { best = k2; pos = j2; }
    if (k3 < best) { best = k3; pos = j3; }
}

// general walk for boxes wider than 3 cells on some axis (threshold radius > cell edge)
static __device__ __noinline__ void nn_walk_wide(const MapView &m, CellBox b, float px, float py, float pz, float limit,
                                          unsigned long long &best, uint32_t &pos)
{
    uint32_t boxed_bits = __float_as_uint(limit);
    for (int cz = b.z0; cz <= b.z1; ++cz)
        for (int cy = b.y0; cy <= b.y1; ++cy) {
            if ((uint32_t)(best >> 32) < boxed_bits) {
                boxed_bits = (uint32_t)(best >> 32);
                shrink_box(b, cell_box(m, px, py, pz, __uint_as_float(boxed_bits)));
            }
            if (cz < b.z0 || cz > b.z1 || cy < b.y0 || cy > b.y1) continue;
            for (int c = b.x0; c <= b.x1; c += 3) {
                uint4 v;
                if (!probe(m, c + 1, cy, cz, v)) continue;
                const int more = b.x1 - c;
                const uint32_t e = more >= 2 ? v.w : (more == 1 ? v.z : v.y);
                for (uint32_t j = v.x; j < e; j += 4) eval4(m, j, e, px, py, pz, best, pos);
            }
        }
}

// Visiting order of the neighbourhood after the own cell.  (sx, sy, sz) point to the NEAR side
// of the own cell on each axis (the side the query is closer to); near neighbours come first so
// the best distance shrinks before the far side is even probed:
//   item 0 own row, near x | 1 own row, far x | 2 row (sy,0) | 3 row (0,sz) | 4 row (sy,sz)
//        5 row (-sy,0) | 6 row (0,-sz) | 7 row (sy,-sz) | 8 row (-sy,sz) | 9 row (-sy,-sz)
// packed as 2-bit fields: multiplier of sy (resp. sz) + 1, field k = item k + 2
constexpr uint32_t pack8(int a0, int a1, int a2, int a3, int a4, int a5, int a6, int a7)
{
    return (uint32_t)(a0 + 1) | (uint32_t)(a1 + 1) << 2 | (uint32_t)(a2 + 1) << 4 | (uint32_t)(a3 + 1) << 6 |
           (uint32_t)(a4 + 1) << 8 | (uint32_t)(a5 + 1) << 10 | (uint32_t)(a6 + 1) << 12 | (uint32_t)(a7 + 1) << 14;
}
constexpr uint32_t kRowY = pack8(1, 0, 1, -1, 0, 1, -1, -1);
constexpr uint32_t kRowZ = pack8(0, 1, 1, 0, -1, -1, 1, -1);

// limit: accept only d2 < limit (strict), like the reference's threshold test
__device__ __forceinline__ NNHit nn_query(const MapView &m, float px, float py, float pz, float limit)
{
    NNHit h;
    h.d2 = limit;
    h.idx = -1;
    h.pos = 0;
    if (!(limit > 0.f) || !isfinite(px) || !isfinite(py) || !isfinite(pz)) return h;
    CellBox b = cell_box(m, px, py, pz, limit);
    if (b.x0 > b.x1 || b.y0 > b.y1 || b.z0 > b.z1) return h;
    const unsigned long long none = (unsigned long long)__float_as_uint(limit) << 32;
    unsigned long long best = none;
    uint32_t pos = 0;
    const int cx = cell_coord(px, m.ox, m.inv_h, m.nx), cy = cell_coord(py, m.oy, m.inv_h, m.ny),
              cz = cell_coord(pz, m.oz, m.inv_h, m.nz);
    const bool narrow = b.x0 >= cx - 1 && b.x1 <= cx + 1 && b.y0 >= cy - 1 && b.y1 <= cy + 1 && b.z0 >= cz - 1 &&
                        b.z1 <= cz + 1;
    NN_STAT(3, 1);
    if (!narrow) {
        NN_STAT(4, 1);
        nn_walk_wide(m, b, px, py, pz, limit, best, pos);
    } else {
        // Every run of this query comes from an entry centred on column cx.  Instead of
        // re-deriving the box, each of the six neighbour directions gets a threshold t such that
        // best_d2 < t proves the box no longer reaches that neighbour: the neighbour's cells
        // start at the float where cell_coord() flips, which differs from o + c*h by a few
        // roundings; `slack` over-covers that, so a neighbour is only ever skipped when the
        // monotone box of the header comment excludes it too.
        const float h = __frcp_rn(m.inv_h);
        float t_lo[3], t_hi[3];
        {
            const float pk[3] = {px, py, pz}, ok[3] = {m.ox, m.oy, m.oz};
            const int ck[3] = {cx, cy, cz};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float lo = __fadd_rn(ok[k], __fmul_rn((float)ck[k], h));        // ~ lower face of the own cell
                const float hi = __fadd_rn(ok[k], __fmul_rn((float)(ck[k] + 1), h));  // ~ upper face
                const float slack = __fmul_rn(4e-6f, fabsf(pk[k]) + fabsf(ok[k]) + __fmul_rn(fabsf((float)ck[k]) + 2.f, h));
                const float rl = __fsub_rd(__fsub_rd(pk[k], lo), slack), rh = __fsub_rd(__fsub_rd(hi, pk[k]), slack);
                // rb = sqrt(best)*(1+1e-6)+1e-18 < r  <=  best < r^2*(1-4e-6) - 1e-30   (r > 0)
                t_lo[k] = rl > 0.f ? __fsub_rd(__fmul_rd(__fmul_rd(rl, rl), 0.999996f), 1e-30f) : 0.f;
                t_hi[k] = rh > 0.f ? __fsub_rd(__fmul_rd(__fmul_rd(rh, rh), 0.999996f), 1e-30f) : 0.f;
            }
        }
        // rows / columns outside the initial box (grid edge, or radius < cell) are never visited
        const bool in_xl = b.x0 <= cx - 1, in_xh = b.x1 >= cx + 1, in_x0 = cx >= b.x0 && cx <= b.x1;
        const float uy = __fmul_rn(__fsub_rn(py, m.oy), m.inv_h), uz = __fmul_rn(__fsub_rn(pz, m.oz), m.inv_h),
                    ux = __fmul_rn(__fsub_rn(px, m.ox), m.inv_h);
        const bool near_left = !((ux - floorf(ux)) > 0.5f);
        const int sy = (uy - floorf(uy)) > 0.5f ? 1 : -1, sz = (uz - floorf(uz)) > 0.5f ? 1 : -1;
        // 1. own cell: one probe, one run; every lane of the warp does this together
        uint4 own = make_uint4(0, 0, 0, 0);
        const bool own_ok = cy >= b.y0 && cy <= b.y1 && cz >= b.z0 && cz <= b.z1 && probe(m, cx, cy, cz, own);
        if (own_ok && in_x0)
            for (uint32_t j = own.y; j < own.z; j += 4) eval4(m, j, own.z, px, py, pz, best, pos);
        // 2. which of the ten neighbour items can still hold a better point?
        // A neighbour block's squared distance is at least the SUM of its per-axis squared gaps
        // (each t_* under-estimates one gap^2; the sum is rounded down), so diagonal rows and the
        // x-cells of a row are pruned by the box bound, not axis by axis.
        float bd = __uint_as_float((uint32_t)(best >> 32));
        uint32_t mask = 0;
        {
            const bool xl = own_ok && in_xl && !(bd < t_lo[0]), xh = own_ok && in_xh && !(bd < t_hi[0]);
            mask = ((near_left ? xl : xh) ? 1u : 0u) | ((near_left ? xh : xl) ? 2u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int dy = sy * ((int)((kRowY >> (2 * k)) & 3u) - 1), dz = sz * ((int)((kRowZ >> (2 * k)) & 3u) - 1);
                const int ry = cy + dy, rz = cz + dz;
                const float lb = __fadd_rd(dy < 0 ? t_lo[1] : (dy > 0 ? t_hi[1] : 0.f), dz < 0 ? t_lo[2] : (dz > 0 ? t_hi[2] : 0.f));
                const bool out = ry < b.y0 || ry > b.y1 || rz < b.z0 || rz > b.z1 || bd < lb;
                mask |= out ? 0u : (4u << k);
            }
        }
        // 3. flat walk over the remaining items: each trip evaluates four candidates or pops an item
        uint32_t j = 0, e = 0;
        while (true) {
            if (j >= e) {
                bool got = false;
                while (mask) {
                    const int s = __ffs(mask) - 1;
                    mask &= mask - 1;
                    bd = __uint_as_float((uint32_t)(best >> 32));
                    const bool xl = in_xl && !(bd < t_lo[0]), xh = in_xh && !(bd < t_hi[0]);
                    if (s < 2) {
                        const bool left = (s == 0) == near_left;
                        if (!(left ? xl : xh)) continue;
                        j = left ? own.x : own.z;
                        e = left ? own.y : own.w;
                    } else {
                        const int k = s - 2;
                        const int dy = sy * ((int)((kRowY >> (2 * k)) & 3u) - 1), dz = sz * ((int)((kRowZ >> (2 * k)) & 3u) - 1);
                        const float lb = __fadd_rd(dy < 0 ? t_lo[1] : (dy > 0 ? t_hi[1] : 0.f), dz < 0 ? t_lo[2] : (dz > 0 ? t_hi[2] : 0.f));
                        if (bd < lb) continue;
                        uint4 v;
                        if (!probe(m, cx, cy + dy, cz + dz, v)) continue;
                        const bool rl = in_xl && !(bd < __fadd_rd(lb, t_lo[0])), rh = in_xh && !(bd < __fadd_rd(lb, t_hi[0]));
                        j = rl ? v.x : (in_x0 ? v.y : v.z);
                        e = rh ? v.w : (in_x0 ? v.z : v.y);
                    }
                    if (j < e) {
                        got = true;
                        break;
                    }
                }
                if (!got) break;
            }
            eval4(m, j, e, px, py, pz, best, pos);
            j += 4;
            NN_STAT(6, 1);
        }
    }
    if (best < none) {
        NN_STAT(5, 1);
        h.d2 = __uint_as_float((uint32_t)(best >> 32));
        h.idx = (int)(uint32_t)(best & 0xFFFFFFFFull);
        h.pos = pos;
    }
    return h;
}

// ---------------------------------------------------------------------------------------------------
// Warp-cooperative form.  Must be called by all 32 lanes of a warp (inactive lanes pass active =
// false).  Each lane scans its own cell exactly like nn_query; the neighbour items that survive
// the pruning mask are then published to a shared-memory queue and drained by ALL lanes, one
// (owner, item) per lane per round, so the neighbour phase no longer runs with a handful of
// lanes.  Results merge through 64-bit atomicMin on the packed (d2, index) key -- a minimum over
// the same candidate set in any order -- so the outcome is identical to the sequential walk.
// The position of the winner travels in a second key (d2, position); when an exact cross-cell
// tie makes the two disagree (checked), the lane falls back to the sequential walk.
// ---------------------------------------------------------------------------------------------------
struct WarpNN {
    float px[32], py[32], pz[32];
    int cx[32], cy[32], cz[32];
    float t_lo[3][32], t_hi[3][32];
    uint4 own[32];
    uint32_t flags[32];  // bit0 in_xl, bit1 in_xh, bit2 in_x0, bit3 near_left, bit4 sy > 0, bit5 sz > 0
    unsigned long long key[32], pkey[32];
    unsigned short queue[32 * 10];
};

__device__ __forceinline__ NNHit nn_query_warp(const MapView &m, float px, float py, float pz, float limit, bool active,
                                               WarpNN &w)
{
    const int lane = threadIdx.x & 31;
    NNHit h;
    h.d2 = limit;
    h.idx = -1;
    h.pos = 0;
    const unsigned long long none = (unsigned long long)__float_as_uint(limit) << 32;
    unsigned long long best = none;
    uint32_t pos = 0, mask = 0;
    active = active && limit > 0.f && isfinite(px) && isfinite(py) && isfinite(pz);
    CellBox b = {0, -1, 0, -1, 0, -1};
    if (active) b = cell_box(m, px, py, pz, limit);
    active = active && !(b.x0 > b.x1 || b.y0 > b.y1 || b.z0 > b.z1);
    const int cx = cell_coord(px, m.ox, m.inv_h, m.nx), cy = cell_coord(py, m.oy, m.inv_h, m.ny),
              cz = cell_coord(pz, m.oz, m.inv_h, m.nz);
    const bool narrow = b.x0 >= cx - 1 && b.x1 <= cx + 1 && b.y0 >= cy - 1 && b.y1 <= cy + 1 && b.z0 >= cz - 1 &&
                        b.z1 <= cz + 1;
    if (active) NN_STAT(3, 1);
    if (active && !narrow) {
        NN_STAT(4, 1);
        nn_walk_wide(m, b, px, py, pz, limit, best, pos);
    }
    float t_lo[3] = {0.f, 0.f, 0.f}, t_hi[3] = {0.f, 0.f, 0.f};
    uint4 own = make_uint4(0, 0, 0, 0);
    uint32_t flags = 0;
    if (active && narrow) {
        const float hcell = __frcp_rn(m.inv_h);
        const float pk[3] = {px, py, pz}, ok[3] = {m.ox, m.oy, m.oz};
        const int ck[3] = {cx, cy, cz};
#pragma unroll
        for (int k = 0; k < 3; ++k) {  // same thresholds as nn_query
            const float lo = __fadd_rn(ok[k], __fmul_rn((float)ck[k], hcell));
            const float hi = __fadd_rn(ok[k], __fmul_rn((float)(ck[k] + 1), hcell));
            const float slack = __fmul_rn(4e-6f, fabsf(pk[k]) + fabsf(ok[k]) + __fmul_rn(fabsf((float)ck[k]) + 2.f, hcell));
            const float rl = __fsub_rd(__fsub_rd(pk[k], lo), slack), rh = __fsub_rd(__fsub_rd(hi, pk[k]), slack);
            t_lo[k] = rl > 0.f ? __fsub_rd(__fmul_rd(__fmul_rd(rl, rl), 0.999996f), 1e-30f) : 0.f;
            t_hi[k] = rh > 0.f ? __fsub_rd(__fmul_rd(__fmul_rd(rh, rh), 0.999996f), 1e-30f) : 0.f;
        }
        const bool in_xl = b.x0 <= cx - 1, in_xh = b.x1 >= cx + 1, in_x0 = cx >= b.x0 && cx <= b.x1;
        const float uy = __fmul_rn(__fsub_rn(py, m.oy), m.inv_h), uz = __fmul_rn(__fsub_rn(pz, m.oz), m.inv_h),
                    ux = __fmul_rn(__fsub_rn(px, m.ox), m.inv_h);
        const bool near_left = !((ux - floorf(ux)) > 0.5f);
        const int sy = (uy - floorf(uy)) > 0.5f ? 1 : -1, sz = (uz - floorf(uz)) > 0.5f ? 1 : -1;
        flags = (in_xl ? 1u : 0u) | (in_xh ? 2u : 0u) | (in_x0 ? 4u : 0u) | (near_left ? 8u : 0u) | (sy > 0 ? 16u : 0u) |
                (sz > 0 ? 32u : 0u);
        const bool own_ok = cy >= b.y0 && cy <= b.y1 && cz >= b.z0 && cz <= b.z1 && probe(m, cx, cy, cz, own);
        if (own_ok && in_x0)
            for (uint32_t j = own.y; j < own.z; j += 4) eval4(m, j, own.z, px, py, pz, best, pos);
        const float bd = __uint_as_float((uint32_t)(best >> 32));
        const bool xl = own_ok && in_xl && !(bd < t_lo[0]), xh = own_ok && in_xh && !(bd < t_hi[0]);
        mask = ((near_left ? xl : xh) ? 1u : 0u) | ((near_left ? xh : xl) ? 2u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int dy = sy * ((int)((kRowY >> (2 * k)) & 3u) - 1), dz = sz * ((int)((kRowZ >> (2 * k)) & 3u) - 1);
            const int ry = cy + dy, rz = cz + dz;
            const float lb = __fadd_rd(dy < 0 ? t_lo[1] : (dy > 0 ? t_hi[1] : 0.f), dz < 0 ? t_lo[2] : (dz > 0 ? t_hi[2] : 0.f));
            const bool out = ry < b.y0 || ry > b.y1 || rz < b.z0 || rz > b.z1 || bd < lb;
            mask |= out ? 0u : (4u << k);
        }
    }
    // ---- publish the lane's query and queue its items ----
    w.px[lane] = px; w.py[lane] = py; w.pz[lane] = pz;
    w.cx[lane] = cx; w.cy[lane] = cy; w.cz[lane] = cz;
#pragma unroll
    for (int k = 0; k < 3; ++k) { w.t_lo[k][lane] = t_lo[k]; w.t_hi[k][lane] = t_hi[k]; }
    w.own[lane] = own;
    w.flags[lane] = flags;
    w.key[lane] = best;
    w.pkey[lane] = best < none ? ((best & 0xFFFFFFFF00000000ull) | pos) : ~0ull;
    const uint32_t cnt = __popc(mask);
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    {
        uint32_t at = incl - cnt, mm = mask;
        while (mm) {
            const int s = __ffs(mm) - 1;
            mm &= mm - 1;
            SSF_CHECK(at < 320);
            w.queue[at++] = (unsigned short)((lane << 4) | s);
        }
    }
    __syncwarp();
    // ---- drain: one (owner, item) per lane per round ----
    for (uint32_t base = 0; base < total; base += 32) {
        const uint32_t i = base + lane;
        if (i < total) {
            const uint32_t it = w.queue[i];
            const int o = it >> 4, s = it & 15;
            const float opx = w.px[o], opy = w.py[o], opz = w.pz[o];
            const uint32_t fl = w.flags[o];
            unsigned long long lbest = *(volatile unsigned long long *)&w.key[o];
            const unsigned long long snap = lbest;
            const float bd = __uint_as_float((uint32_t)(lbest >> 32));
            const bool xl = (fl & 1u) && !(bd < w.t_lo[0][o]), xh = (fl & 2u) && !(bd < w.t_hi[0][o]);
            uint32_t j = 0, e = 0;
            if (s < 2) {
                const bool left = (s == 0) == ((fl & 8u) != 0);
                if (left ? xl : xh) {
                    const uint4 ow = w.own[o];
                    j = left ? ow.x : ow.z;
                    e = left ? ow.y : ow.w;
                }
            } else {
                const int k = s - 2;
                const int dy = ((fl & 16u) ? 1 : -1) * ((int)((kRowY >> (2 * k)) & 3u) - 1);
                const int dz = ((fl & 32u) ? 1 : -1) * ((int)((kRowZ >> (2 * k)) & 3u) - 1);
                const float lb = __fadd_rd(dy < 0 ? w.t_lo[1][o] : (dy > 0 ? w.t_hi[1][o] : 0.f),
                                           dz < 0 ? w.t_lo[2][o] : (dz > 0 ? w.t_hi[2][o] : 0.f));
                uint4 v;
                if (!(bd < lb) && probe(m, w.cx[o], w.cy[o] + dy, w.cz[o] + dz, v)) {
                    const bool rl = (fl & 1u) && !(bd < __fadd_rd(lb, w.t_lo[0][o]));
                    const bool rh = (fl & 2u) && !(bd < __fadd_rd(lb, w.t_hi[0][o]));
                    j = rl ? v.x : ((fl & 4u) ? v.y : v.z);
                    e = rh ? v.w : ((fl & 4u) ? v.z : v.y);
                }
            }
            uint32_t lpos = 0;
            for (; j < e; j += 4) {
                eval4(m, j, e, opx, opy, opz, lbest, lpos);
                NN_STAT(6, 1);
            }
            if (lbest < snap) {
                atomicMin(&w.key[o], lbest);
                atomicMin(&w.pkey[o], (lbest & 0xFFFFFFFF00000000ull) | lpos);
            }
        }
        __syncwarp();
    }
    best = w.key[lane];
    const unsigned long long pk = w.pkey[lane];
    __syncwarp();  // the scratch is reused by the next call
    if (best < none) {
        const int idx = (int)(uint32_t)(best & 0xFFFFFFFFull);
        uint32_t p2 = (uint32_t)(pk & 0xFFFFFFFFull);
        const bool consistent = (pk >> 32) == (best >> 32) && p2 < m.n_pts && __float_as_int(__ldg(&m.pts[p2]).w) == idx;
        if (!consistent) return nn_query(m, px, py, pz, limit);  // exact cross-cell tie: sequential walk decides
        NN_STAT(5, 1);
        h.d2 = __uint_as_float((uint32_t)(best >> 32));
        h.idx = idx;
        h.pos = p2;
    }
    return h;
}

// reference applyTransformation (icp_point_to_point.cpp:103-105): ((T0*x + T1*y) + T2*z) + T3
// with every product and sum rounded (the reference is built without FMA contraction)
__device__ __forceinline__ float affine_row(float a, float b, float c, float d, float x, float y, float z)
{
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, x), __fmul_rn(b, y)), __fmul_rn(c, z)), d);
}

// T column-major 4x4
__device__ __forceinline__ float3 transform_point(const float *T, float x, float y, float z)
{
    float3 r;
    r.x = affine_row(T[0], T[4], T[8], T[12], x, y, z);
    r.y = affine_row(T[1], T[5], T[9], T[13], x, y, z);
    r.z = affine_row(T[2], T[6], T[10], T[14], x, y, z);
    return r;
}

// ---- warp reduction of 32 doubles per thread ---------------------------------------------------
// After the call lane l holds the warp-wide sum of v[l] (fixed tree order -> deterministic).
__device__ __forceinline__ double warp_transpose_reduce32(double (&v)[32])
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const double send = up ? v[i] : v[i + half];
            const double keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

}  // namespace ssf
