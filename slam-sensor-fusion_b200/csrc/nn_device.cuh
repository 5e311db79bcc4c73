// nn_device.cuh -- exact 1-NN of one query against the voxel-hash map (device side of K3).
//
// Replaces kdtree_.nearestKSearch(p, 1, idx, d2) + the `d2 < max_correspondence_dist_` test
// of reference localization/src/icp_point_to_point.cpp:64-70.
//
// Exactness contract (tests/test_nn_parity.py): for every query the result equals an
// O(M) scan of the whole map that keeps the smallest
//     d2 = fl(fl(fl(dx*dx) + fl(dy*dy)) + fl(dz*dz))      (flann::L2_Simple, no FMA)
// with d2 < limit, ties broken by the lowest ORIGINAL map index.
//
// Why the cell walk is exact: a candidate q can only win if d2(q) <= best, which implies
// |p.k - q.k| <= rb on every axis k for rb = sqrtf(best) * (1 + 1e-6).  cell_coord() is
// monotone, so q's cell lies in [cell(p.k - rb), cell(p.k + rb)] when the interval ends
// are rounded outwards (__fsub_rd / __fadd_ru).  Every row/cell in that box is visited,
// and the box is re-derived whenever best shrinks.
#pragma once
#include <cfloat>

#include "common.cuh"

namespace ssf {

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

// probe the table for the entry centred on cell (cx, cy, cz); false when none of
// cx-1, cx, cx+1 holds a point
__device__ __forceinline__ bool probe(const MapView &m, int cx, int cy, int cz, uint4 &v)
{
    const unsigned long long k = cell_key(cx, cy, cz, m.nx);
    uint32_t slot = hash_key(k) & m.hmask;
    while (true) {
        const unsigned long long t = __ldg(&m.hkeys[slot]);
        if (t == k) {
            v = __ldg(&m.hvals[slot]);
            return true;
        }
        if (t == kEmptyKey) return false;
        slot = (slot + 1) & m.hmask;
    }
}

__device__ __forceinline__ void scan_run(const MapView &m, uint32_t s, uint32_t e, float px, float py, float pz,
                                         NNHit &h)
{
    for (uint32_t j = s; j < e; ++j) {
        const float4 q = __ldg(&m.pts[j]);
        const float dx = __fsub_rn(px, q.x), dy = __fsub_rn(py, q.y), dz = __fsub_rn(pz, q.z);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        const int qi = __float_as_int(q.w);
        if (d2 < h.d2 || (d2 == h.d2 && h.idx >= 0 && qi < h.idx)) {
            h.d2 = d2;
            h.idx = qi;
            h.pos = j;
        }
    }
}

// visit cells [b.x0, b.x1] of row (cy, cz)
__device__ __forceinline__ void visit_row(const MapView &m, const CellBox &b, int cy, int cz, float px, float py,
                                          float pz, NNHit &h)
{
    for (int c = b.x0; c <= b.x1; c += 3) {
        uint4 v;
        if (!probe(m, c + 1, cy, cz, v)) continue;
        const int last = b.x1 - c;  // 0, 1 or >= 2 further cells wanted
        const uint32_t e = last >= 2 ? v.w : (last == 1 ? v.z : v.y);
        scan_run(m, v.x, e, px, py, pz, h);
    }
}

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
    // the query's own row first: it almost always holds the answer and shrinks the box
    const int cyc = min(max(cell_coord(py, m.oy, m.inv_h, m.ny), b.y0), b.y1);
    const int czc = min(max(cell_coord(pz, m.oz, m.inv_h, m.nz), b.z0), b.z1);
    visit_row(m, b, cyc, czc, px, py, pz, h);
    float boxed_for = limit;
    for (int cz = b.z0; cz <= b.z1; ++cz) {
        for (int cy = b.y0; cy <= b.y1; ++cy) {
            if (cy == cyc && cz == czc) continue;
            if (h.d2 < boxed_for) {  // best shrank: shrink the box (never grows)
                const CellBox nb = cell_box(m, px, py, pz, h.d2);
                b.x0 = max(b.x0, nb.x0); b.x1 = min(b.x1, nb.x1);
                b.y0 = max(b.y0, nb.y0); b.y1 = min(b.y1, nb.y1);
                b.z0 = max(b.z0, nb.z0); b.z1 = min(b.z1, nb.z1);
                boxed_for = h.d2;
            }
            if (cz < b.z0 || cz > b.z1 || cy < b.y0 || cy > b.y1) continue;
            visit_row(m, b, cy, cz, px, py, pz, h);
        }
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

// ---- warp / block reduction of up to 32 doubles per thread ------------------------------------
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
