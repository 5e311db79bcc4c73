// nn_device.cuh -- exact 1-NN of one query against the voxel-grid map (device side of K3).
//
// Replaces kdtree_.nearestKSearch(p, 1, idx, d2) + the `d2 < max_correspondence_dist_` test
// of reference localization/src/icp_point_to_point.cpp:64-70.
//
// Exactness contract (tests/test_gpu_parity.py): for every query the result equals an O(M)
// scan of the whole map that keeps the smallest
//     d2 = fl(fl(fl(dx*dx) + fl(dy*dy)) + fl(dz*dz))      (flann::L2_Simple, no FMA)
// with d2 < limit, ties broken by the lowest ORIGINAL map index.
//
// The walk: the query's own row first (cells cx-1..cx+1, one directory load), then the rows
// around it ring by ring (ring r = rows with max(|dy|, |dz|) == r).  Each row is reduced to the
// x cells the current best distance can still reach, and those cells are ONE contiguous run of
// the sorted cloud.  The walk ends at the first ring whose nearest row is farther than the best
// distance -- for any threshold, with no a-priori bound on the number of rings.
//
// Why skipping is safe.  Let D be the float d2 of a candidate and S its real squared distance:
// D >= S * (1 - 4e-7) (three products, three differences, two sums, each within 2^-24).  A
// candidate can only change the result if D <= best.  For a row (or x cell) whose every point
// is at least g_y, g_z (g_x) away from the query along the axes, S >= g_x^2 + g_y^2 + g_z^2, so
// the row is skipped only when best < (g_y^2 + g_z^2) * (1 - 2e-6), and the x range kept is
// |dx|^2 <= best * (1 + 2e-6) - (g_y^2 + g_z^2).  The gaps g are UNDER-estimates: they are
// derived from the query's cell coordinate u = fl(fl(p - o) * inv_h) (the same expression
// that bins the map) minus 4e-7 * (|u| + 2) cells of slack, which covers the rounding of u for
// both the query and any map point on the other side of the face, times an under-estimate of
// the cell edge.  (d2, index) pairs are compared as one 64-bit key: d2 >= 0, so its bit pattern
// orders like the float, and the index in the low word breaks ties towards the lowest index.
#pragma once
#include <cfloat>

#include "common.cuh"

namespace ssf {

#ifdef SSF_NN_STATS
// debug build only (make stats): [0] eval4 calls, [1] directory loads, [2] rows visited,
// [3] queries, [4] queries that went past ring 1, [5] matched queries, [6] non-empty runs
static __device__ unsigned long long g_nn_stats[8];  // one copy per translation unit (icp_kernels.cu reports its own)
#define NN_STAT(i, v) atomicAdd(&g_nn_stats[i], (unsigned long long)(v))
#else
#define NN_STAT(i, v) ((void)0)
#endif

struct NNHit {
    float d2;      // best squared distance (== limit when nothing was found)
    int idx;       // original index of the best target point, -1 if none
    uint32_t pos;  // its position in the sorted cloud (valid when idx >= 0)
};

constexpr float kShrink = 0.999998f;  // (1 - 2e-6): applied to squared gaps
constexpr float kGrow = 1.000002f;    // (1 + 2e-6): applied to the best distance

// one-instruction square root (MUFU, relative error <= 2^-23) pushed to a guaranteed upper / lower
// bound; the absolute term covers flushed subnormals
__device__ __forceinline__ float sqrt_up(float x)
{
    float s;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(x));
    return __fadd_ru(__fmul_ru(s, 1.0000005f), 1e-18f);
}
__device__ __forceinline__ float sqrt_dn(float x)
{
    float s;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(x));
    return fmaxf(__fsub_rd(__fmul_rd(s, 0.9999995f), 1e-18f), 0.f);
}

// cell coordinate of v plus under-estimates (metres) of the distance from v to the cells
// c - 1 (dn) and c + 1 (up)
struct AxisGap {
    int c;
    float dn, up;
};

__device__ __forceinline__ AxisGap axis_gap(float v, float o, float inv_h, float hq, int n)
{
    float u = __fmul_rn(__fsub_rn(v, o), inv_h);  // == cell_coord()
    u = fminf(fmaxf(u, -2.0f), (float)n + 1.0f);
    const float fl = floorf(u);
    const float f = __fsub_rn(u, fl);  // exact
    const float su = __fmul_rn(4e-7f, __fadd_rn(fabsf(u), 2.0f));
    AxisGap a;
    a.c = (int)fl;
    a.dn = fmaxf(__fmul_rd(__fsub_rd(f, su), hq), 0.f);
    a.up = fmaxf(__fmul_rd(__fsub_rd(__fsub_rd(1.0f, f), su), hq), 0.f);
    return a;
}

// squared gap, shrunk: a lower bound of what the float d2 of any point behind the gap can be
__device__ __forceinline__ float gap_sq(float g) { return fmaxf(__fsub_rd(__fmul_rd(__fmul_rd(g, g), kShrink), 1e-30f), 0.f); }

// ---- running result of one walk ------------------------------------------------------------------
// key = (d2 bits << 32) | original index of the best candidate so far (or the `none` sentinel).
// CERT walks additionally keep what a later search needs to be SKIPPED (see nn_verify):
//   b2   second-smallest d2 among the candidates evaluated (every point is evaluated at most once)
//   bdm  the pruning bound, (sqrt(best d2) + mu)^2 instead of best d2: rows and cells are only
//        skipped when they lie farther than the best distance PLUS the margin mu, so at the end
//        every point that was never looked at is known to be farther than sqrt(d2 best) + mu.
template <bool CERT>
struct NNBest {
    unsigned long long key;
    uint32_t pos;
    uint32_t skip;  // position of a candidate evaluated up front (the previous neighbour): not counted twice
    float b2, bdm, mu;
    __device__ __forceinline__ float bd() const { return __uint_as_float((uint32_t)(key >> 32)); }
    // bound to prune with
    __device__ __forceinline__ float prune() const { return CERT ? bdm : bd(); }
    __device__ __forceinline__ void refresh()
    {
        if (CERT) {  // (s + mu)^2 with s >= sqrt(bd), everything rounded up
            const float r = __fadd_ru(sqrt_up(bd()), mu);
            bdm = __fmul_ru(r, r);
        }
    }
};

template <bool CERT>
__device__ __forceinline__ void cand_update(NNBest<CERT> &B, const float4 &q, uint32_t j, bool valid, float px,
                                            float py, float pz)
{
    const float dx = __fsub_rn(px, q.x), dy = __fsub_rn(py, q.y), dz = __fsub_rn(pz, q.z);
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    const unsigned long long k = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)__float_as_uint(q.w);
    if (CERT) {  // padding slots (valid == false) must not pose as a second point at the same distance
        const float dv = (valid && j != B.skip) ? d2 : FLT_MAX;
        B.b2 = fminf(B.b2, fmaxf(dv, B.bd()));
    }
    if (k < B.key) { B.key = k; B.pos = j; }
}

// four candidates of run [j, e) per call; indices past the end re-read the last one (harmless)
template <bool CERT>
__device__ __forceinline__ void eval4(const MapView &m, uint32_t j, uint32_t e, float px, float py, float pz,
                                      NNBest<CERT> &B)
{
    NN_STAT(0, 1);
    SSF_CHECK(j < e && e <= m.n_pts);
    const uint32_t last = e - 1;
    const uint32_t j0 = j, j1 = min(j + 1, last), j2 = min(j + 2, last), j3 = min(j + 3, last);
    const float4 q0 = __ldg(&m.pts[j0]), q1 = __ldg(&m.pts[j1]), q2 = __ldg(&m.pts[j2]), q3 = __ldg(&m.pts[j3]);
    cand_update(B, q0, j0, true, px, py, pz);
    cand_update(B, q1, j1, j + 1 <= last, px, py, pz);
    cand_update(B, q2, j2, j + 2 <= last, px, py, pz);
    cand_update(B, q3, j3, j + 3 <= last, px, py, pz);
}

struct NNQuery {
    float px, py, pz;
    int cx, cy, cz;
    float xdn2, xup2;  // shrunk squared gaps to the cells cx - 1 / cx + 1
    float xlim2;       // shrunk squared gap to the nearer of cx - 2 / cx + 2
};

// all points of cells [xa, xb] of row (ry, rz); 0 <= xa, xb < nx, row inside the grid
template <bool CERT>
__device__ __forceinline__ void scan_cells(const MapView &m, const NNQuery &q, int xa, int xb, int ry, int rz,
                                           NNBest<CERT> &B)
{
    if (xa > xb) return;
    for (int bx = xa >> 5; bx <= (xb >> 5); ++bx) {
        NN_STAT(1, 1);
        SSF_CHECK(bx >= 0 && bx < m.nbx && ry >= 0 && ry < m.ny && rz >= 0 && rz < m.nz);
        const uint2 d = __ldg(&m.dir[dir_index(m.nbx, m.nty, bx, ry, rz)]);
        if (!d.x) continue;
        const int a = max(xa - (bx << 5), 0), b = min(xb - (bx << 5), 31);
        const uint32_t i0 = d.y + __popc(d.x & ((1u << a) - 1u));
        const uint32_t i1 = d.y + __popc(d.x & (0xFFFFFFFFu >> (31 - b)));
        SSF_CHECK(i0 <= i1 && i1 <= m.n_cells);
        if (i0 == i1) continue;
        NN_STAT(6, 1);
        uint32_t j = __ldg(&m.cell_start[i0]);
        const uint32_t e = __ldg(&m.cell_start[i1]);
        for (; j < e; j += 4) eval4(m, j, e, q.px, q.py, q.pz, B);
        B.refresh();
    }
}

// one row whose points are all at least sqrt(g) away in (y, z): keep the x cells the best
// distance still reaches.  own: cells cx-1..cx+1 of this row were already scanned.
template <bool CERT>
__device__ __forceinline__ void visit_row(const MapView &m, const NNQuery &q, int ry, int rz, float g, bool own,
                                          NNBest<CERT> &B)
{
    NN_STAT(2, 1);
    const float rem = __fsub_ru(__fmul_ru(B.prune(), kGrow), g);  // real dx^2 of any useful point is <= rem
    int xa, xb;
    if (rem < q.xlim2) {
        xa = q.cx - (rem >= q.xdn2 ? 1 : 0);
        xb = q.cx + (rem >= q.xup2 ? 1 : 0);
    } else {
        const float rx = sqrt_up(fmaxf(rem, 0.f));
        xa = cell_coord(__fsub_rd(q.px, rx), m.ox, m.inv_h, m.nx);
        xb = cell_coord(__fadd_ru(q.px, rx), m.ox, m.inv_h, m.nx);
    }
    xa = max(xa, 0);
    xb = min(xb, m.nx - 1);
    if (own) {
        scan_cells(m, q, xa, min(xb, q.cx - 2), ry, rz, B);
        scan_cells(m, q, max(xa, q.cx + 2), xb, ry, rz, B);
    } else {
        scan_cells(m, q, xa, xb, ry, rz, B);
    }
}

__device__ __forceinline__ float ring_gap(const AxisGap &a, int d, float hq)
{
    return d == 0 ? 0.f
                  : (d > 0 ? __fadd_rd(a.up, __fmul_rd((float)(d - 1), hq)) : __fadd_rd(a.dn, __fmul_rd((float)(-d - 1), hq)));
}

// rings 2, 3, ... (rows farther than the eight around the own row): rare for matched queries
template <bool CERT>
static __device__ __noinline__ void nn_far_rings(const MapView &m, const NNQuery &q, const AxisGap ay, const AxisGap az,
                                                 NNBest<CERT> &B)
{
    for (int rho = 2;; ++rho) {
        const float e = __fmul_rd((float)(rho - 1), m.hq);
        const bool y_up = q.cy + rho <= m.ny - 1, y_dn = q.cy - rho >= 0, z_up = q.cz + rho <= m.nz - 1,
                   z_dn = q.cz - rho >= 0;
        if (!(y_up || y_dn || z_up || z_dn)) return;  // the ring, and every later one, lies outside the grid
        float mn = FLT_MAX;
        if (y_up) mn = fminf(mn, __fadd_rd(ay.up, e));
        if (y_dn) mn = fminf(mn, __fadd_rd(ay.dn, e));
        if (z_up) mn = fminf(mn, __fadd_rd(az.up, e));
        if (z_dn) mn = fminf(mn, __fadd_rd(az.dn, e));
        if (B.prune() < gap_sq(mn)) return;  // later rings are farther still
        const int n_side = 2 * rho + 1;
        for (int t = 0; t < 8 * rho; ++t) {
            int dy, dz;
            if (t < 2 * n_side) {  // the two full rows of the ring: dz = -rho, +rho
                dz = t < n_side ? -rho : rho;
                dy = (t < n_side ? t : t - n_side) - rho;
            } else {  // its two sides: dy = -rho, +rho, |dz| < rho
                const int u = t - 2 * n_side;
                dy = (u & 1) ? rho : -rho;
                dz = (u >> 1) - (rho - 1);
            }
            const int ry = q.cy + dy, rz = q.cz + dz;
            if (ry < 0 || ry >= m.ny || rz < 0 || rz >= m.nz) continue;
            const float g = __fadd_rd(gap_sq(ring_gap(ay, dy, m.hq)), gap_sq(ring_gap(az, dz, m.hq)));
            if (B.prune() < g) continue;
            visit_row(m, q, ry, rz, g, false, B);
        }
    }
}

// Ring 1 in near-side-first order.  (my, mz) are multiples of the NEAR direction of each axis:
//   row 0 (1,0) | 1 (0,1) | 2 (1,1) | 3 (-1,0) | 4 (0,-1) | 5 (1,-1) | 6 (-1,1) | 7 (-1,-1)
// packed as 2-bit fields (value + 1)
constexpr uint32_t pack8(int a0, int a1, int a2, int a3, int a4, int a5, int a6, int a7)
{
    return (uint32_t)(a0 + 1) | (uint32_t)(a1 + 1) << 2 | (uint32_t)(a2 + 1) << 4 | (uint32_t)(a3 + 1) << 6 |
           (uint32_t)(a4 + 1) << 8 | (uint32_t)(a5 + 1) << 10 | (uint32_t)(a6 + 1) << 12 | (uint32_t)(a7 + 1) << 14;
}
constexpr uint32_t kRowY = pack8(1, 0, 1, -1, 0, 1, -1, -1);
constexpr uint32_t kRowZ = pack8(0, 1, 1, 0, -1, -1, 1, -1);

// limit: accept only d2 < limit (strict), like the reference's threshold test.  After the walk
// B.key < (bits(limit) << 32) iff a point was found.
// Near part: own row and ring 1.  Returns true when rings 2.. are still within reach of the bound
// (then nn_walk_far must follow, possibly later and in another thread: B carries all the state).
// seed: position (in the sorted map) of a point worth trying first -- the neighbour found by the
// previous iteration -- or kNoPos.  It only tightens the bound the walk starts with; the result is
// that of the unseeded walk.
template <bool CERT>
__device__ __forceinline__ bool nn_walk_near(const MapView &m, float px, float py, float pz, float limit, float mu,
                                             NNBest<CERT> &B, uint32_t seed = 0xFFFFFFFFu)
{
    const unsigned long long none = (unsigned long long)__float_as_uint(limit) << 32;
    B.key = none;
    B.pos = 0;
    B.skip = 0xFFFFFFFFu;
    B.b2 = FLT_MAX;
    B.mu = mu;
    B.bdm = limit;
    if (seed != 0xFFFFFFFFu) {
        SSF_CHECK(seed < m.n_pts);
        cand_update(B, __ldg(&m.pts[seed]), seed, true, px, py, pz);
        B.skip = seed;
    }
    const bool seeded = B.key < none;
    B.refresh();
    NN_STAT(3, 1);
    {   // farther from the map's bounding box than the (inflated) limit: nothing to look at
        const float ex = fmaxf(fmaxf(__fsub_rd(m.bmin[0], px), __fsub_rd(px, m.bmax[0])), 0.f);
        const float ey = fmaxf(fmaxf(__fsub_rd(m.bmin[1], py), __fsub_rd(py, m.bmax[1])), 0.f);
        const float ez = fmaxf(fmaxf(__fsub_rd(m.bmin[2], pz), __fsub_rd(pz, m.bmax[2])), 0.f);
        const float out2 = __fmul_rd(__fadd_rd(__fadd_rd(__fmul_rd(ex, ex), __fmul_rd(ey, ey)), __fmul_rd(ez, ez)), kShrink);
        if (B.prune() < out2) return false;
    }
    const AxisGap ax = axis_gap(px, m.ox, m.inv_h, m.hq, m.nx), ay = axis_gap(py, m.oy, m.inv_h, m.hq, m.ny),
                  az = axis_gap(pz, m.oz, m.inv_h, m.hq, m.nz);
    NNQuery q;
    q.px = px; q.py = py; q.pz = pz;
    q.cx = ax.c; q.cy = ay.c; q.cz = az.c;
    q.xdn2 = gap_sq(ax.dn);
    q.xup2 = gap_sq(ax.up);
    q.xlim2 = gap_sq(__fadd_rd(fminf(ax.dn, ax.up), m.hq));
    // reach mask: a clear bit proves that nothing lies within sqrt(reach2) of this cell
    if (!seeded && m.reach != nullptr && B.prune() < m.reach2 && q.cx >= 0 && q.cx < m.nx && q.cy >= 0 && q.cy < m.ny &&
        q.cz >= 0 && q.cz < m.nz) {
        const uint32_t w = __ldg(&m.reach[dir_index(m.nbx, m.nty, q.cx >> 5, q.cy, q.cz)]);
        if (!((w >> (q.cx & 31)) & 1u)) {
            NN_STAT(7, 1);
            return false;
        }
    }
    // own row: with a seed, just the cells its distance reaches; else start with the cells
    // cx-1..cx+1, then whatever else of the row is still in reach
    if (q.cy >= 0 && q.cy < m.ny && q.cz >= 0 && q.cz < m.nz) {
        if (seeded) {
            visit_row(m, q, q.cy, q.cz, 0.f, false, B);
        } else {
            scan_cells(m, q, max(q.cx - 1, 0), min(q.cx + 1, m.nx - 1), q.cy, q.cz, B);
            if (!(__fmul_ru(B.prune(), kGrow) < q.xlim2)) visit_row(m, q, q.cy, q.cz, 0.f, true, B);
        }
    }
    // ring 1: which of the eight rows can still hold a better point?
    const int sy = ay.up <= ay.dn ? 1 : -1, sz = az.up <= az.dn ? 1 : -1;
    const float yn2 = gap_sq(fminf(ay.up, ay.dn)), yf2 = gap_sq(fmaxf(ay.up, ay.dn));
    const float zn2 = gap_sq(fminf(az.up, az.dn)), zf2 = gap_sq(fmaxf(az.up, az.dn));
    uint32_t mask = 0;
    {
        const float bd = B.prune();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int my = (int)((kRowY >> (2 * k)) & 3u) - 1, mz = (int)((kRowZ >> (2 * k)) & 3u) - 1;
            const int ry = q.cy + sy * my, rz = q.cz + sz * mz;
            const float g = __fadd_rd(my > 0 ? yn2 : (my < 0 ? yf2 : 0.f), mz > 0 ? zn2 : (mz < 0 ? zf2 : 0.f));
            const bool out = ry < 0 || ry >= m.ny || rz < 0 || rz >= m.nz || bd < g;
            mask |= out ? 0u : (1u << k);
        }
    }
    while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        const int my = (int)((kRowY >> (2 * k)) & 3u) - 1, mz = (int)((kRowZ >> (2 * k)) & 3u) - 1;
        const float g = __fadd_rd(my > 0 ? yn2 : (my < 0 ? yf2 : 0.f), mz > 0 ? zn2 : (mz < 0 ? zf2 : 0.f));
        if (B.prune() < g) continue;
        visit_row(m, q, q.cy + sy * my, q.cz + sz * mz, g, false, B);
    }
    // farther rings only while the bound reaches past ring 1
    const float reach = fminf(fminf(__fadd_rd(ay.up, m.hq), __fadd_rd(ay.dn, m.hq)),
                              fminf(__fadd_rd(az.up, m.hq), __fadd_rd(az.dn, m.hq)));
    if (!(B.prune() < gap_sq(reach))) {
        NN_STAT(4, 1);
        return true;
    }
    return false;
}

// ---- the far rings as ONE loop -----------------------------------------------------------------------
// nn_far_rings nests three loops (rows of a ring, blocks of a row, candidates of a run).  The flat form
// below visits exactly the same rows, runs and candidates in exactly the same order (same bound at every
// decision, same refresh points, hence bit-identical results AND certificates), but as a single loop:
// every iteration, a lane that has no candidates at hand takes one step of its walk (next block of the
// current cell range, or the next row the bound does not rule out), and every lane that has candidates
// evaluates four of them.  Inlined into search_accum_kernel (no call, no copy of the map view to the
// stack): converged launches 144 -> 136 us.  The same treatment of the NEAR part was measured slower
// (first launch 1.46 -> 1.69 ms: the stage logic is issued every iteration for the few lanes that need
// it) and is not kept.
struct WalkCursor {
    int bx, bx_end;  // blocks of the current cell range still to look at (bx > bx_end: none)
    int xa, xb;      // the cell range
    int ry, rz;      // its row
    uint32_t j, e;   // candidates at hand: [j, e) of the sorted cloud
};

__device__ __forceinline__ void cursor_cells(WalkCursor &c, int xa, int xb, int ry, int rz)
{
    c.xa = xa; c.xb = xb; c.ry = ry; c.rz = rz;
    c.bx = 1; c.bx_end = 0;
    if (xa <= xb) { c.bx = xa >> 5; c.bx_end = xb >> 5; }
}

// one block of the cursor's cell range: its run, if any, becomes the candidates at hand
__device__ __forceinline__ void cursor_block(const MapView &m, WalkCursor &c)
{
    NN_STAT(1, 1);
    SSF_CHECK(c.bx >= 0 && c.bx < m.nbx && c.ry >= 0 && c.ry < m.ny && c.rz >= 0 && c.rz < m.nz);
    const uint2 d = __ldg(&m.dir[dir_index(m.nbx, m.nty, c.bx, c.ry, c.rz)]);
    if (d.x) {
        const int a = max(c.xa - (c.bx << 5), 0), b = min(c.xb - (c.bx << 5), 31);
        const uint32_t i0 = d.y + __popc(d.x & ((1u << a) - 1u));
        const uint32_t i1 = d.y + __popc(d.x & (0xFFFFFFFFu >> (31 - b)));
        SSF_CHECK(i0 <= i1 && i1 <= m.n_cells);
        if (i0 != i1) {
            NN_STAT(6, 1);
            c.j = __ldg(&m.cell_start[i0]);
            c.e = __ldg(&m.cell_start[i1]);
        }
    }
    ++c.bx;
}

// cells of a row at squared (y, z) gap g that the bound still reaches (the range part of visit_row)
template <bool CERT>
__device__ __forceinline__ void reach_cells(const MapView &m, const NNQuery &q, float g, const NNBest<CERT> &B, int &xa, int &xb)
{
    NN_STAT(2, 1);
    const float rem = __fsub_ru(__fmul_ru(B.prune(), kGrow), g);  // real dx^2 of any useful point is <= rem
    if (rem < q.xlim2) {
        xa = q.cx - (rem >= q.xdn2 ? 1 : 0);
        xb = q.cx + (rem >= q.xup2 ? 1 : 0);
    } else {
        const float rx = sqrt_up(fmaxf(rem, 0.f));
        xa = cell_coord(__fsub_rd(q.px, rx), m.ox, m.inv_h, m.nx);
        xb = cell_coord(__fadd_ru(q.px, rx), m.ox, m.inv_h, m.nx);
    }
    xa = max(xa, 0);
    xb = min(xb, m.nx - 1);
}

// nn_walk_far as one loop (rings 2, 3, ... until the bound is met; same order, same result)
template <bool CERT>
__device__ __forceinline__ void nn_walk_far_flat(const MapView &m, float px, float py, float pz, NNBest<CERT> &B)
{
    const AxisGap ax = axis_gap(px, m.ox, m.inv_h, m.hq, m.nx), ay = axis_gap(py, m.oy, m.inv_h, m.hq, m.ny),
                  az = axis_gap(pz, m.oz, m.inv_h, m.hq, m.nz);
    NNQuery q;
    q.px = px; q.py = py; q.pz = pz;
    q.cx = ax.c; q.cy = ay.c; q.cz = az.c;
    q.xdn2 = gap_sq(ax.dn);
    q.xup2 = gap_sq(ax.up);
    q.xlim2 = gap_sq(__fadd_rd(fminf(ax.dn, ax.up), m.hq));
    WalkCursor c;
    c.j = c.e = 0;
    c.bx = 1; c.bx_end = 0;
    int rho = 1, t = 0, n_t = 0;  // the first step opens ring 2
    for (;;) {
        if (c.j >= c.e) {
            if (c.bx > c.bx_end) {  // the next row of the rings that the bound does not rule out
                for (;;) {
                    if (t >= n_t) {  // next ring
                        ++rho;
                        const float e = __fmul_rd((float)(rho - 1), m.hq);
                        const bool y_up = q.cy + rho <= m.ny - 1, y_dn = q.cy - rho >= 0, z_up = q.cz + rho <= m.nz - 1,
                                   z_dn = q.cz - rho >= 0;
                        if (!(y_up || y_dn || z_up || z_dn)) return;  // the ring, and every later one, lies outside the grid
                        float mn = FLT_MAX;
                        if (y_up) mn = fminf(mn, __fadd_rd(ay.up, e));
                        if (y_dn) mn = fminf(mn, __fadd_rd(ay.dn, e));
                        if (z_up) mn = fminf(mn, __fadd_rd(az.up, e));
                        if (z_dn) mn = fminf(mn, __fadd_rd(az.dn, e));
                        if (B.prune() < gap_sq(mn)) return;  // later rings are farther still
                        n_t = 8 * rho;
                        t = 0;
                    }
                    const int n_side = 2 * rho + 1;
                    int dy, dz;
                    if (t < 2 * n_side) {  // the two full rows of the ring: dz = -rho, +rho
                        dz = t < n_side ? -rho : rho;
                        dy = (t < n_side ? t : t - n_side) - rho;
                    } else {  // its two sides: dy = -rho, +rho, |dz| < rho
                        const int u = t - 2 * n_side;
                        dy = (u & 1) ? rho : -rho;
                        dz = (u >> 1) - (rho - 1);
                    }
                    ++t;
                    const int ry = q.cy + dy, rz = q.cz + dz;
                    if (ry < 0 || ry >= m.ny || rz < 0 || rz >= m.nz) continue;
                    const float g = __fadd_rd(gap_sq(ring_gap(ay, dy, m.hq)), gap_sq(ring_gap(az, dz, m.hq)));
                    if (B.prune() < g) continue;
                    int xa, xb;
                    reach_cells(m, q, g, B, xa, xb);
                    cursor_cells(c, xa, xb, ry, rz);
                    break;
                }
            }
            if (c.bx <= c.bx_end) cursor_block(m, c);
        }
        if (c.j < c.e) {
            eval4(m, c.j, c.e, px, py, pz, B);
            c.j += 4;
            if (c.j >= c.e) B.refresh();
        }
    }
}

// Far part: rings 2, 3, ... until the bound is met.
template <bool CERT>
__device__ __forceinline__ void nn_walk_far(const MapView &m, float px, float py, float pz, NNBest<CERT> &B)
{
    const AxisGap ax = axis_gap(px, m.ox, m.inv_h, m.hq, m.nx), ay = axis_gap(py, m.oy, m.inv_h, m.hq, m.ny),
                  az = axis_gap(pz, m.oz, m.inv_h, m.hq, m.nz);
    NNQuery q;
    q.px = px; q.py = py; q.pz = pz;
    q.cx = ax.c; q.cy = ay.c; q.cz = az.c;
    q.xdn2 = gap_sq(ax.dn);
    q.xup2 = gap_sq(ax.up);
    q.xlim2 = gap_sq(__fadd_rd(fminf(ax.dn, ax.up), m.hq));
    nn_far_rings(m, q, ay, az, B);
}

// The complete walk in one thread.
template <bool CERT>
__device__ __forceinline__ void nn_walk(const MapView &m, float px, float py, float pz, float limit, float mu,
                                        NNBest<CERT> &B)
{
    if (nn_walk_near(m, px, py, pz, limit, mu, B)) nn_walk_far(m, px, py, pz, B);
}

__device__ __forceinline__ NNHit nn_query(const MapView &m, float px, float py, float pz, float limit)
{
    NNHit h;
    h.d2 = limit;
    h.idx = -1;
    h.pos = 0;
    if (!(limit > 0.f) || m.n_pts == 0 || !isfinite(px) || !isfinite(py) || !isfinite(pz)) return h;
    NNBest<false> B;
    nn_walk<false>(m, px, py, pz, limit, 0.f, B);
    if ((uint32_t)(B.key >> 32) < __float_as_uint(limit)) {
        NN_STAT(5, 1);
        h.d2 = B.bd();
        h.idx = (int)(uint32_t)(B.key & 0xFFFFFFFFull);
        h.pos = B.pos;
    }
    return h;
}

// ---- certificates: skipping the search when the query has barely moved ---------------------------
// A CERT walk at query position p leaves, besides the result, a radius L such that EVERY map point
// other than the reported neighbour lies farther than L from p (real distances):
//   * evaluated points: their float d2 >= b2, i.e. real distance >= sqrt(b2 * (1 - 4e-7));
//   * points never looked at: farther than sqrt(d2 best) + mu (see NNBest).
// If the same source point is searched again at p' (next iteration, pose moved a little) and
//       d2(p', neighbour) < limit   and   d2(p', neighbour) < (L - |p' - p|)^2 * (1 - 2e-6),
// the neighbour is still the strict minimum, so the exact search would return the same index and
// the d2 evaluated here -- bit for bit -- and is skipped.  For a query without a neighbour the
// same L bounds ALL points: it stays unmatched while (L - |p' - p|)^2 * (1 - 2e-6) >= limit.
constexpr uint32_t kNoPos = 0xFFFFFFFFu;
// A certificate is 8 bytes per query: (radius L with the low 5 mantissa bits replaced by the
// iteration that issued it, neighbour position).  The query position of that iteration is
// recomputed from the scan's pose history (ScanState poses of the first kCertHist iterations).
constexpr int kCertHist = 32;
__device__ __forceinline__ uint32_t cert_pack(float L, int iteration)
{
    return (__float_as_uint(fmaxf(L, 0.f)) & ~31u) | ((uint32_t)iteration & 31u);  // truncation only shrinks L
}

// L of a finished CERT walk (rounded down)
__device__ __forceinline__ float cert_radius(const NNBest<true> &B)
{
    const float s2 = __fmul_rd(sqrt_dn(__fmul_rd(B.b2, 0.999999f)), 0.999999f);
    const float s1 = __fadd_rd(sqrt_dn(B.bd()), B.mu);  // bd() == limit when nothing was found
    return fminf(s1, s2);
}

// (cx, cy, cz) = where the query was when the certificate of radius L was issued, pos = neighbour
// position or kNoPos.  Returns true when the old result stands; then key = packed (d2, index) of
// the neighbour at the new position (or the none sentinel).
__device__ __forceinline__ bool nn_verify(const MapView &m, float cx, float cy, float cz, float L, uint32_t pos, float px,
                                          float py, float pz, float limit, unsigned long long &key)
{
    const float mx = __fsub_rn(px, cx), my = __fsub_rn(py, cy), mz = __fsub_rn(pz, cz);
    const float mv2 = __fadd_ru(__fadd_ru(__fmul_ru(mx, mx), __fmul_ru(my, my)), __fmul_ru(mz, mz));
    const float mv = __fmul_ru(sqrt_up(mv2), 1.000001f);  // >= |p' - p|
    const float rest = __fsub_rd(L, mv);
    if (!(rest > 0.f)) return false;
    const float thr = __fmul_rd(__fmul_rd(rest, rest), kShrink);  // every OTHER point has float d2 > thr
    const unsigned long long none = (unsigned long long)__float_as_uint(limit) << 32;
    if (pos == kNoPos) {
        key = none;
        return thr >= limit;
    }
    SSF_CHECK(pos < m.n_pts);
    const float4 q = __ldg(&m.pts[pos]);
    const float dx = __fsub_rn(px, q.x), dy = __fsub_rn(py, q.y), dz = __fsub_rn(pz, q.z);
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)__float_as_uint(q.w);
    return d2 < limit && d2 < thr;
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

// 16 doubles per thread: after the call lane l holds the warp-wide sum of v[l & 15]
__device__ __forceinline__ double warp_transpose_reduce16(double (&v)[16])
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int half = 8; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const double send = up ? v[i] : v[i + half];
            const double keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

}  // namespace ssf
