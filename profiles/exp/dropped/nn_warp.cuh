// ARCHIVED EXPERIMENT (round 2) -- not part of the build.  Warp-cooperative candidate evaluation for the
// fused search kernel: parity-green (70 / 70 GPU tests), but slower than the one-thread-per-query walk it was
// meant to replace (first launches of a 256-scan step 1.85 / 1.26 / 1.22 / 0.60 ms against 1.46 / 0.90 / 0.58 /
// 0.42 ms; 1152 M warp instructions in the first launch against 929 M).  See DESIGN.md section 4 and
// profiles/r2/SUMMARY.md.  It was included from icp_kernels.cu in place of nn_device.cuh, with
//     WarpWalk<CERT> W(map, s_wq[warp], {s_q, s_key, s_b2, s_skip, s_pos}, mu, none_hi, lane);
//     far = warp_walk_near<CERT>(W, has, r, limit);  ...  warp_walk_far<CERT>(W, has, r);
// in the S phase (one __shared__ WarpQueue per warp).
//
// nn_warp.cuh -- the exact walk of nn_device.cuh, one query per lane, with the CANDIDATE evaluation
// shared by the warp (device side of K3 inside search_accum_kernel).
//
// Replaces kdtree_.nearestKSearch(p, 1, idx, d2) + the `d2 < max_correspondence_dist_` test of
// reference localization/src/icp_point_to_point.cpp:64-70, exactly like nn_device.cuh (same result,
// bit for bit: smallest float d2, ties to the lowest original index).
//
// Why.  In the one-thread-per-query walk half of all issued instructions were the candidate loop, at
// 9 of 32 lanes: the lanes of a warp reach runs of different length at different times.  Here a lane
// still walks its own rows (directory load, pruning, x range -- cheap, per-query geometry), but the
// runs it finds are not evaluated by that lane.  They are cut into GROUPS of four consecutive points
// and appended to a ring of pending groups owned by the warp; whenever 32 groups are pending, every
// lane takes one group -- whichever query it belongs to -- evaluates its four candidates against that
// query and merges the result into the query's running best in shared memory.  The candidate work is
// thereby spread evenly over the lanes regardless of which lane found it.
//
// Running best of a query (shared memory, indexed by the query's slot r in the tile):
//   key[r]  (d2 bits << 32) | original index, merged with a 64-bit atomic min
//   pos[r]  position of that point in the sorted cloud: after the merges of a batch are done
//           (__syncwarp) the lane whose group minimum IS the key writes it -- keys are unique per point
//   b2[r]   CERT walks: second-smallest d2 among everything evaluated.  Each group contributes the
//           second-smallest of its own candidates and max(d2 of its minimum, d2 of the key it met in
//           the atomic); the smallest such value over all groups is the second-smallest overall
//           (whichever of the two smallest group minima is merged later meets the other one).
//   skip[r] CERT walks: position of the seed (evaluated up front by the owner, not counted twice)
// Pruning reads key[r] whenever a lane looks at its next row; a stale (larger) bound only means a few
// more candidates -- every bound ever used is the distance of a real point, so nothing nearer than the
// final result is skipped, and equal distances are never pruned (strict comparisons, as before).
#pragma once
#include "nn_device.cuh"

namespace ssf {

constexpr uint32_t kEntCap = 256;    // pending groups per warp (ring buffer, power of two)
constexpr uint32_t kEntPerStep = 4;  // groups one lane may append per step (16 candidates)
constexpr uint32_t kFull = 0xffffffffu;

struct WarpQueue {
    uint32_t ent_j[kEntCap];         // first candidate of the group (position in the sorted cloud)
    unsigned char ent_o[kEntCap];    // owner lane | (candidates - 1) << 5
    unsigned short r_of[32];         // tile slot of the query each lane owns
};

struct WalkArrays {
    const float4 *q;          // transformed queries of the tile
    unsigned long long *key;  // running best per query
    uint32_t *b2;             // bits of the second-smallest d2 (CERT)
    uint32_t *skip;           // seed position (CERT)
    uint32_t *pos;            // in: seed position or kNoPos; out: position of the best
};

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t lane, uint32_t &total)
{
    uint32_t s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, s, d);
        if (lane >= (uint32_t)d) s += t;
    }
    total = __shfl_sync(kFull, s, 31);
    return s - v;
}

template <bool CERT>
struct WarpWalk {
    const MapView &m;
    WarpQueue &wq;
    const WalkArrays a;
    const uint32_t lane;
    const float mu;
    const unsigned long long none;  // key of "nothing found": (bits of the limit) << 32
    uint32_t head, tail;            // warp-uniform

    // none_hi: bit pattern of the rejection threshold (squared distance)
    __device__ __forceinline__ WarpWalk(const MapView &m_, WarpQueue &wq_, const WalkArrays &a_, float mu_, uint32_t none_hi,
                                        uint32_t lane_)
        : m(m_), wq(wq_), a(a_), lane(lane_), mu(mu_), none((unsigned long long)none_hi << 32), head(0), tail(0)
    {
    }

    // the bound to prune with, from the running best of query r
    __device__ __forceinline__ float bound(uint32_t r) const
    {
        const float bd = __uint_as_float((uint32_t)(a.key[r] >> 32));
        if (!CERT) return bd;
        const float s = __fadd_ru(sqrt_up(bd), mu);  // (sqrt(bd) + mu)^2, everything rounded up
        return __fmul_ru(s, s);
    }

    // `count` (<= 32, warp-uniform) pending groups, one per lane
    __device__ __forceinline__ void eval(uint32_t count)
    {
        unsigned long long k1 = ~0ull;
        uint32_t p1 = 0, r = 0;
        if (lane < count) {
            NN_STAT(0, 1);
            const uint32_t slot = (head + lane) & (kEntCap - 1);
            const uint32_t j = wq.ent_j[slot], o = wq.ent_o[slot];
            const uint32_t last = j + (o >> 5);
            r = wq.r_of[o & 31u];
            const float4 p = a.q[r];
            SSF_CHECK(last < m.n_pts);
            const uint32_t j1 = min(j + 1, last), j2 = min(j + 2, last), j3 = min(j + 3, last);
            const float4 c0 = __ldg(&m.pts[j]), c1 = __ldg(&m.pts[j1]), c2 = __ldg(&m.pts[j2]), c3 = __ldg(&m.pts[j3]);
            float d[4];
            {
                const float4 c[4] = {c0, c1, c2, c3};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float dx = __fsub_rn(p.x, c[i].x), dy = __fsub_rn(p.y, c[i].y), dz = __fsub_rn(p.z, c[i].z);
                    d[i] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                }
            }
            const float mn = fminf(fminf(d[0], d[1]), fminf(d[2], d[3]));
            const uint32_t bd_bits = (uint32_t)(a.key[r] >> 32);  // possibly stale: never below the current one
#ifdef SSF_DBG_NOFAST
            if (false) {
#else
            if (__float_as_uint(mn) > bd_bits) {
#endif
                // none of the four can be the result (d2 >= 0: the bit patterns order like the floats)
                if (CERT) atomicMin(&a.b2[r], __float_as_uint(mn));
            } else {
                // minimum and (CERT) second-smallest of the group, candidates in order
                const uint32_t w[4] = {__float_as_uint(c0.w), __float_as_uint(c1.w), __float_as_uint(c2.w),
                                       __float_as_uint(c3.w)};
                const uint32_t jj[4] = {j, j1, j2, j3};
                const uint32_t skip = CERT ? a.skip[r] : kNoPos;
                unsigned long long lk = ((unsigned long long)__float_as_uint(d[0]) << 32) | w[0];
                int arg = 0;
                p1 = j;
#pragma unroll
                for (int i = 1; i < 4; ++i) {  // (padding repeats the last point: an equal key never replaces)
                    const unsigned long long k = ((unsigned long long)__float_as_uint(d[i]) << 32) | w[i];
                    if (k < lk) { lk = k; arg = i; p1 = jj[i]; }
                }
                float m2 = FLT_MAX;
                if (CERT) {  // the group's other points; padding and the seed are not second points
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const bool valid = i != arg && j + (uint32_t)i <= last && jj[i] != skip;
                        m2 = fminf(m2, valid ? d[i] : FLT_MAX);
                    }
                }
                k1 = lk;
                const unsigned long long old = atomicMin(&a.key[r], k1);
                if (CERT) {
                    // the same point twice (only the seed can be, and a seed is below `none`): it is not its
                    // own runner-up
                    const float other = (old == k1 && k1 < none) ? FLT_MAX : fmaxf(__uint_as_float((uint32_t)(k1 >> 32)),
                                                                     __uint_as_float((uint32_t)(old >> 32)));
                    atomicMin(&a.b2[r], __float_as_uint(fminf(m2, other)));
                }
            }
        }
        __syncwarp();
        if (k1 != ~0ull && a.key[r] == k1) a.pos[r] = p1;
        __syncwarp();
    }

    __device__ __forceinline__ void drain()
    {
        while (tail != head) {
            const uint32_t n = min(tail - head, 32u);
            eval(n);
            head += n;
        }
    }

    // Collective: every lane names the cells [xa, xb] of row (ry, rz) of ITS query (xa > xb: nothing).
    // The points of those cells are appended to the ring; full batches of 32 groups are evaluated.
    __device__ __forceinline__ void emit_cells(int xa, int xb, int ry, int rz)
    {
        int bx = 1, bx_end = 0;
        if (xa <= xb) { bx = xa >> 5; bx_end = xb >> 5; }
        uint32_t j = 0, e = 0;
        while (true) {
            while (j >= e && bx <= bx_end) {  // this lane's next non-empty run
                NN_STAT(1, 1);
                SSF_CHECK(bx >= 0 && bx < m.nbx && ry >= 0 && ry < m.ny && rz >= 0 && rz < m.nz);
                const uint2 d = __ldg(&m.dir[dir_index(m.nbx, m.nty, bx, ry, rz)]);
                if (d.x) {
                    const int lo = max(xa - (bx << 5), 0), hi = min(xb - (bx << 5), 31);
                    const uint32_t i0 = d.y + __popc(d.x & ((1u << lo) - 1u));
                    const uint32_t i1 = d.y + __popc(d.x & (0xFFFFFFFFu >> (31 - hi)));
                    SSF_CHECK(i0 <= i1 && i1 <= m.n_cells);
                    if (i0 != i1) {
                        NN_STAT(6, 1);
                        j = __ldg(&m.cell_start[i0]);
                        e = __ldg(&m.cell_start[i1]);
                    }
                }
                ++bx;
            }
            const bool has = j < e;
            if (!__any_sync(kFull, has)) break;
            const uint32_t n4 = has ? min((e - j + 3u) >> 2, kEntPerStep) : 0u;
            uint32_t total;
            const uint32_t off = tail + warp_excl_scan(n4, lane, total);
#pragma unroll
            for (uint32_t t = 0; t < kEntPerStep; ++t) {
                if (t < n4) {
                    const uint32_t slot = (off + t) & (kEntCap - 1);
                    const uint32_t jt = j + 4u * t;
                    wq.ent_j[slot] = jt;
                    wq.ent_o[slot] = (unsigned char)(lane | ((min(e - jt, 4u) - 1u) << 5));
                }
            }
            j += 4u * n4;
            tail += total;
            __syncwarp();
            while (tail - head >= 32u) {
                eval(32u);
                head += 32u;
            }
        }
    }
};

// x cells of a row at squared (y, z) gap g that the bound still reaches (visit_row of nn_device.cuh)
__device__ __forceinline__ void row_range(const MapView &m, const NNQuery &q, float g, float bound, int &xa, int &xb)
{
    const float rem = __fsub_ru(__fmul_ru(bound, kGrow), g);  // real dx^2 of any useful point is <= rem
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

__device__ __forceinline__ NNQuery make_query(const MapView &m, float px, float py, float pz, AxisGap &ay, AxisGap &az)
{
    const AxisGap ax = axis_gap(px, m.ox, m.inv_h, m.hq, m.nx);
    ay = axis_gap(py, m.oy, m.inv_h, m.hq, m.ny);
    az = axis_gap(pz, m.oz, m.inv_h, m.hq, m.nz);
    NNQuery q;
    q.px = px; q.py = py; q.pz = pz;
    q.cx = ax.c; q.cy = ay.c; q.cz = az.c;
    q.xdn2 = gap_sq(ax.dn);
    q.xup2 = gap_sq(ax.up);
    q.xlim2 = gap_sq(__fadd_rd(fminf(ax.dn, ax.up), m.hq));
    return q;
}

// Near part (own row and ring 1) for the queries of the warp's lanes: lane with has == true owns the
// query in tile slot r.  Collective -- every lane of the warp must call it.  Returns true when rings
// 2.. are still within reach of that query's bound (warp_walk_far must follow).  On return key[r],
// pos[r] (and b2[r], skip[r] for CERT) hold the state of the walk.
template <bool CERT>
__device__ __forceinline__ bool warp_walk_near(WarpWalk<CERT> &W, bool has, uint32_t r, float limit)
{
    const MapView &m = W.m;
    const unsigned long long none = W.none;
    bool alive = has, seeded = false;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (has) {
        const float4 p = W.a.q[r];
        px = p.x; py = p.y; pz = p.z;
        const uint32_t seed = W.a.pos[r];
        unsigned long long key = none;
        float b2 = FLT_MAX;
        if (seed != kNoPos) {  // the previous neighbour first: it only tightens the bound the walk starts with
            SSF_CHECK(seed < m.n_pts);
            const float4 c = __ldg(&m.pts[seed]);
            const float dx = __fsub_rn(px, c.x), dy = __fsub_rn(py, c.y), dz = __fsub_rn(pz, c.z);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const unsigned long long k = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)__float_as_uint(c.w);
            b2 = fmaxf(d2, limit);
            if (k < key) key = k;
        }
        seeded = key < none;
        W.a.key[r] = key;
        if (CERT) {
            W.a.b2[r] = __float_as_uint(b2);
            W.a.skip[r] = seed;
        }
        W.wq.r_of[W.lane] = (unsigned short)r;
        NN_STAT(3, 1);
    }
    __syncwarp();
    float bd = has ? W.bound(r) : 0.f;
    if (alive) {  // farther from the map's bounding box than the (inflated) limit: nothing to look at
        const float ex = fmaxf(fmaxf(__fsub_rd(m.bmin[0], px), __fsub_rd(px, m.bmax[0])), 0.f);
        const float ey = fmaxf(fmaxf(__fsub_rd(m.bmin[1], py), __fsub_rd(py, m.bmax[1])), 0.f);
        const float ez = fmaxf(fmaxf(__fsub_rd(m.bmin[2], pz), __fsub_rd(pz, m.bmax[2])), 0.f);
        const float out2 = __fmul_rd(__fadd_rd(__fadd_rd(__fmul_rd(ex, ex), __fmul_rd(ey, ey)), __fmul_rd(ez, ez)), kShrink);
        if (bd < out2) alive = false;
    }
    AxisGap ay, az;
    const NNQuery q = make_query(m, px, py, pz, ay, az);
    // reach mask: a clear bit proves that nothing lies within sqrt(reach2) of this cell
    if (alive && !seeded && m.reach != nullptr && bd < m.reach2 && q.cx >= 0 && q.cx < m.nx && q.cy >= 0 && q.cy < m.ny &&
        q.cz >= 0 && q.cz < m.nz) {
        const uint32_t w = __ldg(&m.reach[dir_index(m.nbx, m.nty, q.cx >> 5, q.cy, q.cz)]);
        if (!((w >> (q.cx & 31)) & 1u)) {
            NN_STAT(7, 1);
            alive = false;
        }
    }
    // own row: with a seed, just the cells its distance reaches; else the cells cx-1..cx+1 first,
    // then whatever else of the row is still in reach
    const bool row_in = alive && q.cy >= 0 && q.cy < m.ny && q.cz >= 0 && q.cz < m.nz;
    {
        int xa = 1, xb = 0;
        if (row_in) {
            if (seeded) {
                NN_STAT(2, 1);
                row_range(m, q, 0.f, bd, xa, xb);
            } else {
                xa = max(q.cx - 1, 0);
                xb = min(q.cx + 1, m.nx - 1);
            }
        }
        W.emit_cells(xa, xb, q.cy, q.cz);
        W.drain();
    }
    if (has) bd = W.bound(r);
    {
        const bool rest = row_in && !seeded && !(__fmul_ru(bd, kGrow) < q.xlim2);
        if (__any_sync(kFull, rest)) {
            int xa = 1, xb = 0;
            if (rest) {
                NN_STAT(2, 1);
                row_range(m, q, 0.f, bd, xa, xb);
            }
            W.emit_cells(xa, rest ? min(xb, q.cx - 2) : 0, q.cy, q.cz);
            W.emit_cells(rest ? max(xa, q.cx + 2) : 1, xb, q.cy, q.cz);
            W.drain();
            if (has) bd = W.bound(r);
        }
    }
    // ring 1: which of the eight rows can still hold a better point?
    const int sy = ay.up <= ay.dn ? 1 : -1, sz = az.up <= az.dn ? 1 : -1;
    const float yn2 = gap_sq(fminf(ay.up, ay.dn)), yf2 = gap_sq(fmaxf(ay.up, ay.dn));
    const float zn2 = gap_sq(fminf(az.up, az.dn)), zf2 = gap_sq(fmaxf(az.up, az.dn));
    uint32_t mask = 0;
    if (alive) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int my = (int)((kRowY >> (2 * k)) & 3u) - 1, mz = (int)((kRowZ >> (2 * k)) & 3u) - 1;
            const int ry = q.cy + sy * my, rz = q.cz + sz * mz;
            const float g = __fadd_rd(my > 0 ? yn2 : (my < 0 ? yf2 : 0.f), mz > 0 ? zn2 : (mz < 0 ? zf2 : 0.f));
            const bool out = ry < 0 || ry >= m.ny || rz < 0 || rz >= m.nz || bd < g;
            mask |= out ? 0u : (1u << k);
        }
    }
    while (__any_sync(kFull, mask != 0u)) {
        int xa = 1, xb = 0, ry = 0, rz = 0;
        if (mask) bd = W.bound(r);
        while (mask) {  // this lane's next row that its bound does not rule out
            const int k = __ffs(mask) - 1;
            mask &= mask - 1;
            const int my = (int)((kRowY >> (2 * k)) & 3u) - 1, mz = (int)((kRowZ >> (2 * k)) & 3u) - 1;
            const float g = __fadd_rd(my > 0 ? yn2 : (my < 0 ? yf2 : 0.f), mz > 0 ? zn2 : (mz < 0 ? zf2 : 0.f));
            if (bd < g) continue;
            NN_STAT(2, 1);
            row_range(m, q, g, bd, xa, xb);
            ry = q.cy + sy * my;
            rz = q.cz + sz * mz;
            break;
        }
        W.emit_cells(xa, xb, ry, rz);
    }
    W.drain();
    if (!alive) return false;
    bd = W.bound(r);
    // farther rings only while the bound reaches past ring 1
    const float reach = fminf(fminf(__fadd_rd(ay.up, m.hq), __fadd_rd(ay.dn, m.hq)),
                              fminf(__fadd_rd(az.up, m.hq), __fadd_rd(az.dn, m.hq)));
    if (!(bd < gap_sq(reach))) {
        NN_STAT(4, 1);
        return true;
    }
    return false;
}

// Far part: rings 2, 3, ... until every lane's bound is met.  Collective; key[r], pos[r] (b2[r],
// skip[r]) carry the state over from warp_walk_near, possibly run by another warp.
template <bool CERT>
__device__ __forceinline__ void warp_walk_far(WarpWalk<CERT> &W, bool has, uint32_t r)
{
    const MapView &m = W.m;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (has) {
        const float4 p = W.a.q[r];
        px = p.x; py = p.y; pz = p.z;
        W.wq.r_of[W.lane] = (unsigned short)r;
    }
    __syncwarp();
    AxisGap ay, az;
    const NNQuery q = make_query(m, px, py, pz, ay, az);
    bool done = !has;
    for (int rho = 2;; ++rho) {
        if (!done) {
            const float e = __fmul_rd((float)(rho - 1), m.hq);
            const bool y_up = q.cy + rho <= m.ny - 1, y_dn = q.cy - rho >= 0, z_up = q.cz + rho <= m.nz - 1,
                       z_dn = q.cz - rho >= 0;
            if (!(y_up || y_dn || z_up || z_dn)) {
                done = true;  // the ring, and every later one, lies outside the grid
            } else {
                float mn = FLT_MAX;
                if (y_up) mn = fminf(mn, __fadd_rd(ay.up, e));
                if (y_dn) mn = fminf(mn, __fadd_rd(ay.dn, e));
                if (z_up) mn = fminf(mn, __fadd_rd(az.up, e));
                if (z_dn) mn = fminf(mn, __fadd_rd(az.dn, e));
                if (W.bound(r) < gap_sq(mn)) done = true;  // later rings are farther still
            }
        }
        if (__all_sync(kFull, done)) return;
        const int n_side = 2 * rho + 1, n_t = 8 * rho;
        int t = done ? n_t : 0;
        while (__any_sync(kFull, t < n_t)) {
            int xa = 1, xb = 0, ry = 0, rz = 0;
            if (t < n_t) {
                const float bd = W.bound(r);
                while (t < n_t) {  // this lane's next row of the ring that its bound does not rule out
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
                    const int yy = q.cy + dy, zz = q.cz + dz;
                    if (yy < 0 || yy >= m.ny || zz < 0 || zz >= m.nz) continue;
                    const float g = __fadd_rd(gap_sq(ring_gap(ay, dy, m.hq)), gap_sq(ring_gap(az, dz, m.hq)));
                    if (bd < g) continue;
                    NN_STAT(2, 1);
                    row_range(m, q, g, bd, xa, xb);
                    ry = yy;
                    rz = zz;
                    break;
                }
            }
            W.emit_cells(xa, xb, ry, rz);
        }
        W.drain();
    }
}

}  // namespace ssf
