/*
 * ssf_synth.c -- deterministic synthetic street world: map clouds and LiDAR scans.
 *
 * This is the data source of bench.py and the tests.  It is
 * NOT part of the registration path and NOT part of the oracle: the reference ships
 * no sample data (its clouds are expected under $HOME/Desktop/map_data at run time,
 * reference localization/src/localization_node.cpp:7), so every workload here is
 * procedurally generated from a seed.
 *
 * World (all SI units, map frame):
 *   ground    z = 0.05 sin(x/7) cos(y/9), everywhere outside building footprints
 *   buildings one axis-aligned box per 52 m block cell (40 m block + 12 m street),
 *             footprint [52 bi + 6, 52 bi + 46] x [52 bj + 6, 52 bj + 46],
 *             height U(6, 30) m hashed from (seed, bi, bj)
 *   poles     12 vertical cylinders per block, 1 m outside each facade, every 15 m,
 *             radius U(0.15, 0.4) m, height U(3, 8) m
 *
 * Randomness is counter based (SplitMix64 of a structured key), so a point does not
 * depend on generation order and loops can be split across threads.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC ssf_synth.c -o libssf_synth.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PITCH 52.0
#define B_LO 6.0
#define B_HI 46.0
#define N_POLES 12

static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static inline uint64_t key4(uint64_t seed, uint64_t a, uint64_t b, uint64_t c, uint64_t d)
{
    uint64_t h = splitmix64(seed ^ 0x5353465f53594e54ull);
    h = splitmix64(h ^ a);
    h = splitmix64(h ^ b);
    h = splitmix64(h ^ c);
    h = splitmix64(h ^ d);
    return h;
}

/* uniform in [0,1) from a key and a draw number */
static inline double u01(uint64_t key, uint64_t draw)
{
    return (double)(splitmix64(key + 0x632BE59BD9B4E019ull * (draw + 1)) >> 11) * (1.0 / 9007199254740992.0);
}

static inline void gauss2(uint64_t key, uint64_t draw, double *g0, double *g1)
{
    double u = u01(key, draw), v = u01(key, draw + 1);
    if (u < 1e-300) u = 1e-300;
    double r = sqrt(-2.0 * log(u));
    *g0 = r * cos(6.283185307179586 * v);
    *g1 = r * sin(6.283185307179586 * v);
}

static inline double ground_z(double x, double y) { return 0.05 * sin(x / 7.0) * cos(y / 9.0); }

static inline double building_height(uint64_t seed, int64_t bi, int64_t bj)
{
    return 6.0 + 24.0 * u01(key4(seed, 1, (uint64_t)bi, (uint64_t)bj, 0), 0);
}

static inline int64_t block_of(double v) { return (int64_t)floor(v / PITCH); }

static inline int inside_footprint(double x, double y)
{
    double fx = x - PITCH * floor(x / PITCH), fy = y - PITCH * floor(y / PITCH);
    return fx > B_LO && fx < B_HI && fy > B_LO && fy < B_HI;
}

/* pole k (0..11) of block (bi,bj): centre, radius, height */
static inline void pole_of(uint64_t seed, int64_t bi, int64_t bj, int k, double *cx, double *cy, double *r, double *h)
{
    static const double along[3] = {5.0, 20.0, 35.0};
    double bx = PITCH * (double)bi, by = PITCH * (double)bj;
    int side = k / 3, j = k % 3;
    switch (side) {
    case 0: *cx = bx + B_LO + along[j]; *cy = by + B_LO - 1.0; break; /* south */
    case 1: *cx = bx + B_LO + along[j]; *cy = by + B_HI + 1.0; break; /* north */
    case 2: *cx = bx + B_LO - 1.0; *cy = by + B_LO + along[j]; break; /* west  */
    default: *cx = bx + B_HI + 1.0; *cy = by + B_LO + along[j]; break; /* east  */
    }
    uint64_t key = key4(seed, 2, (uint64_t)bi, (uint64_t)bj, (uint64_t)k);
    *r = 0.15 + 0.25 * u01(key, 0);
    *h = 3.0 + 5.0 * u01(key, 1);
}

/* ------------------------------------------------------------------------------------------
 * Map sampling.  Candidate points are enumerated in a fixed order (ground rows, then per
 * block: 4 facades, 12 poles); ssf_synth_map_count returns how many candidates the square
 * [-half, half]^2 holds, ssf_synth_map_fill keeps exactly m_keep of them by Bresenham
 * thinning over the candidate number, which keeps the spatial density uniform.
 * ---------------------------------------------------------------------------------------- */

typedef struct {
    uint64_t seed;
    double half, s, sigma;
    int64_t n_cand, m_keep; /* thinning; n_cand == 0 -> count only */
    float *xyz;             /* m_keep x 4 (w = 1) or NULL */
    float *nrm;             /* m_keep x 4 (w = 0) or NULL */
} mapgen_t;

static inline int64_t keep_slot(const mapgen_t *g, int64_t cand)
{
    /* candidate cand is kept iff floor((cand+1)*m/n) > floor(cand*m/n); slot = floor(cand*m/n) */
    __int128 a = (__int128)cand * g->m_keep / g->n_cand;
    __int128 b = (__int128)(cand + 1) * g->m_keep / g->n_cand;
    return b > a ? (int64_t)a : -1;
}

static inline void emit(const mapgen_t *g, int64_t cand, uint64_t key, double x, double y, double z, double nx,
                        double ny, double nz)
{
    int64_t slot = keep_slot(g, cand);
    if (slot < 0) return;
    double g0, g1, g2, g3;
    gauss2(key, 8, &g0, &g1);
    gauss2(key, 10, &g2, &g3);
    float *p = g->xyz + 4 * slot;
    p[0] = (float)(x + g->sigma * g0);
    p[1] = (float)(y + g->sigma * g1);
    p[2] = (float)(z + g->sigma * g2);
    p[3] = 1.0f;
    if (g->nrm) {
        float *n = g->nrm + 4 * slot;
        n[0] = (float)nx; n[1] = (float)ny; n[2] = (float)nz; n[3] = 0.0f;
    }
}

/* ground row j: returns number of candidates in the row; emits if filling */
static int64_t ground_row(const mapgen_t *g, int64_t j, int64_t n_side, int64_t cand0)
{
    int64_t c = 0;
    for (int64_t i = 0; i < n_side; ++i) {
        uint64_t key = key4(g->seed, 3, (uint64_t)i, (uint64_t)j, 0);
        double x = -g->half + ((double)i + u01(key, 0)) * g->s;
        double y = -g->half + ((double)j + u01(key, 1)) * g->s;
        if (inside_footprint(x, y)) continue;
        if (g->n_cand) {
            double dzdx = 0.05 / 7.0 * cos(x / 7.0) * cos(y / 9.0);
            double dzdy = -0.05 / 9.0 * sin(x / 7.0) * sin(y / 9.0);
            double inv = 1.0 / sqrt(dzdx * dzdx + dzdy * dzdy + 1.0);
            emit(g, cand0 + c, key, x, y, ground_z(x, y), -dzdx * inv, -dzdy * inv, inv);
        }
        ++c;
    }
    return c;
}

/* one block: facades + poles, clipped to the map square */
static int64_t block_surfaces(const mapgen_t *g, int64_t bi, int64_t bj, int64_t cand0)
{
    int64_t c = 0;
    const double bx = PITCH * (double)bi, by = PITCH * (double)bj;
    const double h = building_height(g->seed, bi, bj);
    const int64_t nu = (int64_t)ceil((B_HI - B_LO) / g->s), nv = (int64_t)ceil(h / g->s);
    for (int face = 0; face < 4; ++face) {
        for (int64_t v = 0; v < nv; ++v)
            for (int64_t u = 0; u < nu; ++u) {
                uint64_t key = key4(g->seed, 4 + (uint64_t)face, (uint64_t)bi, (uint64_t)bj, (uint64_t)(v * nu + u));
                double a = B_LO + ((double)u + u01(key, 0)) * g->s;
                double z = ((double)v + u01(key, 1)) * g->s;
                if (a > B_HI || z > h) continue;
                double x, y, nx = 0, ny = 0;
                switch (face) {
                case 0: x = bx + a; y = by + B_LO; ny = -1; break;
                case 1: x = bx + a; y = by + B_HI; ny = 1; break;
                case 2: x = bx + B_LO; y = by + a; nx = -1; break;
                default: x = bx + B_HI; y = by + a; nx = 1; break;
                }
                if (fabs(x) > g->half || fabs(y) > g->half) continue;
                if (g->n_cand) emit(g, cand0 + c, key, x, y, z, nx, ny, 0.0);
                ++c;
            }
    }
    for (int k = 0; k < N_POLES; ++k) {
        double cx, cy, r, ph;
        pole_of(g->seed, bi, bj, k, &cx, &cy, &r, &ph);
        if (fabs(cx) + r > g->half || fabs(cy) + r > g->half) continue;
        int64_t nt = (int64_t)ceil(6.283185307179586 * r / g->s), nz = (int64_t)ceil(ph / g->s);
        double gz = ground_z(cx, cy);
        for (int64_t v = 0; v < nz; ++v)
            for (int64_t u = 0; u < nt; ++u) {
                uint64_t key = key4(g->seed, 8 + (uint64_t)k, (uint64_t)bi, (uint64_t)bj, (uint64_t)(v * nt + u));
                double th = 6.283185307179586 * ((double)u + u01(key, 0)) / (double)nt;
                double z = ((double)v + u01(key, 1)) * g->s;
                if (z > ph) continue;
                if (g->n_cand) emit(g, cand0 + c, key, cx + r * cos(th), cy + r * sin(th), gz + z, cos(th), sin(th), 0.0);
                ++c;
            }
    }
    return c;
}

static int64_t mapgen_run(mapgen_t *g)
{
    const int64_t n_side = (int64_t)ceil(2.0 * g->half / g->s);
    const int64_t b0 = block_of(-g->half), b1 = block_of(g->half);
    const int64_t nb = b1 - b0 + 1;
    const int64_t n_units = n_side + nb * nb;
    int64_t *cnt = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_units + 1));
    if (!cnt) return -1;
    mapgen_t counter = *g;
    counter.n_cand = 0;
    /* pass 1: candidates per unit (ground row or block) */
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t u = 0; u < n_units; ++u) {
        if (u < n_side) cnt[u + 1] = ground_row(&counter, u, n_side, 0);
        else {
            int64_t b = u - n_side;
            cnt[u + 1] = block_surfaces(&counter, b0 + b % nb, b0 + b / nb, 0);
        }
    }
    cnt[0] = 0;
    for (int64_t u = 0; u < n_units; ++u) cnt[u + 1] += cnt[u];
    const int64_t total = cnt[n_units];
    if (g->n_cand) {
        if (g->n_cand != total) { free(cnt); return -2; }
#pragma omp parallel for schedule(dynamic, 8)
        for (int64_t u = 0; u < n_units; ++u) {
            if (u < n_side) ground_row(g, u, n_side, cnt[u]);
            else {
                int64_t b = u - n_side;
                block_surfaces(g, b0 + b % nb, b0 + b / nb, cnt[u]);
            }
        }
    }
    free(cnt);
    return total;
}

/* Number of candidate surface samples of the square [-half, half]^2 at grid spacing s. */
int64_t ssf_synth_map_count(uint64_t seed, double half, double spacing)
{
    mapgen_t g = {seed, half, spacing, 0.0, 0, 0, NULL, NULL};
    return mapgen_run(&g);
}

/* Fill exactly m_keep points (xyz: m_keep x 4 floats; nrm optional m_keep x 4).
 * n_cand must be the value ssf_synth_map_count returned for the same (seed, half, spacing).
 * Returns m_keep, or <0 on error. */
int64_t ssf_synth_map_fill(uint64_t seed, double half, double spacing, double sigma, int64_t n_cand, int64_t m_keep,
                           float *xyz, float *nrm)
{
    if (m_keep <= 0 || m_keep > n_cand || !xyz) return -3;
    mapgen_t g = {seed, half, spacing, sigma, n_cand, m_keep, xyz, nrm};
    int64_t t = mapgen_run(&g);
    return t < 0 ? t : m_keep;
}

/* ------------------------------------------------------------------------------------------
 * LiDAR scan: analytic ray cast from pose T (row-major 4x4 double, sensor -> map).
 * Output points are in the SENSOR frame (x forward), float4 with w = 1; rays with no hit
 * inside max_range are dropped.  Returns the number of points written (<= beams*azimuths).
 * ---------------------------------------------------------------------------------------- */
static double cast_ray(uint64_t seed, const double o[3], const double d[3], double max_range)
{
    double best = max_range;
    /* ground: fixed-point iteration on the height field, starting from the z = 0 plane */
    if (d[2] < -1e-6) {
        double t = -o[2] / d[2];
        for (int it = 0; it < 6; ++it) t = (ground_z(o[0] + t * d[0], o[1] + t * d[1]) - o[2]) / d[2];
        if (t > 0.0 && t < best && !inside_footprint(o[0] + t * d[0], o[1] + t * d[1])) best = t;
    }
    double x1 = o[0] + max_range * d[0], y1 = o[1] + max_range * d[1];
    int64_t bi0 = block_of(fmin(o[0], x1)), bi1 = block_of(fmax(o[0], x1));
    int64_t bj0 = block_of(fmin(o[1], y1)), bj1 = block_of(fmax(o[1], y1));
    for (int64_t bj = bj0; bj <= bj1; ++bj)
        for (int64_t bi = bi0; bi <= bi1; ++bi) {
            /* building box, slab test; only entry from outside counts */
            double lo[3] = {PITCH * (double)bi + B_LO, PITCH * (double)bj + B_LO, -1.0};
            double hi[3] = {PITCH * (double)bi + B_HI, PITCH * (double)bj + B_HI, building_height(seed, bi, bj)};
            double tn = 0.0, tf = best;
            int ok = 1;
            for (int k = 0; k < 3 && ok; ++k) {
                if (fabs(d[k]) < 1e-12) {
                    if (o[k] < lo[k] || o[k] > hi[k]) ok = 0;
                } else {
                    double ta = (lo[k] - o[k]) / d[k], tb = (hi[k] - o[k]) / d[k];
                    if (ta > tb) { double s = ta; ta = tb; tb = s; }
                    if (ta > tn) tn = ta;
                    if (tb < tf) tf = tb;
                    if (tn > tf) ok = 0;
                }
            }
            if (ok && tn > 0.0 && tn < best) best = tn;
            for (int k = 0; k < N_POLES; ++k) {
                double cx, cy, r, h;
                pole_of(seed, bi, bj, k, &cx, &cy, &r, &h);
                double ox = o[0] - cx, oy = o[1] - cy;
                double a = d[0] * d[0] + d[1] * d[1];
                if (a < 1e-12) continue;
                double b = ox * d[0] + oy * d[1], c = ox * ox + oy * oy - r * r;
                double disc = b * b - a * c;
                if (disc < 0.0) continue;
                double t = (-b - sqrt(disc)) / a;
                if (t <= 0.0 || t >= best) continue;
                double z = o[2] + t * d[2], gz = ground_z(cx, cy);
                if (z < gz - 0.5 || z > gz + h) continue;
                best = t;
            }
        }
    return best < max_range ? best : -1.0;
}

int64_t ssf_synth_scan(uint64_t world_seed, uint64_t scan_seed, const double *T, int beams, int azimuths,
                       double fov_lo_deg, double fov_hi_deg, double max_range, double range_sigma, float *xyz_out)
{
    const int64_t n_rays = (int64_t)beams * azimuths;
    double *rng = (double *)malloc(sizeof(double) * (size_t)n_rays);
    if (!rng) return -1;
    const double o[3] = {T[3], T[7], T[11]};
    const double deg = 3.14159265358979323846 / 180.0;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t ray = 0; ray < n_rays; ++ray) {
        int b = (int)(ray / azimuths), a = (int)(ray % azimuths);
        double el = deg * (beams > 1 ? fov_lo_deg + (fov_hi_deg - fov_lo_deg) * (double)b / (double)(beams - 1) : fov_lo_deg);
        double az = 6.283185307179586 * (double)a / (double)azimuths;
        double ds[3] = {cos(el) * cos(az), cos(el) * sin(az), sin(el)};
        double d[3];
        for (int k = 0; k < 3; ++k) d[k] = T[4 * k + 0] * ds[0] + T[4 * k + 1] * ds[1] + T[4 * k + 2] * ds[2];
        double t = cast_ray(world_seed, o, d, max_range);
        if (t > 0.0) {
            double g0, g1;
            gauss2(key4(scan_seed, 99, (uint64_t)ray, 0, 0), 0, &g0, &g1);
            t += range_sigma * g0;
            if (t < 0.05) t = 0.05;
        }
        rng[ray] = t;
    }
    int64_t n = 0;
    for (int64_t ray = 0; ray < n_rays; ++ray) {
        double t = rng[ray];
        if (t <= 0.0) continue;
        int b = (int)(ray / azimuths), a = (int)(ray % azimuths);
        double el = deg * (beams > 1 ? fov_lo_deg + (fov_hi_deg - fov_lo_deg) * (double)b / (double)(beams - 1) : fov_lo_deg);
        double az = 6.283185307179586 * (double)a / (double)azimuths;
        xyz_out[4 * n + 0] = (float)(t * cos(el) * cos(az));
        xyz_out[4 * n + 1] = (float)(t * cos(el) * sin(az));
        xyz_out[4 * n + 2] = (float)(t * sin(el));
        xyz_out[4 * n + 3] = 1.0f;
        ++n;
    }
    free(rng);
    return n;
}

/* Height of the ground under (x, y); exported so pose generators can place the sensor. */
double ssf_synth_ground(double x, double y) { return ground_z(x, y); }
