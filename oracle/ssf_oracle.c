/*
 * ssf_oracle.c -- CPU ORACLE for the scan-to-map registration hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under slam-sensor-fusion_b200/ (the product) may
 * include, link or call this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the checker or the
 * reported CPU baseline.
 *
 * PINNING.  The reference (viniciusvidal2/slam-sensor-fusion) ships no behavioural tests, golden
 * vectors or fixtures (only three ament lint stubs under localization_python/test/), and its
 * build needs PCL, FLANN, Eigen and ROS 2, none of which is installed here.  So:
 *   (1) oracle/_ref/libssf_ref.so is the reference's OWN icp_point_to_point.cpp,
 *       brute_force_alignment.cpp and point_cloud_processing.hpp compiled UNMODIFIED against
 *       stand-in Eigen/PCL headers (oracle/ref_stubs/, recipe in oracle/Makefile), and
 *       tests/test_ref_pin.py requires this file to agree with it BIT FOR BIT (pose, error,
 *       iterations, convergence flag, abort sentinel, debug text; fine / coarse / re-search-
 *       every-pass / abort cases, config 1 at full size; the pose-grid scorer; the three
 *       pre-processing functions).  Everything the reference spells out in its own text is
 *       therefore pinned to that text.
 *   (2) What stays restated from published algorithms, in the stand-ins as well as here  [ext]
 *       (versions unpinned in the reference's CMakeLists.txt:14-26):
 *   - pcl::KdTreeFLANN<PointXYZ>::nearestKSearch(k=1): exact 1-NN, squared L2 distance
 *     accumulated left to right in float without FMA (flann::L2_Simple; pinned bit-equal
 *     against cv2.flann, the same FLANN lineage).  FLANN's tie order is traversal order; the
 *     contract here is "lowest target index wins".
 *   - Eigen::JacobiSVD<Matrix3f>: two-sided Jacobi with the real 2x2 kernel; the summation
 *     order of Eigen's blocked GEMM for the 3x3 cross-covariance (host-cache dependent).
 *   - pcl::VoxelGrid<PointXYZ>::applyFilter: see ssf_oracle_voxel_grid.
 *       These are pinned by tests/test_oracle.py: O(N*M) brute force, cv2.flann KDTREE_SINGLE
 *       and scipy cKDTree for the NN; numpy.linalg.svd for the Kabsch step; a numpy group-by
 *       for the voxel grid; analytic ground-truth poses of the synthetic world for the loop.
 *   (3) The GN / Open3D-flow modes have no reference counterpart in C++ (Open3D is absent and
 *       unpinned): this file defines them; numpy/scipy restatements pin them (test_oracle.py).
 *
 * Build (oracle/Makefile): gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC
 * -ffp-contract=off mirrors the reference build, which passes no -march/-ffast-math flag
 * (localization/CMakeLists.txt:5-11), i.e. plain SSE2 float math with no FMA contraction.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ======================================================================================
 * 1. Exact nearest neighbour: KD-tree (leaf <= 15 like KDTreeSingleIndexParams(15)) and
 *    brute force.  Replaces kdtree_.setInputCloud (icp_point_to_point.cpp:54) and
 *    kdtree_.nearestKSearch(.., 1, ..) (icp_point_to_point.cpp:68).
 * ==================================================================================== */

typedef struct {
    int32_t left, right; /* children, -1 for a leaf */
    int32_t start, count; /* leaf: range in perm[] */
    int32_t dim;
    float divlow, divhigh; /* max of left subtree / min of right subtree along dim */
} kd_node_t;

typedef struct {
    int64_t n;
    float *pts;    /* n x 3, ORIGINAL order (kept for index -> point) */
    float *lpts;   /* n x 3, leaf order (cache friendly) */
    int32_t *perm; /* leaf order -> original index */
    kd_node_t *nodes;
    int32_t n_nodes, cap_nodes;
    float bb_lo[3], bb_hi[3];
} kdtree_t;

#define KD_LEAF 15

static inline float sqdist3(const float *a, const float *b)
{
    /* flann::L2_Simple: result += diff*diff, left to right, float, no FMA */
    float d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    float r = d0 * d0;
    r += d1 * d1;
    r += d2 * d2;
    return r;
}

static void kd_select(const float *pts, int32_t *idx, int64_t lo, int64_t hi, int64_t k, int dim)
{
    /* quickselect on idx[lo..hi) by pts[.][dim], ties by index so the tree is deterministic */
    while (hi - lo > 1) {
        int64_t mid = lo + (hi - lo) / 2;
        int32_t pi = idx[mid];
        float pv = pts[3 * (int64_t)pi + dim];
        int64_t i = lo, j = hi - 1;
        while (i <= j) {
            while (pts[3 * (int64_t)idx[i] + dim] < pv || (pts[3 * (int64_t)idx[i] + dim] == pv && idx[i] < pi)) ++i;
            while (pts[3 * (int64_t)idx[j] + dim] > pv || (pts[3 * (int64_t)idx[j] + dim] == pv && idx[j] > pi)) --j;
            if (i <= j) { int32_t t = idx[i]; idx[i] = idx[j]; idx[j] = t; ++i; --j; }
        }
        if (k <= j) hi = j + 1;
        else if (k >= i) lo = i;
        else return;
    }
}

static int32_t kd_new_node(kdtree_t *t)
{
    if (t->n_nodes == t->cap_nodes) {
        t->cap_nodes = t->cap_nodes ? 2 * t->cap_nodes : 1024;
        t->nodes = (kd_node_t *)realloc(t->nodes, sizeof(kd_node_t) * (size_t)t->cap_nodes);
    }
    return t->n_nodes++;
}

static int32_t kd_build_rec(kdtree_t *t, int64_t lo, int64_t hi)
{
    int32_t id = kd_new_node(t);
    if (hi - lo <= KD_LEAF) {
        kd_node_t nd = {-1, -1, (int32_t)lo, (int32_t)(hi - lo), 0, 0.f, 0.f};
        t->nodes[id] = nd;
        return id;
    }
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int64_t i = lo; i < hi; ++i) {
        const float *p = t->pts + 3 * (int64_t)t->perm[i];
        for (int k = 0; k < 3; ++k) { if (p[k] < mn[k]) mn[k] = p[k]; if (p[k] > mx[k]) mx[k] = p[k]; }
    }
    int dim = 0;
    if (mx[1] - mn[1] > mx[dim] - mn[dim]) dim = 1;
    if (mx[2] - mn[2] > mx[dim] - mn[dim]) dim = 2;
    int64_t mid = lo + (hi - lo) / 2;
    kd_select(t->pts, t->perm, lo, hi, mid, dim);
    float dl = -FLT_MAX, dh = FLT_MAX;
    for (int64_t i = lo; i < mid; ++i) { float v = t->pts[3 * (int64_t)t->perm[i] + dim]; if (v > dl) dl = v; }
    for (int64_t i = mid; i < hi; ++i) { float v = t->pts[3 * (int64_t)t->perm[i] + dim]; if (v < dh) dh = v; }
    int32_t l = kd_build_rec(t, lo, mid);
    int32_t r = kd_build_rec(t, mid, hi);
    kd_node_t nd = {l, r, 0, 0, dim, dl, dh};
    t->nodes[id] = nd;
    return id;
}

void *ssf_oracle_kdtree_build(const float *xyz, int64_t n, int stride_floats)
{
    kdtree_t *t = (kdtree_t *)calloc(1, sizeof(kdtree_t));
    t->n = n;
    t->pts = (float *)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
    t->lpts = (float *)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
    t->perm = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int k = 0; k < 3; ++k) { t->bb_lo[k] = FLT_MAX; t->bb_hi[k] = -FLT_MAX; }
    for (int64_t i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) {
            float v = xyz[i * stride_floats + k];
            t->pts[3 * i + k] = v;
            if (v < t->bb_lo[k]) t->bb_lo[k] = v;
            if (v > t->bb_hi[k]) t->bb_hi[k] = v;
        }
        t->perm[i] = (int32_t)i;
    }
    if (n > 0) kd_build_rec(t, 0, n);
    for (int64_t i = 0; i < n; ++i) memcpy(t->lpts + 3 * i, t->pts + 3 * (int64_t)t->perm[i], 3 * sizeof(float));
    return t;
}

void ssf_oracle_kdtree_free(void *tree)
{
    kdtree_t *t = (kdtree_t *)tree;
    if (!t) return;
    free(t->pts); free(t->lpts); free(t->perm); free(t->nodes); free(t);
}

typedef struct { const kdtree_t *t; const float *q; float best; int32_t best_idx; } kd_query_t;

static void kd_search_rec(kd_query_t *s, int32_t node, double mindist, double dists[3])
{
    const kd_node_t *nd = &s->t->nodes[node];
    if (nd->left < 0) {
        const float *lp = s->t->lpts + 3 * (int64_t)nd->start;
        const int32_t *pi = s->t->perm + nd->start;
        for (int32_t i = 0; i < nd->count; ++i) {
            float d = sqdist3(s->q, lp + 3 * i);
            if (d < s->best || (d == s->best && pi[i] < s->best_idx)) { s->best = d; s->best_idx = pi[i]; }
        }
        return;
    }
    double val = s->q[nd->dim];
    double diff1 = val - (double)nd->divlow, diff2 = val - (double)nd->divhigh;
    int32_t near_c, far_c;
    double cut;
    if (diff1 + diff2 < 0) { near_c = nd->left; far_c = nd->right; cut = diff2 * diff2; }
    else { near_c = nd->right; far_c = nd->left; cut = diff1 * diff1; }
    kd_search_rec(s, near_c, mindist, dists);
    double saved = dists[nd->dim];
    double far_min = mindist + cut - saved;
    /* far_min is the exact squared distance to the far child's box (in double).  The float
     * distance of any point inside is >= far_min*(1 - 2^-22); the factor keeps pruning
     * conservative so equal-distance candidates are still visited (tie -> lowest index). */
    if (far_min * (1.0 - 1e-6) <= (double)s->best) {
        dists[nd->dim] = cut;
        kd_search_rec(s, far_c, far_min, dists);
        dists[nd->dim] = saved;
    }
}

static void kd_nn_one(const kdtree_t *t, const float *q, int32_t *idx, float *d2)
{
    kd_query_t s = {t, q, FLT_MAX, INT32_MAX};
    if (t->n == 0) { *idx = -1; *d2 = FLT_MAX; return; }
    double dists[3], mind = 0.0;
    for (int k = 0; k < 3; ++k) {
        double d = 0.0;
        if (q[k] < t->bb_lo[k]) d = (double)t->bb_lo[k] - q[k];
        else if (q[k] > t->bb_hi[k]) d = (double)q[k] - t->bb_hi[k];
        dists[k] = d * d;
        mind += dists[k];
    }
    kd_search_rec(&s, 0, mind, dists);
    *idx = s.best_idx;
    *d2 = s.best;
}

/* k=1 search for nq queries; threads<=1 is the faithful single-threaded form. */
void ssf_oracle_kdtree_nn(const void *tree, const float *q, int64_t nq, int q_stride, int32_t *idx, float *d2,
                          int threads)
{
    const kdtree_t *t = (const kdtree_t *)tree;
#ifdef _OPENMP
    if (threads > 1) {
#pragma omp parallel for num_threads(threads) schedule(dynamic, 256)
        for (int64_t i = 0; i < nq; ++i) kd_nn_one(t, q + i * q_stride, idx + i, d2 + i);
        return;
    }
#endif
    (void)threads;
    for (int64_t i = 0; i < nq; ++i) kd_nn_one(t, q + i * q_stride, idx + i, d2 + i);
}

void ssf_oracle_nn_brute(const float *map, int64_t m, int m_stride, const float *q, int64_t nq, int q_stride,
                         int32_t *idx, float *d2)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < nq; ++i) {
        float best = FLT_MAX;
        int32_t bi = m > 0 ? INT32_MAX : -1;
        for (int64_t j = 0; j < m; ++j) {
            float d = sqdist3(q + i * q_stride, map + j * m_stride);
            if (d < best) { best = d; bi = (int32_t)j; } /* ascending j: first minimum = lowest index */
        }
        idx[i] = bi;
        d2[i] = best;
    }
}

int ssf_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ======================================================================================
 * 2. Small fixed-size linear algebra (Eigen is not installed).  Matrices are column-major
 *    like Eigen::Matrix4f / Matrix3f: M(r,c) = m[c*rows + r].
 * ==================================================================================== */
#define M3(m, r, c) (m)[(c)*3 + (r)]
#define M4(m, r, c) (m)[(c)*4 + (r)]

static void mat4_mul_f(const float *A, const float *B, float *C)
{
    float t[16];
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            float s = M4(A, r, 0) * M4(B, 0, c);
            s += M4(A, r, 1) * M4(B, 1, c);
            s += M4(A, r, 2) * M4(B, 2, c);
            s += M4(A, r, 3) * M4(B, 3, c);
            t[c * 4 + r] = s;
        }
    memcpy(C, t, sizeof(t));
}

/* rows p,q of W <- J * [row p; row q],  J = [[c, s], [-s, c]]   (Eigen applyOnTheLeft) */
static void rot_left3(float *W, int p, int q, float c, float s)
{
    for (int k = 0; k < 3; ++k) {
        float x = M3(W, p, k), y = M3(W, q, k);
        M3(W, p, k) = c * x + s * y;
        M3(W, q, k) = -s * x + c * y;
    }
}
/* cols p,q of W <- [col p, col q] * J   (Eigen applyOnTheRight) */
static void rot_right3(float *W, int p, int q, float c, float s)
{
    for (int k = 0; k < 3; ++k) {
        float x = M3(W, k, p), y = M3(W, k, q);
        M3(W, k, p) = c * x - s * y;
        M3(W, k, q) = s * x + c * y;
    }
}

/* Two-sided Jacobi SVD of a 3x3 float matrix, H = U diag(S) V^T, singular values sorted
 * descending -- the published algorithm of Eigen::JacobiSVD for a square real matrix
 * (used at icp_point_to_point.cpp:137 with ComputeFullU | ComputeFullV). */
static void jacobi_svd3_f(const float *H, float *U, float *S, float *V)
{
    const float precision = 2.0f * FLT_EPSILON, tiny = FLT_MIN;
    float W[9];
    float scale = 0.f;
    for (int i = 0; i < 9; ++i) { float a = fabsf(H[i]); if (a > scale) scale = a; }
    if (!(scale > 0.f) || !isfinite(scale)) scale = 1.f;
    for (int i = 0; i < 9; ++i) { W[i] = H[i] / scale; U[i] = V[i] = (i % 4 == 0) ? 1.f : 0.f; }
    float maxdiag = fmaxf(fabsf(M3(W, 0, 0)), fmaxf(fabsf(M3(W, 1, 1)), fabsf(M3(W, 2, 2))));
    int finished = 0, sweeps = 0;
    while (!finished && sweeps++ < 64) {
        finished = 1;
        for (int p = 1; p < 3; ++p)
            for (int q = 0; q < p; ++q) {
                float thr = fmaxf(tiny, precision * maxdiag);
                if (!(fabsf(M3(W, p, q)) > thr || fabsf(M3(W, q, p)) > thr)) continue;
                finished = 0;
                /* real_2x2_jacobi_svd on [[W(p,p), W(p,q)], [W(q,p), W(q,q)]] */
                float m00 = M3(W, p, p), m01 = M3(W, p, q), m10 = M3(W, q, p), m11 = M3(W, q, q);
                float t = m00 + m11, d = m10 - m01, c1, s1;
                if (fabsf(d) < tiny) { s1 = 0.f; c1 = 1.f; }
                else { float u = t / d, tmp = sqrtf(1.f + u * u); s1 = 1.f / tmp; c1 = u / tmp; }
                /* m <- rot1 applied on the left */
                float a00 = c1 * m00 + s1 * m10, a01 = c1 * m01 + s1 * m11, a11 = -s1 * m01 + c1 * m11;
                /* makeJacobi on the now symmetric 2x2 (x = a00, y = a01, z = a11) */
                float cr, sr, deno = 2.f * fabsf(a01);
                if (deno < tiny) { cr = 1.f; sr = 0.f; }
                else {
                    float tau = (a00 - a11) / deno, w = sqrtf(tau * tau + 1.f);
                    float tt = tau > 0.f ? 1.f / (tau + w) : 1.f / (tau - w);
                    float sign_t = tt > 0.f ? 1.f : -1.f, nn = 1.f / sqrtf(tt * tt + 1.f);
                    sr = -sign_t * (a01 / fabsf(a01)) * fabsf(tt) * nn;
                    cr = nn;
                }
                /* j_left = rot1 * j_right^T */
                float cl = c1 * cr - s1 * (-sr), sl = c1 * (-sr) + s1 * cr;
                rot_left3(W, p, q, cl, sl);
                rot_right3(U, p, q, cl, -sl); /* U.applyOnTheRight(p, q, j_left.transpose()) */
                rot_right3(W, p, q, cr, sr);
                rot_right3(V, p, q, cr, sr);
                maxdiag = fmaxf(maxdiag, fmaxf(fabsf(M3(W, p, p)), fabsf(M3(W, q, q))));
            }
    }
    for (int i = 0; i < 3; ++i) {
        float a = fabsf(M3(W, i, i));
        S[i] = a;
        if (a != 0.f) { float f = M3(W, i, i) / a; for (int k = 0; k < 3; ++k) M3(U, k, i) *= f; }
    }
    for (int i = 0; i < 3; ++i) S[i] *= scale;
    for (int i = 0; i < 3; ++i) {
        int pos = i;
        for (int k = i + 1; k < 3; ++k) if (S[k] > S[pos]) pos = k;
        if (S[pos] == 0.f) break;
        if (pos != i) {
            float ts = S[i]; S[i] = S[pos]; S[pos] = ts;
            for (int k = 0; k < 3; ++k) {
                float tu = M3(U, k, i); M3(U, k, i) = M3(U, k, pos); M3(U, k, pos) = tu;
                float tv = M3(V, k, i); M3(V, k, i) = M3(V, k, pos); M3(V, k, pos) = tv;
            }
        }
    }
}

void ssf_oracle_svd3(const float *H, float *U, float *S, float *V) { jacobi_svd3_f(H, U, S, V); }

static void mat3_mul_abt_f(const float *A, const float *B, float *C) /* C = A * B^T */
{
    /* fixed 3x3 lazy product: each coefficient is a size-3 reduction, a0 + (a1 + a2)  [ext] */
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            float s12 = M3(A, r, 1) * M3(B, c, 1) + M3(A, r, 2) * M3(B, c, 2);
            M3(C, r, c) = M3(A, r, 0) * M3(B, c, 0) + s12;
        }
}

static float det3_f(const float *R)
{
    return M3(R, 0, 0) * (M3(R, 1, 1) * M3(R, 2, 2) - M3(R, 1, 2) * M3(R, 2, 1)) -
           M3(R, 0, 1) * (M3(R, 1, 0) * M3(R, 2, 2) - M3(R, 1, 2) * M3(R, 2, 0)) +
           M3(R, 0, 2) * (M3(R, 1, 0) * M3(R, 2, 1) - M3(R, 1, 1) * M3(R, 2, 0));
}

/* ======================================================================================
 * 3. REFERENCE mode: restatement of ICPPointToPoint (icp_point_to_point.cpp).
 *    Clouds are kept as three float planes like Eigen::MatrixX3f (column-major).
 * ==================================================================================== */

typedef struct {
    float max_correspondence_dist; /* compared against the SQUARED distance (cpp:70) */
    int32_t num_iterations;
    float acceptable_mean_error;
    float transformation_epsilon;
} ssf_oracle_params;

typedef struct {
    float transformation[16]; /* column-major 4x4 */
    float error;
    int32_t iterations;
    int32_t has_converged;
    int32_t n_searches;
    int32_t k_final;
    int32_t aborted; /* 1: fewer than 10 correspondences on the first search (cpp:196-200) */
} ssf_oracle_result;

/* Optional trace of every correspondence search so a GPU kernel can be checked on the
 * oracle's own query arrays.  All arrays are caller allocated; cap_searches searches of up
 * to n_source queries each are recorded. */
typedef struct {
    int32_t cap_searches;
    int32_t n_source;
    int32_t *count;   /* [cap]            queries in search s */
    float *queries;   /* [cap][n][3]      query coordinates (x,y,z interleaved) */
    int32_t *rows;    /* [cap][n]         original source row of each query */
    int32_t *idx;     /* [cap][n]         NN index found */
    float *d2;        /* [cap][n]         squared distance */
    float *iter_err;  /* [num_iterations] error measured at the top of pass i (NaN if not run) */
    int32_t *iter_searched; /* [num_iterations] 1 if pass i re-searched */
} ssf_oracle_trace;

typedef struct { int64_t n; float *x, *y, *z; int32_t *row; } cloud_t;

static void cloud_alloc(cloud_t *c, int64_t n)
{
    c->n = n;
    size_t m = (size_t)(n > 0 ? n : 1);
    c->x = (float *)malloc(sizeof(float) * m);
    c->y = (float *)malloc(sizeof(float) * m);
    c->z = (float *)malloc(sizeof(float) * m);
    c->row = (int32_t *)malloc(sizeof(int32_t) * m);
}
static void cloud_free(cloud_t *c) { free(c->x); free(c->y); free(c->z); free(c->row); }

/* applyTransformation, icp_point_to_point.cpp:99-110: expression order kept */
static void apply_transformation(const float *T, cloud_t *c)
{
    for (int64_t i = 0; i < c->n; ++i) {
        const float x = c->x[i], y = c->y[i], z = c->z[i];
        const float tx = M4(T, 0, 0) * x + M4(T, 0, 1) * y + M4(T, 0, 2) * z + M4(T, 0, 3);
        const float ty = M4(T, 1, 0) * x + M4(T, 1, 1) * y + M4(T, 1, 2) * z + M4(T, 1, 3);
        const float tz = M4(T, 2, 0) * x + M4(T, 2, 1) * y + M4(T, 2, 2) * z + M4(T, 2, 3);
        c->x[i] = tx; c->y[i] = ty; c->z[i] = tz;
    }
}

/* sourceTargetCorrespondences, icp_point_to_point.cpp:57-84: k=1 search per row, keep the
 * row iff d2 < max_correspondence_dist_ (squared vs unsquared, cpp:70), then shrink both
 * clouds to the kept rows in original order (cpp:77-83). */
static void source_target_correspondences(const kdtree_t *tree, float thr, cloud_t *src, cloud_t *tgt,
                                          ssf_oracle_trace *tr, int search_no, int threads, int32_t *corr_out)
{
    const int64_t n = src->n;
    float *q = (float *)calloc(3 * (size_t)(n > 0 ? n : 1), sizeof(float));
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    float *d2 = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) { q[3 * i] = src->x[i]; q[3 * i + 1] = src->y[i]; q[3 * i + 2] = src->z[i]; }
    ssf_oracle_kdtree_nn(tree, q, n, 3, idx, d2, threads);
    if (tr && search_no < tr->cap_searches) {
        size_t off = (size_t)search_no * (size_t)tr->n_source;
        tr->count[search_no] = (int32_t)n;
        memcpy(tr->queries + 3 * off, q, sizeof(float) * 3 * (size_t)n);
        memcpy(tr->rows + off, src->row, sizeof(int32_t) * (size_t)n);
        memcpy(tr->idx + off, idx, sizeof(int32_t) * (size_t)n);
        memcpy(tr->d2 + off, d2, sizeof(float) * (size_t)n);
    }
    int64_t k = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (idx[i] >= 0 && d2[i] < thr) {
            const float *p = tree->pts + 3 * (int64_t)idx[i];
            src->x[k] = src->x[i]; src->y[k] = src->y[i]; src->z[k] = src->z[i]; src->row[k] = src->row[i];
            tgt->x[k] = p[0]; tgt->y[k] = p[1]; tgt->z[k] = p[2];
            if (corr_out) corr_out[src->row[i]] = idx[i];
            ++k;
        } else if (corr_out) corr_out[src->row[i]] = -1;
    }
    src->n = tgt->n = k;
    free(q); free(idx); free(d2);
}

/* calculateErrorMetric, icp_point_to_point.cpp:161-170: mean UNSQUARED distance,
 * sequential float accumulation */
static float calculate_error_metric(const cloud_t *s, const cloud_t *t)
{
    float error = 0.0f;
    for (int64_t i = 0; i < s->n; ++i) {
        float dx = s->x[i] - t->x[i], dy = s->y[i] - t->y[i], dz = s->z[i] - t->z[i];
        /* Eigen's fixed-size-3 reduction tree (redux_novec_unroller): a0 + (a1 + a2)  [ext] */
        float yz = dy * dy + dz * dz;
        float sq = dx * dx + yz;
        error += sqrtf(sq);
    }
    return error / (float)s->n;
}

/* calculateStepBestTransformation, icp_point_to_point.cpp:112-159 (Kabsch) */
static void calculate_step_best_transformation(const cloud_t *s, const cloud_t *t, float *T_step)
{
    const int64_t n = s->n;
    float cs[3] = {0, 0, 0}, ct[3] = {0, 0, 0};
    for (int64_t i = 0; i < n; ++i) { /* cpp:117-121 */
        cs[0] += s->x[i]; cs[1] += s->y[i]; cs[2] += s->z[i];
        ct[0] += t->x[i]; ct[1] += t->y[i]; ct[2] += t->z[i];
    }
    for (int k = 0; k < 3; ++k) { cs[k] /= (float)n; ct[k] /= (float)n; } /* cpp:122-123 */
    float H[9] = {0};
    for (int64_t i = 0; i < n; ++i) { /* cpp:126-134: zero-mean copies, H = Ps^T * Pt */
        float a[3] = {s->x[i] - cs[0], s->y[i] - cs[1], s->z[i] - cs[2]};
        float b[3] = {t->x[i] - ct[0], t->y[i] - ct[1], t->z[i] - ct[2]};
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) M3(H, r, c) += a[r] * b[c];
    }
    float U[9], S[3], V[9], R[9];
    jacobi_svd3_f(H, U, S, V);          /* cpp:137-139 */
    mat3_mul_abt_f(V, U, R);            /* cpp:142 R = V U^T */
    if (det3_f(R) < 0) {                /* cpp:145-149 */
        for (int k = 0; k < 3; ++k) M3(V, k, 2) *= -1.f;
        mat3_mul_abt_f(V, U, R);
    }
    float tr[3]; /* cpp:152 t = ct - R cs */
    for (int r = 0; r < 3; ++r) {
        float a12 = M3(R, r, 1) * cs[1] + M3(R, r, 2) * cs[2];
        float acc = M3(R, r, 0) * cs[0] + a12; /* size-3 reduction a0 + (a1 + a2)  [ext] */
        tr[r] = ct[r] - acc;
    }
    for (int i = 0; i < 16; ++i) T_step[i] = (i % 5 == 0) ? 1.f : 0.f;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) M4(T_step, r, c) = M3(R, r, c);
    for (int r = 0; r < 3; ++r) M4(T_step, r, 3) = tr[r];
}

void ssf_oracle_kabsch(const float *src_xyz, const float *tgt_xyz, int64_t n, float *T_step)
{
    cloud_t s, t;
    cloud_alloc(&s, n); cloud_alloc(&t, n);
    for (int64_t i = 0; i < n; ++i) {
        s.x[i] = src_xyz[3 * i]; s.y[i] = src_xyz[3 * i + 1]; s.z[i] = src_xyz[3 * i + 2];
        t.x[i] = tgt_xyz[3 * i]; t.y[i] = tgt_xyz[3 * i + 1]; t.z[i] = tgt_xyz[3 * i + 2];
    }
    calculate_step_best_transformation(&s, &t, T_step);
    cloud_free(&s); cloud_free(&t);
}

/* calculateAlignment, icp_point_to_point.cpp:185-254.
 * corr_out (optional, n_source ints): final correspondence per ORIGINAL source row, -1 if
 * the row was dropped.  Divergence from the reference, by contract: a re-search that leaves
 * zero correspondences (reference: 0/0 -> NaN pose, cpp:122-123,169) stops the loop here. */
int ssf_oracle_icp_reference(const void *tree_v, const float *src_xyz, int64_t n_source, int src_stride,
                             const float *T_init, const ssf_oracle_params *prm, ssf_oracle_result *res,
                             ssf_oracle_trace *tr, int32_t *corr_out, int threads)
{
    const kdtree_t *tree = (const kdtree_t *)tree_v;
    memcpy(res->transformation, T_init, sizeof(float) * 16); /* ICPResult(initial_transform_), cpp:188 */
    res->error = 1e6f; res->iterations = 0; res->has_converged = 0;
    res->n_searches = 0; res->k_final = 0; res->aborted = 0;
    if (tr) for (int i = 0; i < prm->num_iterations; ++i) { tr->iter_err[i] = NAN; tr->iter_searched[i] = 0; }

    cloud_t P, Q;
    cloud_alloc(&P, n_source); cloud_alloc(&Q, n_source);
    for (int64_t i = 0; i < n_source; ++i) { /* convertPclToEigen cpp:86-97 + copy cpp:191 */
        P.x[i] = src_xyz[i * src_stride]; P.y[i] = src_xyz[i * src_stride + 1]; P.z[i] = src_xyz[i * src_stride + 2];
        P.row[i] = (int32_t)i;
    }
    if (corr_out) for (int64_t i = 0; i < n_source; ++i) corr_out[i] = -1;
    apply_transformation(T_init, &P);                                                     /* cpp:192 */
    source_target_correspondences(tree, prm->max_correspondence_dist, &P, &Q, tr, res->n_searches++, threads, corr_out); /* cpp:195 */
    res->k_final = (int32_t)P.n;
    if (P.n < 10) { /* cpp:196-200 */
        res->aborted = 1;
        cloud_free(&P); cloud_free(&Q);
        return 0;
    }
    float T[16];
    memcpy(T, T_init, sizeof(T));                                                         /* cpp:203 */
    int iterations_taken = 0;
    float last_error = FLT_MAX;                                                           /* cpp:205 */
    for (int i = 0; i < prm->num_iterations; ++i) {                                       /* cpp:206 */
        const float error = calculate_error_metric(&P, &Q);                               /* cpp:209 */
        if (tr) tr->iter_err[i] = error;
        if (error < prm->acceptable_mean_error) { last_error = error; break; }            /* cpp:215-219 */
        if (fabsf(last_error - error) < prm->transformation_epsilon) {                    /* cpp:221-224 */
            source_target_correspondences(tree, prm->max_correspondence_dist, &P, &Q, tr, res->n_searches++, threads, corr_out);
            if (tr) tr->iter_searched[i] = 1;
            if (P.n == 0) { last_error = error; break; } /* contract: see header comment */
        }
        float T_step[16];
        calculate_step_best_transformation(&P, &Q, T_step);                               /* cpp:226 */
        mat4_mul_f(T_step, T, T);                                                         /* cpp:228 */
        apply_transformation(T_step, &P);                                                 /* cpp:230 */
        last_error = error;                                                               /* cpp:232 */
        ++iterations_taken;                                                               /* cpp:234 */
    }
    memcpy(res->transformation, T, sizeof(T));                                            /* cpp:249 */
    res->error = last_error;
    res->iterations = iterations_taken;
    res->has_converged = last_error < prm->acceptable_mean_error;
    res->k_final = (int32_t)P.n;
    cloud_free(&P); cloud_free(&Q);
    return 0;
}

/* ======================================================================================
 * 4. GN modes (north-star solver; no reference counterpart, so this restatement IS the
 *    contract -- SURVEY.md Appendix B.5) and the Open3D control flow of the Python twin
 *    (localization_python/localization_python/localization_node.py:233-237).
 *
 *    Every iteration: P = fl(T)*src with the reference's float expression
 *    (icp_point_to_point.cpp:103-105), k=1 search for EVERY source point, accept iff
 *    d2 < max_correspondence_dist, accumulate in double in source order.
 * ==================================================================================== */

static int cholesky_solve6(const double *A_in, const double *b, double *x)
{
    double L[36];
    memcpy(L, A_in, sizeof(L));
    for (int j = 0; j < 6; ++j) {
        double d = L[j * 6 + j];
        for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k];
        if (!(d > 0.0)) return -1;
        d = sqrt(d);
        L[j * 6 + j] = d;
        for (int i = j + 1; i < 6; ++i) {
            double s = L[i * 6 + j];
            for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
            L[i * 6 + j] = s / d;
        }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * y[k];
        y[i] = s / L[i * 6 + i];
    }
    for (int i = 5; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * x[k];
        x[i] = s / L[i * 6 + i];
    }
    return 0;
}

/* T_step (double, column-major) = [exp([w]x) | t] */
static void se3_from_twist(const double *x, double *Ts)
{
    double wx = x[0], wy = x[1], wz = x[2];
    double th2 = wx * wx + wy * wy + wz * wz, th = sqrt(th2);
    double a, b; /* R = I + a [w]x + b [w]x^2 */
    if (th < 1e-8) { a = 1.0 - th2 / 6.0; b = 0.5 - th2 / 24.0; }
    else { a = sin(th) / th; b = (1.0 - cos(th)) / th2; }
    double K[9] = {0, wz, -wy, -wz, 0, wx, wy, -wx, 0}; /* column-major [w]x */
    double K2[9];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += K[k * 3 + r] * K[c * 3 + k];
            K2[c * 3 + r] = s;
        }
    for (int i = 0; i < 16; ++i) Ts[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) Ts[c * 4 + r] = (r == c ? 1.0 : 0.0) + a * K[c * 3 + r] + b * K2[c * 3 + r];
    Ts[12] = x[3]; Ts[13] = x[4]; Ts[14] = x[5];
}

static void compose_round(const double *Ts, float *T) /* T <- fl(Ts * T) */
{
    double r[16];
    for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 4; ++rr) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += Ts[k * 4 + rr] * (double)T[c * 4 + k];
            r[c * 4 + rr] = s;
        }
    for (int i = 0; i < 16; ++i) T[i] = (float)r[i];
}

static void transform_all(const float *T, const float *src, int64_t n, int stride, float *P)
{
    for (int64_t i = 0; i < n; ++i) {
        const float x = src[i * stride], y = src[i * stride + 1], z = src[i * stride + 2];
        P[3 * i + 0] = M4(T, 0, 0) * x + M4(T, 0, 1) * y + M4(T, 0, 2) * z + M4(T, 0, 3);
        P[3 * i + 1] = M4(T, 1, 0) * x + M4(T, 1, 1) * y + M4(T, 1, 2) * z + M4(T, 1, 3);
        P[3 * i + 2] = M4(T, 2, 0) * x + M4(T, 2, 1) * y + M4(T, 2, 2) * z + M4(T, 2, 3);
    }
}

/* mode 0: point-to-point  r = p - q (3 rows),  J = [-[p]x | I]
 * mode 1: point-to-plane  r = n.(p - q),       J = [p x n ; n]^T   (needs normals, stride 4)
 * Loop (the contract):
 *   for it < num_iterations: search; K < 10 on the first pass -> abort like cpp:196-200,
 *   K < 6 later -> stop;  err = sqrt(sum r^2 / K);  err < acceptable -> converged, stop;
 *   solve (J^T J) x = -J^T r by Cholesky in double (failure -> stop);
 *   T <- fl(exp(x) * T); ++iterations;  max|x_i| < transformation_epsilon -> converged, stop.
 */
int ssf_oracle_icp_gn(const void *tree_v, const float *normals, const float *src_xyz, int64_t n_source,
                      int src_stride, const float *T_init, const ssf_oracle_params *prm, int mode,
                      ssf_oracle_result *res, int32_t *corr_out, int threads)
{
    const kdtree_t *tree = (const kdtree_t *)tree_v;
    memcpy(res->transformation, T_init, sizeof(float) * 16);
    res->error = 1e6f; res->iterations = 0; res->has_converged = 0;
    res->n_searches = 0; res->k_final = 0; res->aborted = 0;
    if (mode == 1 && !normals) return -1;
    float T[16];
    memcpy(T, T_init, sizeof(T));
    size_t m = (size_t)(n_source > 0 ? n_source : 1);
    float *P = (float *)malloc(sizeof(float) * 3 * m), *d2 = (float *)malloc(sizeof(float) * m);
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * m);
    float err = 1e6f;
    int converged = 0, it = 0;
    for (int i = 0; i < prm->num_iterations; ++i) {
        transform_all(T, src_xyz, n_source, src_stride, P);
        ssf_oracle_kdtree_nn(tree, P, n_source, 3, idx, d2, threads);
        res->n_searches++;
        double A[36] = {0}, b[6] = {0}, sr2 = 0.0;
        int64_t K = 0;
        for (int64_t j = 0; j < n_source; ++j) {
            int ok = idx[j] >= 0 && d2[j] < prm->max_correspondence_dist;
            if (corr_out) corr_out[j] = ok ? idx[j] : -1;
            if (!ok) continue;
            ++K;
            const double p[3] = {P[3 * j], P[3 * j + 1], P[3 * j + 2]};
            const float *qf = tree->pts + 3 * (int64_t)idx[j];
            const double e[3] = {p[0] - qf[0], p[1] - qf[1], p[2] - qf[2]};
            if (mode == 1) {
                const float *nf = normals + 4 * (int64_t)idx[j];
                const double n[3] = {nf[0], nf[1], nf[2]};
                const double a[6] = {p[1] * n[2] - p[2] * n[1], p[2] * n[0] - p[0] * n[2], p[0] * n[1] - p[1] * n[0],
                                     n[0], n[1], n[2]};
                const double r = n[0] * e[0] + n[1] * e[1] + n[2] * e[2];
                for (int u = 0; u < 6; ++u) { b[u] += a[u] * r; for (int v = 0; v < 6; ++v) A[u * 6 + v] += a[u] * a[v]; }
                sr2 += r * r;
            } else {
                /* rows of J: [ -[p]x | I ];  -[p]x = [[0, pz, -py], [-pz, 0, px], [py, -px, 0]] */
                const double J[3][6] = {{0, p[2], -p[1], 1, 0, 0}, {-p[2], 0, p[0], 0, 1, 0}, {p[1], -p[0], 0, 0, 0, 1}};
                for (int rr = 0; rr < 3; ++rr)
                    for (int u = 0; u < 6; ++u) {
                        b[u] += J[rr][u] * e[rr];
                        for (int v = 0; v < 6; ++v) A[u * 6 + v] += J[rr][u] * J[rr][v];
                    }
                sr2 += e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
            }
        }
        res->k_final = (int32_t)K;
        if (i == 0 && K < 10) { res->aborted = 1; free(P); free(d2); free(idx); return 0; }
        if (K < 6) break;
        err = (float)sqrt(sr2 / (double)K);
        if (err < prm->acceptable_mean_error) { converged = 1; break; }
        double x[6], nb[6];
        for (int u = 0; u < 6; ++u) nb[u] = -b[u];
        if (cholesky_solve6(A, nb, x) != 0) break;
        double Ts[16];
        se3_from_twist(x, Ts);
        compose_round(Ts, T);
        ++it;
        double mx = 0;
        for (int u = 0; u < 6; ++u) if (fabs(x[u]) > mx) mx = fabs(x[u]);
        if (mx < (double)prm->transformation_epsilon) { converged = 1; break; }
    }
    memcpy(res->transformation, T, sizeof(T));
    res->error = err; res->iterations = it; res->has_converged = converged;
    free(P); free(d2); free(idx);
    return 0;
}

/* Kabsch / Umeyama (no scaling) from double moments: R = V diag(1,1,det) U^T of
 * H = sum (p - pbar)(q - qbar)^T, via the symmetric eigen-decomposition-free route:
 * one-sided Jacobi in double. */
static void jacobi_svd3_d(const double *H, double *U, double *S, double *V)
{
    /* one-sided Jacobi (Hestenes) on columns of A = H: A V = U S */
    double A[9];
    memcpy(A, H, sizeof(A));
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int k = 0; k < 3; ++k) {
                    alpha += A[p * 3 + k] * A[p * 3 + k];
                    beta += A[q * 3 + k] * A[q * 3 + k];
                    gamma += A[p * 3 + k] * A[q * 3 + k];
                }
                if (gamma == 0.0) continue;
                double lim = fabs(gamma) / sqrt(alpha * beta + 1e-300);
                if (lim > off) off = lim;
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < 3; ++k) {
                    double ap = A[p * 3 + k], aq = A[q * 3 + k];
                    A[p * 3 + k] = c * ap - s * aq; A[q * 3 + k] = s * ap + c * aq;
                    double vp = V[p * 3 + k], vq = V[q * 3 + k];
                    V[p * 3 + k] = c * vp - s * vq; V[q * 3 + k] = s * vp + c * vq;
                }
            }
        if (off < 1e-15) break;
    }
    for (int j = 0; j < 3; ++j) {
        double nrm = sqrt(A[j * 3] * A[j * 3] + A[j * 3 + 1] * A[j * 3 + 1] + A[j * 3 + 2] * A[j * 3 + 2]);
        S[j] = nrm;
    }
    /* sort descending */
    for (int i = 0; i < 3; ++i) {
        int pos = i;
        for (int k = i + 1; k < 3; ++k) if (S[k] > S[pos]) pos = k;
        if (pos != i) {
            double ts = S[i]; S[i] = S[pos]; S[pos] = ts;
            for (int k = 0; k < 3; ++k) {
                double ta = A[i * 3 + k]; A[i * 3 + k] = A[pos * 3 + k]; A[pos * 3 + k] = ta;
                double tv = V[i * 3 + k]; V[i * 3 + k] = V[pos * 3 + k]; V[pos * 3 + k] = tv;
            }
        }
    }
    /* U columns; complete a rank-deficient basis by cross products */
    for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) U[j * 3 + k] = S[j] > 1e-300 ? A[j * 3 + k] / S[j] : 0.0;
    if (!(S[1] > S[0] * 1e-14)) { /* rank <= 1: pick any unit vector orthogonal to u0 */
        double *u0 = U, *u1 = U + 3;
        int k = fabs(u0[0]) < fabs(u0[1]) ? (fabs(u0[0]) < fabs(u0[2]) ? 0 : 2) : (fabs(u0[1]) < fabs(u0[2]) ? 1 : 2);
        double e[3] = {0, 0, 0}; e[k] = 1.0;
        double d = e[0] * u0[0] + e[1] * u0[1] + e[2] * u0[2];
        for (int i = 0; i < 3; ++i) u1[i] = e[i] - d * u0[i];
        double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
        for (int i = 0; i < 3; ++i) u1[i] /= n1;
    }
    if (!(S[2] > S[0] * 1e-14)) { /* rank <= 2: u2 = u0 x u1 (sign fixed later by det) */
        double *u0 = U, *u1 = U + 3, *u2 = U + 6;
        u2[0] = u0[1] * u1[2] - u0[2] * u1[1];
        u2[1] = u0[2] * u1[0] - u0[0] * u1[2];
        u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
    }
}

static double det3_d(const double *R)
{
    return R[0] * (R[4] * R[8] - R[7] * R[5]) - R[3] * (R[1] * R[8] - R[7] * R[2]) + R[6] * (R[1] * R[5] - R[4] * R[2]);
}

/* Open3D control flow (Appendix B.4) on float32 geometry:
 *   search(T0); loop i < max_iteration: T <- fl(Kabsch(P,Q) * T); search(T);
 *   stop when |d fitness| < 1e-6 and |d rmse| < 1e-6.
 * max_correspondence_dist is compared with d2 (the caller squares a metric radius).
 * result.error = inlier_rmse, fitness returned separately. */
int ssf_oracle_icp_o3d(const void *tree_v, const float *src_xyz, int64_t n_source, int src_stride,
                       const float *T_init, const ssf_oracle_params *prm, ssf_oracle_result *res, float *fitness_out,
                       int32_t *corr_out, int threads)
{
    const kdtree_t *tree = (const kdtree_t *)tree_v;
    float T[16];
    memcpy(T, T_init, sizeof(T));
    memcpy(res->transformation, T_init, sizeof(T));
    res->error = 0.f; res->iterations = 0; res->has_converged = 0; res->n_searches = 0; res->k_final = 0; res->aborted = 0;
    size_t m = (size_t)(n_source > 0 ? n_source : 1);
    float *P = (float *)malloc(sizeof(float) * 3 * m), *d2 = (float *)malloc(sizeof(float) * m);
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * m);
    double fitness = 0, rmse = 0, prev_fit = 0, prev_rmse = 0;
    int it = 0, converged = 0;
    for (int pass = 0; pass <= prm->num_iterations; ++pass) {
        transform_all(T, src_xyz, n_source, src_stride, P);
        ssf_oracle_kdtree_nn(tree, P, n_source, 3, idx, d2, threads);
        res->n_searches++;
        int64_t K = 0;
        double sd2 = 0, sp[3] = {0}, sq[3] = {0};
        for (int64_t j = 0; j < n_source; ++j) {
            int ok = idx[j] >= 0 && d2[j] < prm->max_correspondence_dist;
            if (corr_out) corr_out[j] = ok ? idx[j] : -1;
            if (!ok) continue;
            ++K;
            const float *qf = tree->pts + 3 * (int64_t)idx[j];
            for (int k = 0; k < 3; ++k) { sp[k] += P[3 * j + k]; sq[k] += qf[k]; double e = (double)P[3 * j + k] - qf[k]; sd2 += e * e; }
        }
        prev_fit = fitness; prev_rmse = rmse;
        fitness = n_source > 0 ? (double)K / (double)n_source : 0.0;
        rmse = K > 0 ? sqrt(sd2 / (double)K) : 0.0;
        res->k_final = (int32_t)K;
        if (pass > 0 && fabs(prev_fit - fitness) < 1e-6 && fabs(prev_rmse - rmse) < 1e-6) { converged = 1; break; }
        if (pass == prm->num_iterations || K == 0) break;
        double H[9] = {0};
        for (int k = 0; k < 3; ++k) { sp[k] /= (double)K; sq[k] /= (double)K; }
        for (int64_t j = 0; j < n_source; ++j) {
            if (!(idx[j] >= 0 && d2[j] < prm->max_correspondence_dist)) continue;
            const float *qf = tree->pts + 3 * (int64_t)idx[j];
            double a[3] = {P[3 * j] - sp[0], P[3 * j + 1] - sp[1], P[3 * j + 2] - sp[2]};
            double b[3] = {qf[0] - sq[0], qf[1] - sq[1], qf[2] - sq[2]};
            for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) H[c * 3 + r] += a[r] * b[c];
        }
        double U[9], S[3], V[9], R[9];
        /* H = sum a b^T = U S V^T  ->  R = V U^T;  jacobi_svd3_d factors its argument as A V = U S */
        jacobi_svd3_d(H, U, S, V);
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) {
            double s = 0; for (int k = 0; k < 3; ++k) s += V[k * 3 + r] * U[k * 3 + c]; R[c * 3 + r] = s; }
        if (det3_d(R) < 0) {
            for (int k = 0; k < 3; ++k) V[6 + k] = -V[6 + k];
            for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) {
                double s = 0; for (int k = 0; k < 3; ++k) s += V[k * 3 + r] * U[k * 3 + c]; R[c * 3 + r] = s; }
        }
        double Ts[16];
        for (int i = 0; i < 16; ++i) Ts[i] = (i % 5 == 0) ? 1.0 : 0.0;
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) Ts[c * 4 + r] = R[c * 3 + r];
        for (int r = 0; r < 3; ++r) Ts[12 + r] = sq[r] - (R[0 * 3 + r] * sp[0] + R[1 * 3 + r] * sp[1] + R[2 * 3 + r] * sp[2]);
        compose_round(Ts, T);
        ++it;
    }
    memcpy(res->transformation, T, sizeof(T));
    res->error = (float)rmse; res->iterations = it; res->has_converged = converged;
    if (fitness_out) *fitness_out = (float)fitness;
    free(P); free(d2); free(idx);
    return 0;
}

/* ======================================================================================
 * 5. pcl::VoxelGrid<PointXYZ>::applyFilter with setLeafSize(l,l,l), as called at
 *    global_map_frames_manager.cpp:143-146:
 *    inverse leaf in float, min/max over finite points, overflow refusal (output = input),
 *    idx = (floor(p*inv) - min_b) . (1, dx, dx*dy), STABLE sort by idx (contract: in-voxel
 *    order = ascending original index), centroid = float running sum / count.
 *    Returns the number of output points; *status = 1 when the overflow guard fired.
 * ==================================================================================== */
typedef struct { int32_t idx; int32_t pt; } vox_pair_t;

static int vox_cmp(const void *a, const void *b)
{
    const vox_pair_t *x = (const vox_pair_t *)a, *y = (const vox_pair_t *)b;
    if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
    return x->pt < y->pt ? -1 : (x->pt > y->pt ? 1 : 0);
}

int64_t ssf_oracle_voxel_grid(const float *in, int64_t n, int stride, float leaf, float *out, int32_t *status)
{
    *status = 0;
    const float inv = 1.0f / leaf;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    int64_t n_finite = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = in + i * stride;
        if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
        ++n_finite;
        for (int k = 0; k < 3; ++k) { if (p[k] < mn[k]) mn[k] = p[k]; if (p[k] > mx[k]) mx[k] = p[k]; }
    }
    if (n_finite == 0) return 0;
    int64_t d64[3];
    for (int k = 0; k < 3; ++k) d64[k] = (int64_t)((mx[k] - mn[k]) * inv) + 1;
    if (d64[0] * d64[1] * d64[2] > (int64_t)INT32_MAX) {
        *status = 1;
        for (int64_t i = 0; i < n; ++i) { memcpy(out + 4 * i, in + i * stride, 3 * sizeof(float)); out[4 * i + 3] = 1.0f; }
        return n;
    }
    int32_t minb[3], maxb[3], divb[3];
    for (int k = 0; k < 3; ++k) {
        minb[k] = (int32_t)floorf(mn[k] * inv);
        maxb[k] = (int32_t)floorf(mx[k] * inv);
        divb[k] = maxb[k] - minb[k] + 1;
    }
    const int32_t mul[3] = {1, divb[0], divb[0] * divb[1]};
    vox_pair_t *pairs = (vox_pair_t *)malloc(sizeof(vox_pair_t) * (size_t)n_finite);
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = in + i * stride;
        if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
        int32_t i0 = (int32_t)(floorf(p[0] * inv) - (float)minb[0]);
        int32_t i1 = (int32_t)(floorf(p[1] * inv) - (float)minb[1]);
        int32_t i2 = (int32_t)(floorf(p[2] * inv) - (float)minb[2]);
        pairs[m].idx = i0 * mul[0] + i1 * mul[1] + i2 * mul[2];
        pairs[m].pt = (int32_t)i;
        ++m;
    }
    qsort(pairs, (size_t)m, sizeof(vox_pair_t), vox_cmp);
    int64_t n_out = 0, i = 0;
    while (i < m) {
        int64_t j = i;
        float c[3] = {0, 0, 0};
        while (j < m && pairs[j].idx == pairs[i].idx) {
            const float *p = in + (int64_t)pairs[j].pt * stride;
            c[0] += p[0]; c[1] += p[1]; c[2] += p[2];
            ++j;
        }
        const float cnt = (float)(j - i);
        out[4 * n_out + 0] = c[0] / cnt;
        out[4 * n_out + 1] = c[1] / cnt;
        out[4 * n_out + 2] = c[2] / cnt;
        out[4 * n_out + 3] = 1.0f;
        ++n_out;
        i = j;
    }
    free(pairs);
    return n_out;
}

/* open3d::geometry::PointCloud::VoxelDownSample (localization_python/.../localization_node.py:47)  [ext]:
 * origin = min_bound - voxel / 2, index = floor((p - origin) / voxel) per axis, AccumulatedPoint per voxel
 * (points added in input order), centroid = sum / count -- all in double.  Open3D iterates an
 * unordered_map for the output; the contract is ascending (z, y, x) voxel index.  Output float4 rows
 * (centroid rounded to float).  Returns the number of output points. */
typedef struct { uint64_t idx; int32_t pt; } vox64_pair_t;
static int vox64_cmp(const void *a, const void *b)
{
    const vox64_pair_t *x = (const vox64_pair_t *)a, *y = (const vox64_pair_t *)b;
    if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
    return x->pt < y->pt ? -1 : (x->pt > y->pt ? 1 : 0);
}

int64_t ssf_oracle_voxel_grid_o3d(const float *in, int64_t n, int stride, double voxel, float *out)
{
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    int64_t n_finite = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = in + i * stride;
        if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
        ++n_finite;
        for (int k = 0; k < 3; ++k) { if (p[k] < mn[k]) mn[k] = p[k]; if (p[k] > mx[k]) mx[k] = p[k]; }
    }
    if (n_finite == 0) return 0;
    double b[3];
    uint64_t d[3];
    for (int k = 0; k < 3; ++k) { b[k] = mn[k] - voxel * 0.5; d[k] = (uint64_t)floor((mx[k] - b[k]) / voxel) + 1u; }
    vox64_pair_t *pairs = (vox64_pair_t *)malloc(sizeof(vox64_pair_t) * (size_t)n_finite);
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = in + i * stride;
        if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
        const uint64_t i0 = (uint64_t)(int64_t)floor(((double)p[0] - b[0]) / voxel);
        const uint64_t i1 = (uint64_t)(int64_t)floor(((double)p[1] - b[1]) / voxel);
        const uint64_t i2 = (uint64_t)(int64_t)floor(((double)p[2] - b[2]) / voxel);
        pairs[m].idx = i0 + i1 * d[0] + i2 * d[0] * d[1];
        pairs[m].pt = (int32_t)i;
        ++m;
    }
    qsort(pairs, (size_t)m, sizeof(vox64_pair_t), vox64_cmp);
    int64_t n_out = 0, i = 0;
    while (i < m) {
        int64_t j = i;
        double c[3] = {0, 0, 0};
        while (j < m && pairs[j].idx == pairs[i].idx) {
            const float *p = in + (int64_t)pairs[j].pt * stride;
            c[0] += p[0]; c[1] += p[1]; c[2] += p[2];
            ++j;
        }
        const double cnt = (double)(j - i);
        out[4 * n_out + 0] = (float)(c[0] / cnt);
        out[4 * n_out + 1] = (float)(c[1] / cnt);
        out[4 * n_out + 2] = (float)(c[2] / cnt);
        out[4 * n_out + 3] = 1.0f;
        ++n_out;
        i = j;
    }
    free(pairs);
    return n_out;
}

/* ======================================================================================
 * 6. Cloud pre-processing, reference localization/include/localization/point_cloud_processing.hpp.
 *    Outputs are float4 rows (x, y, z, 1); each function returns the number of rows written.
 * ==================================================================================== */

/* applyUniformSubsample, hpp:55-74: indices 0, step, 2*step, ...; cloud unchanged if size < step */
int64_t ssf_oracle_subsample(const float *in, int64_t n, int stride, int64_t step, float *out)
{
    int64_t m = 0;
    if (n < step) step = 1; /* hpp:58-61 returns early: every point stays */
    for (int64_t i = 0; i < n; i += step) {
        memcpy(out + 4 * m, in + i * stride, 3 * sizeof(float));
        out[4 * m + 3] = 1.0f;
        ++m;
    }
    return m;
}

/* removeFloor, hpp:76-92: keep z > 0 */
int64_t ssf_oracle_remove_floor(const float *in, int64_t n, int stride, float *out)
{
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i)
        if (in[i * stride + 2] > 0) {
            memcpy(out + 4 * m, in + i * stride, 3 * sizeof(float));
            out[4 * m + 3] = 1.0f;
            ++m;
        }
    return m;
}

typedef struct { float d2; int32_t idx; } rad_pair_t;
static int rad_cmp(const void *a, const void *b)
{
    const rad_pair_t *x = (const rad_pair_t *)a, *y = (const rad_pair_t *)b;
    if (x->d2 != y->d2) return x->d2 < y->d2 ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

/* cropPointCloudThroughRadius, hpp:31-53: pcl::search::KdTree::radiusSearch(center, radius) then
 * ExtractIndices.  [ext] FLANN radius search keeps dist < radius^2 (strict), PCL passes
 * float(radius*radius) and asks for results sorted by distance (FLANN orders ties by index). */
int64_t ssf_oracle_crop_radius(const float *in, int64_t n, int stride, const float *center, double radius, float *out,
                               int32_t *idx_out)
{
    const float r2 = (float)(radius * radius);
    rad_pair_t *pairs = (rad_pair_t *)malloc(sizeof(rad_pair_t) * (size_t)(n > 0 ? n : 1));
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = in + i * stride;
        if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
        float d = sqdist3(center, p);
        if (d < r2) { pairs[m].d2 = d; pairs[m].idx = (int32_t)i; ++m; }
    }
    qsort(pairs, (size_t)m, sizeof(rad_pair_t), rad_cmp);
    for (int64_t j = 0; j < m; ++j) {
        memcpy(out + 4 * j, in + (int64_t)pairs[j].idx * stride, 3 * sizeof(float));
        out[4 * j + 3] = 1.0f;
        if (idx_out) idx_out[j] = pairs[j].idx;
    }
    free(pairs);
    return m;
}

/* ======================================================================================
 * 7. BruteForceAlignment::alignClouds, reference localization/src/brute_force_alignment.cpp:65-136,
 *    with the test sequences of createTestTransformSequences (cpp:148-180).
 *    [ext] Eigen pieces restated: AngleAxisf(yaw, UnitZ).toRotationMatrix() =
 *    [[c,-s,0],[s,c,0],[0,0,(1-c)+c]] in float; fixed 4x4 products accumulate k = 0..3 in order.
 * ==================================================================================== */
typedef struct {
    float x_step, y_step, z_step;
    float x_range, y_range, z_range;
    float yaw_step, yaw_range;
    float mean_error_threshold;
} ssf_oracle_bfa_params;

static int bfa_sequence(float range, float step, float *out, int cap)
{
    int n = 0;
    for (int i = 0; (float)i < range / (2 * step) + 1; ++i) { /* cpp:160-179: both signs, i = 0 twice */
        if (n + 2 > cap) break;
        out[n++] = -i * step;
        out[n++] = i * step;
    }
    return n;
}

/* Number of candidate poses and (optionally) their transforms prev * T(x, y, z, yaw) in loop order. */
int64_t ssf_oracle_bfa_poses(const float *T_prev, const ssf_oracle_bfa_params *p, float *T_out /* n x 16 or NULL */)
{
    float xs[4096], ys[4096], zs[4096], ws[4096];
    int nx = bfa_sequence(p->x_range, p->x_step, xs, 4096), ny = bfa_sequence(p->y_range, p->y_step, ys, 4096);
    int nz = bfa_sequence(p->z_range, p->z_step, zs, 4096), nw = bfa_sequence(p->yaw_range, p->yaw_step, ws, 4096);
    int64_t n = 0;
    for (int a = 0; a < nx; ++a)
        for (int b = 0; b < ny; ++b)
            for (int c = 0; c < nz; ++c)
                for (int d = 0; d < nw; ++d) { /* cpp:80-92 */
                    if (T_out) {
                        float T[16];
                        for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
                        const float cs = cosf(ws[d]), sn = sinf(ws[d]);
                        M4(T, 0, 0) = cs; M4(T, 0, 1) = -sn; M4(T, 1, 0) = sn; M4(T, 1, 1) = cs;
                        M4(T, 2, 2) = (1.f - cs) + cs;
                        M4(T, 0, 3) = xs[a]; M4(T, 1, 3) = ys[b]; M4(T, 2, 3) = zs[c];
                        mat4_mul_f(T_prev, T, T_out + 16 * n);
                    }
                    ++n;
                }
    return n;
}

/* scores_out (optional): mean squared NN distance of every candidate that was evaluated, NaN for
 * the ones skipped by the early return.  no_early_exit != 0 evaluates all of them (for tests). */
int ssf_oracle_bfa_align(const void *tree_v, const float *src, int64_t n_src, int stride, const float *T_prev,
                         const ssf_oracle_bfa_params *p, int no_early_exit, float *T_best_out, float *best_score_out,
                         int32_t *success_out, float *scores_out, int threads)
{
    const kdtree_t *tree = (const kdtree_t *)tree_v;
    const int64_t n_pose = ssf_oracle_bfa_poses(T_prev, p, NULL);
    float *Ts = (float *)malloc(sizeof(float) * 16 * (size_t)n_pose);
    ssf_oracle_bfa_poses(T_prev, p, Ts);
    float *q = (float *)malloc(sizeof(float) * 3 * (size_t)(n_src > 0 ? n_src : 1));
    float *d2 = (float *)malloc(sizeof(float) * (size_t)(n_src > 0 ? n_src : 1));
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_src > 0 ? n_src : 1));
    float best_score = FLT_MAX, best_T[16];
    for (int i = 0; i < 16; ++i) best_T[i] = (i % 5 == 0) ? 1.f : 0.f; /* cpp:68 */
    int success = 0, decided = 0;
    if (scores_out) for (int64_t k = 0; k < n_pose; ++k) scores_out[k] = NAN;
    for (int64_t k = 0; k < n_pose; ++k) {
        const float *T = Ts + 16 * k;
        transform_all(T, src, n_src, stride, q);                           /* cpp:98-99 */
        ssf_oracle_kdtree_nn(tree, q, n_src, 3, idx, d2, threads);         /* cpp:100-102, unbounded */
        float score = 0.0f;
        for (int64_t i = 0; i < n_src; ++i) score += d2[i];                /* cpp:103 */
        score /= (float)n_src;                                             /* cpp:105 */
        if (scores_out) scores_out[k] = score;
        if (decided) continue;
        if (score < best_score) { best_score = score; memcpy(best_T, T, sizeof(best_T)); } /* cpp:108-112 */
        if (score < p->mean_error_threshold) {                             /* cpp:114-119 */
            memcpy(best_T, T, sizeof(best_T));
            success = 1;
            decided = 1;
            if (!no_early_exit) break;
        }
    }
    if (!success && best_score < p->mean_error_threshold) success = 1;     /* cpp:128-134 (unreachable in practice) */
    memcpy(T_best_out, best_T, sizeof(best_T));
    *best_score_out = best_score;
    *success_out = success;
    free(Ts); free(q); free(d2); free(idx);
    return 0;
}
