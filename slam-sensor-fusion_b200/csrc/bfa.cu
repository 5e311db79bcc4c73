// bfa.cu -- row N1: BruteForceAlignment::alignClouds on the device
// (reference localization/src/brute_force_alignment.cpp:65-136, sequences :148-180).
//
// The reference scores every pose of a (x, y, z, yaw) grid -- 18*18*4*6 = 7776 with the node's
// settings (localization_node.cpp:38-43) -- by the mean squared distance of every source point
// to its (unbounded) nearest map point, one KD-tree query at a time.  Here:
//   host    the candidate transforms prev * T(x,y,z,yaw), in the reference's loop order and float
//           arithmetic (glibc cosf/sinf like the reference build);
//   K-bfa1  one thread per (pose, point): transform, unbounded exact NN (the search radius doubles
//           until a point is found), d2 -> matrix [point][pose];
//   K-bfa2  one thread per pose: score = sequential float sum over points (cpp:103), so the
//           scores -- and therefore the first-below-threshold / best-so-far decisions taken on
//           the host in loop order -- are bit-identical to the CPU loop.
#include <cmath>
#include <vector>

#include "bfa.cuh"
#include "nn_device.cuh"

namespace ssf {

__device__ __forceinline__ float nn_unbounded_d2(const MapView &m, float px, float py, float pz)
{
    if (m.n_pts == 0) return FLT_MAX;  // empty map
    const float h = __frcp_rn(m.inv_h);
    float limit = __fmul_rn(__fmul_rn(h, h), 4.0f);
    while (true) {  // a non-empty map always answers once the limit passes the true distance
        const NNHit hit = nn_query(m, px, py, pz, limit);
        if (hit.idx >= 0) return hit.d2;
        if (!(limit < 1e30f)) return FLT_MAX;
        limit = __fmul_rn(limit, 4.0f);
    }
}

__global__ void __launch_bounds__(128)
    bfa_d2_kernel(MapView map, const float4 *__restrict__ src, uint32_t pt0, uint32_t n_pts_chunk,
                  const float *__restrict__ poses, uint32_t n_pose, float *__restrict__ d2)
{
    const uint32_t pose = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pt = blockIdx.y;
    if (pose >= n_pose || pt >= n_pts_chunk) return;
    const float4 s = src[pt0 + pt];
    const float *T = poses + 16 * (size_t)pose;
    const float3 p = transform_point(T, s.x, s.y, s.z);
    d2[(size_t)pt * n_pose + pose] = nn_unbounded_d2(map, p.x, p.y, p.z);
}

__global__ void __launch_bounds__(128) bfa_sum_kernel(const float *__restrict__ d2, uint32_t n_pts_chunk,
                                                      uint32_t n_pose, float *__restrict__ score)
{
    const uint32_t pose = blockIdx.x * blockDim.x + threadIdx.x;
    if (pose >= n_pose) return;
    float s = score[pose];
    for (uint32_t pt = 0; pt < n_pts_chunk; ++pt) s = __fadd_rn(s, d2[(size_t)pt * n_pose + pose]);
    score[pose] = s;
}

static int sequence(float range, float step, std::vector<float> &out)
{
    out.clear();
    if (!(step > 0.f) || !(range >= 0.f) || range / step > 4000.f) {
        set_error("brute-force alignment: invalid step/range (%g / %g)", step, range);
        return SSF_ERR_INVALID;
    }
    for (int i = 0; (float)i < range / (2 * step) + 1; ++i) {  // cpp:160-179: both signs, i = 0 twice
        out.push_back(-i * step);
        out.push_back(i * step);
    }
    return SSF_OK;
}

static void mat4_mul_host(const float *A, const float *B, float *C)
{
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            float s = A[0 * 4 + r] * B[c * 4 + 0];
            s += A[1 * 4 + r] * B[c * 4 + 1];
            s += A[2 * 4 + r] * B[c * 4 + 2];
            s += A[3 * 4 + r] * B[c * 4 + 3];
            C[c * 4 + r] = s;
        }
}

int bfa_poses_host(const float *T_prev, const ssf_bfa_params &p, std::vector<float> &poses)
{
    std::vector<float> xs, ys, zs, ws;
    SSF_TRY(sequence(p.x_range, p.x_step, xs));
    SSF_TRY(sequence(p.y_range, p.y_step, ys));
    SSF_TRY(sequence(p.z_range, p.z_step, zs));
    SSF_TRY(sequence(p.yaw_range, p.yaw_step, ws));
    poses.clear();
    poses.reserve(16 * xs.size() * ys.size() * zs.size() * ws.size());
    for (float x : xs)
        for (float y : ys)
            for (float z : zs)
                for (float yaw : ws) {  // cpp:80-92
                    float T[16], out[16];
                    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
                    const float cs = cosf(yaw), sn = sinf(yaw);
                    T[0] = cs; T[4] = -sn; T[1] = sn; T[5] = cs;
                    T[10] = (1.f - cs) + cs;  // Eigen AngleAxis::toRotationMatrix diagonal
                    T[12] = x; T[13] = y; T[14] = z;
                    mat4_mul_host(T_prev, T, out);
                    poses.insert(poses.end(), out, out + 16);
                }
    return SSF_OK;
}

int bfa_scores_device(const MapView &map, BfaWork &w, size_t n_src, const std::vector<float> &poses,
                      std::vector<float> &scores, cudaStream_t st)
{
    const size_t n_pose = poses.size() / 16;
    scores.assign(n_pose, 0.f);
    if (n_pose == 0) return SSF_OK;
    SSF_TRY(w.poses.reserve(poses.size()));
    SSF_TRY(w.score.reserve(n_pose));
    SSF_CUDA(cudaMemcpyAsync(w.poses.p, poses.data(), poses.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    SSF_CUDA(cudaMemsetAsync(w.score.p, 0, n_pose * sizeof(float), st));
    size_t chunk = (size_t)(256u << 20) / (n_pose * sizeof(float));  // d2 matrix <= 256 MB
    if (chunk < 1) chunk = 1;
    if (chunk > 32768) chunk = 32768;
    if (chunk > n_src) chunk = n_src ? n_src : 1;
    SSF_TRY(w.d2.reserve(chunk * n_pose));
    for (size_t p0 = 0; p0 < n_src; p0 += chunk) {
        const uint32_t m = (uint32_t)((n_src - p0) < chunk ? (n_src - p0) : chunk);
        dim3 grid((unsigned)((n_pose + 127) / 128), m);
        bfa_d2_kernel<<<grid, 128, 0, st>>>(map, w.src.p, (uint32_t)p0, m, w.poses.p, (uint32_t)n_pose, w.d2.p);
        SSF_LAUNCHED();
        g_queries.fetch_add((uint64_t)m * n_pose, std::memory_order_relaxed);
        bfa_sum_kernel<<<(unsigned)((n_pose + 127) / 128), 128, 0, st>>>(w.d2.p, m, (uint32_t)n_pose, w.score.p);
        SSF_LAUNCHED();
    }
    SSF_CUDA(cudaMemcpyAsync(scores.data(), w.score.p, n_pose * sizeof(float), cudaMemcpyDeviceToHost, st));
    SSF_CUDA(cudaStreamSynchronize(st));
    for (size_t k = 0; k < n_pose; ++k) scores[k] = scores[k] / (float)n_src;  // cpp:105
    return SSF_OK;
}

}  // namespace ssf
