// bfa.cuh -- brute-force pose-grid alignment (see bfa.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace ssf {

struct BfaWork {
    DevBuf<float4> src;
    DevBuf<float> poses, d2, score;
};

int bfa_poses_host(const float *T_prev, const ssf_bfa_params &p, std::vector<float> &poses);
int bfa_scores_device(const MapView &map, BfaWork &w, size_t n_src, const std::vector<float> &poses,
                      std::vector<float> &scores, cudaStream_t st);

}  // namespace ssf
