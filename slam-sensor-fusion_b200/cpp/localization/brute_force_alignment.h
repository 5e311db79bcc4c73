// brute_force_alignment.h -- drop-in replacement for the reference header of the same name
// (viniciusvidal2/slam-sensor-fusion, localization/include/localization/brute_force_alignment.h:1-115).
//
// Same class name, setters, alignClouds(), firstAlignmentCompleted() and getBestTransformation() as the
// reference (brute_force_alignment.cpp:3-146), so localization_node.cpp (:38-43 parameters, :216-219 and
// :245 use) compiles and behaves unchanged.  The 4-deep loop over candidate poses and the per-point
// unbounded nearest-neighbour search (cpp:80-119) run on the GPU through ssf_bfa_align of libssf_gpu.so
// (include/ssf/ssf.h); scores, the first-below-threshold rule and the best-so-far carry-over are
// bit-identical to the CPU loop (tests/test_bfa.py, tests/test_ref_pin.py).  No CPU path in here.
#ifndef BRUTE_FORCE_ALIGNMENT_H
#define BRUTE_FORCE_ALIGNMENT_H

#include <ssf/ssf.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <vector>

#ifndef SSF_SHIM_STANDALONE
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include <Eigen/Core>
#else
#include <localization/ssf_standalone_types.h>
#endif

using PointT = pcl::PointXYZ;

class BruteForceAlignment
{
public:
    BruteForceAlignment()
    {
        map_T_sensor_previous_ = Eigen::Matrix4f::Identity();
        map_T_sensor_best_ = Eigen::Matrix4f::Identity();
        source_cloud_ = pcl::PointCloud<PointT>::Ptr(new pcl::PointCloud<PointT>());
        target_cloud_ = pcl::PointCloud<PointT>::Ptr(new pcl::PointCloud<PointT>());
        // reference defaults (brute_force_alignment.h:93-101)
        params_.x_step = params_.y_step = params_.z_step = 0.1f;
        params_.yaw_step = (float)(M_PI / 90.0f);
        params_.x_range = params_.y_range = params_.z_range = 0.5f;
        params_.yaw_range = (float)(M_PI / 6.0f);
        params_.mean_error_threshold = 0.1f;
        ssf_icp_params p{};
        p.max_correspondence_dist = 1.f;
        p.num_iterations = 1;
        const char *dev = std::getenv("SSF_DEVICE");
        if (ssf_ctx_create(dev ? std::atoi(dev) : 0, &ctx_) != SSF_OK || ssf_icp_create(ctx_, &p, &icp_) != SSF_OK)
            fail("construction");
    }
    ~BruteForceAlignment()
    {
        if (icp_) ssf_icp_destroy(icp_);
        if (ctx_) ssf_ctx_destroy(ctx_);
    }
    BruteForceAlignment(const BruteForceAlignment &) = delete;
    BruteForceAlignment &operator=(const BruteForceAlignment &) = delete;

    void setXYZStep(const float x_step, const float y_step, const float z_step)
    {
        params_.x_step = x_step;
        params_.y_step = y_step;
        params_.z_step = z_step;
    }
    void setXYZRange(const float x, const float y, const float z)
    {
        params_.x_range = x;
        params_.y_range = y;
        params_.z_range = z;
    }
    void setRotationStep(const float yaw_step) { params_.yaw_step = yaw_step; }
    void setRotationRange(const float yaw) { params_.yaw_range = yaw; }
    void setMeanErrorThreshold(const float error_threshold) { params_.mean_error_threshold = error_threshold; }
    void setSourceCloud(const pcl::PointCloud<PointT>::Ptr &cloud) { source_cloud_ = cloud; }
    /// The target's search index is built here (the reference builds its KD-tree inside every alignClouds
    /// call, cpp:72-73; the node sets the target once per attempt, localization_node.cpp:217).
    void setTargetCloud(const pcl::PointCloud<PointT>::Ptr &cloud)
    {
        target_cloud_ = cloud;
        target_dirty_ = true;
    }
    /// Only the first guess is taken (cpp:44-51: the trace() == 4.0f test).
    void setInitialGuess(const Eigen::Matrix4f &initial_guess)
    {
        const Eigen::Matrix4f &P = map_T_sensor_previous_;
        if ((P(0, 0) + P(1, 1)) + (P(2, 2) + P(3, 3)) == 4.0f) map_T_sensor_previous_ = initial_guess;
    }
    void resetFirstAlignment(const bool value) { first_alignment_completed_ = value; }

    /// alignClouds (cpp:65-136).  Never throws; a device failure reports false and keeps the state.
    bool alignClouds()
    {
        if (!icp_) return false;
        if (target_dirty_) {
            const float *t = target_cloud_->points.empty() ? nullptr : &target_cloud_->points[0].x;
            if (ssf_icp_set_target(icp_, t, target_cloud_->points.size(), sizeof(PointT), nullptr, 0) != SSF_OK) {
                fail("setTargetCloud");
                return false;
            }
            target_dirty_ = false;
        }
        if (source_cloud_->points.empty()) return false;  // the reference divides by zero here (cpp:105)
        Eigen::Matrix4f best = Eigen::Matrix4f::Identity();
        float score = 0.f;
        int ok = 0;
        if (ssf_bfa_align(icp_, &source_cloud_->points[0].x, source_cloud_->points.size(), sizeof(PointT),
                          map_T_sensor_previous_.data(), &params_, best.data(), &score, &ok, nullptr) != SSF_OK) {
            fail("alignClouds");
            return false;
        }
        last_score_ = score;
        if (ok) {  // cpp:114-119 (and :128-134)
            map_T_sensor_best_ = best;
            first_alignment_completed_ = true;
            return true;
        }
        map_T_sensor_previous_ = best;  // cpp:126: the best candidate is the next starting pose
        return false;
    }
    bool firstAlignmentCompleted() const { return first_alignment_completed_; }
    Eigen::Matrix4f getBestTransformation() const
    {
        return first_alignment_completed_ ? map_T_sensor_best_ : map_T_sensor_previous_;
    }
    /// Extra: best mean squared distance of the last alignClouds call.
    float lastBestScore() const { return last_score_; }

private:
    void fail(const char *what) const
    {
        std::cerr << "[BFA ERROR] libssf_gpu " << what << " failed: " << ssf_last_error() << std::endl;
    }
    bool first_alignment_completed_{false};
    bool target_dirty_{true};
    ssf_bfa_params params_{};
    pcl::PointCloud<PointT>::Ptr source_cloud_;
    pcl::PointCloud<PointT>::Ptr target_cloud_;
    Eigen::Matrix4f map_T_sensor_previous_;
    Eigen::Matrix4f map_T_sensor_best_;
    float last_score_{0.f};
    ssf_ctx *ctx_ = nullptr;
    ssf_icp *icp_ = nullptr;
};

#endif  // BRUTE_FORCE_ALIGNMENT_H
