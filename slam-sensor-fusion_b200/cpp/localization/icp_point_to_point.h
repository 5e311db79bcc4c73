// icp_point_to_point.h -- drop-in replacement for the reference header of the same name
// (viniciusvidal2/slam-sensor-fusion, localization/include/localization/icp_point_to_point.h).
//
// Same type names, constructor, setters and calculateAlignment() as the reference class
// (icp_point_to_point.h:28-85), so localization_node.cpp (:28-29, :223-236, :303, :335-338)
// compiles and behaves unchanged -- but every call goes through the C ABI of libssf_gpu.so
// (include/ssf/ssf.h) to sm_100a kernels; there is no CPU path in here.
//
// With PCL and Eigen present (a ROS 2 workspace) the real pcl::PointCloud<pcl::PointXYZ>::Ptr
// and Eigen::Matrix4f are used.  Without them (this repository's tests) define
// SSF_SHIM_STANDALONE before including: minimal stand-ins with the same memory layout
// (16-byte xyz_ points, column-major 4x4 float) take their place.
#ifndef ICP_POINT_TO_POINT_H
#define ICP_POINT_TO_POINT_H

#include <ssf/ssf.h>

#include <cstdio>
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

// Result record with the reference's field names, defaults and constructors
// (icp_point_to_point.h:28-39): identity transform, error 1e6, 0 iterations, not converged.
struct ICPResult {
    Eigen::Matrix4f transformation = Eigen::Matrix4f::Identity();
    float error = 1e6f;
    int iterations = 0;
    bool has_converged = false;

    ICPResult() = default;
    ICPResult(const Eigen::Matrix4f &T) : transformation(T) {}  // implicit, like the reference (h:32)
    ICPResult(const Eigen::Matrix4f &T, float err, int its, bool converged)
        : transformation(T), error(err), iterations(its), has_converged(converged)
    {
    }
};

class ICPPointToPoint
{
public:
    /// Same arguments as the reference constructor (icp_point_to_point.cpp:3-12).
    ICPPointToPoint(const float max_correspondence_dist, const int num_iterations, const float acceptable_mean_error,
                    const float transformation_epsilon)
    {
        params_.max_correspondence_dist = max_correspondence_dist;
        params_.num_iterations = num_iterations;
        params_.acceptable_mean_error = acceptable_mean_error;
        params_.transformation_epsilon = transformation_epsilon;
        params_.mode = envInt("SSF_MODE", SSF_MODE_REFERENCE);
        params_.reduce = envInt("SSF_REDUCE", SSF_REDUCE_STRICT);
        params_.debug = 0;  // the reference leaves debug_mode_ uninitialised (icp_point_to_point.h:135)
        params_.source_voxel_leaf = 0.f;
        initial_transform_ = Eigen::Matrix4f::Identity();
        if (ssf_ctx_create(envInt("SSF_DEVICE", 0), &ctx_) != SSF_OK ||
            ssf_icp_create(ctx_, &params_, &icp_) != SSF_OK)
            fail("construction");
    }
    ~ICPPointToPoint()
    {
        if (icp_) ssf_icp_destroy(icp_);
        if (ctx_) ssf_ctx_destroy(ctx_);
    }
    ICPPointToPoint(const ICPPointToPoint &) = delete;
    ICPPointToPoint &operator=(const ICPPointToPoint &) = delete;

    void setMaxCorrespondenceDist(const float max_correspondence_dist)
    {
        params_.max_correspondence_dist = max_correspondence_dist;
        push();
    }
    void setNumIterations(const int num_iterations)
    {
        params_.num_iterations = num_iterations;
        push();
    }
    void setTransformationEpsilon(const float transformation_epsilon)
    {
        params_.transformation_epsilon = transformation_epsilon;
        push();
    }
    void setAcceptableMeanError(const float acceptable_error)
    {
        params_.acceptable_mean_error = acceptable_error;
        push();
    }
    void setInitialTransformation(const Eigen::Matrix4f &initial_transformation)
    {
        initial_transform_ = initial_transformation;
        if (icp_ && ssf_icp_set_initial(icp_, initial_transform_.data()) != SSF_OK) fail("setInitialTransformation");
    }
    void setSourcePointCloud(const pcl::PointCloud<PointT>::Ptr &source_cloud)
    {
        const float *p = source_cloud->points.empty() ? nullptr : &source_cloud->points[0].x;
        if (icp_ && ssf_icp_set_source(icp_, p, source_cloud->points.size(), sizeof(PointT)) != SSF_OK)
            fail("setSourcePointCloud");
    }
    void setTargetPointCloud(const pcl::PointCloud<PointT>::Ptr &target_cloud)
    {
        const float *p = target_cloud->points.empty() ? nullptr : &target_cloud->points[0].x;
        if (icp_ && ssf_icp_set_target(icp_, p, target_cloud->points.size(), sizeof(PointT), nullptr, 0) != SSF_OK)
            fail("setTargetPointCloud");
    }
    void setDebugMode(bool debug_mode)
    {
        params_.debug = debug_mode ? 1 : 0;
        push();
    }

    /// calculateAlignment (icp_point_to_point.cpp:185-254).  Never throws; a device failure
    /// yields the same sentinel result the reference returns for "not enough correspondences".
    ICPResult calculateAlignment()
    {
        ICPResult out(initial_transform_);
        ssf_icp_result r;
        if (!icp_ || ssf_icp_align(icp_, &r) != SSF_OK) {
            fail("calculateAlignment");
            return out;
        }
        last_ = r;
        if (r.aborted) return out;  // {T = initial, error = 1e6, iterations = 0, has_converged = false}
        std::memcpy(out.transformation.data(), r.transformation, sizeof(r.transformation));
        out.error = r.error;
        out.iterations = r.iterations;
        out.has_converged = r.has_converged != 0;
        return out;
    }

    /// Extras the reference does not have: diagnostics of the last alignment, and the C handle (for
    /// ssf::ResidentMap::cropToTarget and ssf_batch_*).  An ICPPointToPoint and the free functions of
    /// point_cloud_processing.hpp may sit on different contexts of the same device.
    const ssf_icp_result &lastDeviceResult() const { return last_; }
    ssf_icp *handle() const { return icp_; }

private:
    static int envInt(const char *name, int dflt)
    {
        const char *v = std::getenv(name);
        return v ? std::atoi(v) : dflt;
    }
    void push()
    {
        if (icp_ && ssf_icp_set_params(icp_, &params_) != SSF_OK) fail("set_params");
    }
    void fail(const char *what) const
    {
        std::cerr << "[ICP ERROR] libssf_gpu " << what << " failed: " << ssf_last_error() << std::endl;
    }

    ssf_ctx *ctx_ = nullptr;
    ssf_icp *icp_ = nullptr;
    ssf_icp_params params_{};
    ssf_icp_result last_{};
    Eigen::Matrix4f initial_transform_;
};

#endif  // ICP_POINT_TO_POINT_H
