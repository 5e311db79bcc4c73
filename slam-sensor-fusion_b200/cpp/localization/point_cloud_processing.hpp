// point_cloud_processing.hpp -- drop-in replacement for the reference header of the same name
// (viniciusvidal2/slam-sensor-fusion, localization/include/localization/point_cloud_processing.hpp:31-92).
//
// The same three free functions with the same signatures; callers localization_node.cpp:20, 211-213, 292,
// 296, 302 compile unchanged.  Each runs on the GPU through libssf_gpu.so (ssf_cloud_crop_radius /
// ssf_cloud_subsample / ssf_cloud_remove_floor, include/ssf/ssf.h) and returns exactly the cloud the
// reference returns: the crop in pcl::search::KdTree::radiusSearch order (ascending distance, equal
// distances by index) without building a KD-tree over the cloud first (hpp:37-45).
//
// Beyond the reference: ssf::ResidentMap keeps the whole map cloud in HBM, so the re-crop of
// localization_node.cpp:300-305 (crop + setTargetPointCloud) becomes a device-side window change:
//     resident_map.cropToTarget(map_T_sensor_, 10.0, *icp_);     // instead of lines :302-303
#ifndef POINT_CLOUD_PROCESSING_H
#define POINT_CLOUD_PROCESSING_H

#include <ssf/ssf.h>

#include <cstdlib>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#ifndef SSF_SHIM_STANDALONE
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include <Eigen/Core>
#else
#include <localization/ssf_standalone_types.h>
#endif

using PointT = pcl::PointXYZ;

namespace ssf {
/// One context per process for the free functions below (created on first use, never destroyed: the
/// functions may be called from static destructors of the node).
inline ssf_ctx *sharedContext()
{
    static ssf_ctx *ctx = [] {
        ssf_ctx *c = nullptr;
        const char *dev = std::getenv("SSF_DEVICE");
        if (ssf_ctx_create(dev ? std::atoi(dev) : 0, &c) != SSF_OK)
            std::cerr << "[SSF ERROR] libssf_gpu context creation failed: " << ssf_last_error() << std::endl;
        return c;
    }();
    return ctx;
}
inline void storeCloud(const std::vector<float> &xyzw, std::size_t n, pcl::PointCloud<pcl::PointXYZ> &cloud)
{
    cloud.points.resize(n);
    for (std::size_t i = 0; i < n; ++i) cloud.points[i] = pcl::PointXYZ(xyzw[4 * i], xyzw[4 * i + 1], xyzw[4 * i + 2]);
    cloud.width = static_cast<std::uint32_t>(n);
    cloud.height = 1;
    cloud.is_dense = true;
}
}  // namespace ssf

static inline void cropPointCloudThroughRadius(const Eigen::Matrix4f &T, const double radius,
                                               pcl::PointCloud<PointT>::Ptr &cloud,
                                               pcl::PointCloud<PointT>::Ptr &cropped_cloud)
{
    const std::size_t n = cloud->points.size();
    const float center[3] = {T(0, 3), T(1, 3), T(2, 3)};
    std::vector<float> out(4 * (n ? n : 1));
    std::size_t n_out = 0;
    ssf_ctx *ctx = ssf::sharedContext();
    if (!ctx || ssf_cloud_crop_radius(ctx, n ? &cloud->points[0].x : nullptr, n, sizeof(PointT), center, radius, out.data(),
                                      &n_out, nullptr) != SSF_OK) {
        std::cerr << "[SSF ERROR] cropPointCloudThroughRadius: " << ssf_last_error() << std::endl;
        return;
    }
    ssf::storeCloud(out, n_out, *cropped_cloud);
}

static inline void applyUniformSubsample(pcl::PointCloud<PointT>::Ptr &cloud, const std::size_t point_step)
{
    const std::size_t n = cloud->points.size();
    if (n < point_step) return;  // hpp:58-61
    std::vector<float> out(4 * (n ? n : 1));
    std::size_t n_out = 0;
    ssf_ctx *ctx = ssf::sharedContext();
    if (!ctx || ssf_cloud_subsample(ctx, n ? &cloud->points[0].x : nullptr, n, sizeof(PointT), point_step, out.data(),
                                    &n_out) != SSF_OK) {
        std::cerr << "[SSF ERROR] applyUniformSubsample: " << ssf_last_error() << std::endl;
        return;
    }
    ssf::storeCloud(out, n_out, *cloud);
}

static inline void removeFloor(pcl::PointCloud<PointT>::Ptr &cloud)
{
    const std::size_t n = cloud->points.size();
    std::vector<float> out(4 * (n ? n : 1));
    std::size_t n_out = 0;
    ssf_ctx *ctx = ssf::sharedContext();
    if (!ctx || ssf_cloud_remove_floor(ctx, n ? &cloud->points[0].x : nullptr, n, sizeof(PointT), out.data(), &n_out) != SSF_OK) {
        std::cerr << "[SSF ERROR] removeFloor: " << ssf_last_error() << std::endl;
        return;
    }
    ssf::storeCloud(out, n_out, *cloud);
}

namespace ssf {
/// The map cloud resident in HBM (ssf_map): uploaded (or merged from the recorder's PCD tiles) once.
class ResidentMap
{
public:
    ResidentMap() = default;
    ~ResidentMap() { reset(); }
    ResidentMap(const ResidentMap &) = delete;
    ResidentMap &operator=(const ResidentMap &) = delete;
    /// upload a cloud the node already holds (map_cloud_ after localization_node.cpp:19-20)
    bool upload(const pcl::PointCloud<pcl::PointXYZ>::Ptr &cloud)
    {
        reset();
        const std::size_t n = cloud->points.size();
        return check(ssf_map_create(sharedContext(), n ? &cloud->points[0].x : nullptr, n, sizeof(pcl::PointXYZ), &map_),
                     "ResidentMap::upload");
    }
    /// GlobalMapFramesManager::getMapCloud(voxel_size) (global_map_frames_manager.cpp:93-151) straight into HBM
    bool loadFolder(const std::string &data_folder, const std::string &map_name, float voxel_size, bool save = true)
    {
        reset();
        return check(ssf_map_from_pcd_folder(sharedContext(), data_folder.c_str(), map_name.c_str(), voxel_size, save ? 1 : 0,
                                             &map_),
                     "ResidentMap::loadFolder");
    }
    bool subsample(std::size_t point_step) { return map_ && check(ssf_map_subsample(map_, point_step), "ResidentMap::subsample"); }
    std::size_t size() const { return ssf_map_size(map_); }
    /// cropPointCloudThroughRadius on the resident cloud; the result comes back to the host (debug topics)
    bool crop(const Eigen::Matrix4f &T, double radius, pcl::PointCloud<pcl::PointXYZ>::Ptr &cropped_cloud)
    {
        if (!map_) return false;
        const float center[3] = {T(0, 3), T(1, 3), T(2, 3)};
        std::vector<float> out(4 * (size() ? size() : 1));
        std::size_t n_out = 0;
        if (!check(ssf_map_crop_radius(map_, center, radius, out.data(), size(), &n_out, nullptr), "ResidentMap::crop")) return false;
        storeCloud(out, n_out, *cropped_cloud);
        return true;
    }
    /// crop + setTargetPointCloud (localization_node.cpp:302-303) without leaving HBM.  `icp_handle` is
    /// ICPPointToPoint::handle() of the drop-in class.
    bool cropToTarget(const Eigen::Matrix4f &T, double radius, ssf_icp *icp_handle, std::size_t *n_out = nullptr)
    {
        if (!map_ || !icp_handle) return false;
        const float center[3] = {T(0, 3), T(1, 3), T(2, 3)};
        return check(ssf_map_crop_to_target(map_, icp_handle, center, radius, n_out), "ResidentMap::cropToTarget");
    }
    ssf_map *handle() const { return map_; }

private:
    void reset()
    {
        if (map_) ssf_map_destroy(map_);
        map_ = nullptr;
    }
    static bool check(int rc, const char *what)
    {
        if (rc != SSF_OK) std::cerr << "[SSF ERROR] " << what << ": " << ssf_last_error() << std::endl;
        return rc == SSF_OK;
    }
    ssf_map *map_ = nullptr;
};
}  // namespace ssf

#endif  // POINT_CLOUD_PROCESSING_H
