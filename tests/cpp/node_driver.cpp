// node_driver.cpp -- drives the three C++ drop-in headers the way localization_node.cpp drives the
// reference's (reference localization/src/localization_node.cpp: constructor :19-43, performCoarseAlignment
// :200-261, localizationCallback :290-338) and writes every intermediate cloud size / result to a file that
// tests/test_cpp_node_flow.py compares with the oracle and with oracle/_ref.
//   node_driver <map.f32> <scan.f32> <prior.f32 (16, column-major)> <out.f32>
#define SSF_SHIM_STANDALONE
#include <localization/brute_force_alignment.h>
#include <localization/icp_point_to_point.h>
#include <localization/point_cloud_processing.hpp>

#include <cstdio>
#include <string>
#include <vector>

static pcl::PointCloud<PointT>::Ptr load_cloud(const char *path)
{
    auto cloud = std::make_shared<pcl::PointCloud<PointT>>();
    FILE *f = std::fopen(path, "rb");
    if (!f) { std::perror(path); std::exit(2); }
    float v[4];
    while (std::fread(v, sizeof(float), 4, f) == 4) cloud->points.emplace_back(v[0], v[1], v[2]);
    std::fclose(f);
    return cloud;
}

static void put(std::vector<float> &out, const Eigen::Matrix4f &T)
{
    for (int i = 0; i < 16; ++i) out.push_back(T.data()[i]);
}
static void put(std::vector<float> &out, const ICPResult &r)
{
    put(out, r.transformation);
    out.push_back(r.error);
    out.push_back((float)r.iterations);
    out.push_back(r.has_converged ? 1.f : 0.f);
}
static float checksum(const pcl::PointCloud<PointT> &c)
{
    double s = 0;
    for (std::size_t i = 0; i < c.points.size(); ++i) s += (double)(i % 97 + 1) * (c.points[i].x + 2.0 * c.points[i].y + 3.0 * c.points[i].z);
    return (float)s;
}

int main(int argc, char **argv)
{
    if (argc < 5) { std::fprintf(stderr, "usage: %s map scan prior out\n", argv[0]); return argc == 2 && std::string(argv[1]) == "--help" ? 0 : 2; }
    auto map_cloud_ = load_cloud(argv[1]);
    auto scan_cloud = load_cloud(argv[2]);
    Eigen::Matrix4f map_T_sensor_ = Eigen::Matrix4f::Identity();
    {
        FILE *f = std::fopen(argv[3], "rb");
        if (!f || std::fread(map_T_sensor_.data(), sizeof(float), 16, f) != 16) { std::perror(argv[3]); return 2; }
        std::fclose(f);
    }
    std::vector<float> out;
    // ---- constructor, localization_node.cpp:19-43 -------------------------------------------------------
    ssf::ResidentMap resident;                    // beyond the reference: the map also stays in HBM
    resident.upload(map_cloud_);
    applyUniformSubsample(map_cloud_, 3);         // :20
    resident.subsample(3);
    out.push_back((float)map_cloud_->size());
    out.push_back((float)resident.size());
    auto icp_ = std::make_shared<ICPPointToPoint>(0.5f, 10, 0.05f, 1e-5f);  // :24-29
    icp_->setDebugMode(false);
    auto brute_force_alignment_ = std::make_shared<BruteForceAlignment>();   // :38-43 (ranges reduced 3x to keep the test short)
    brute_force_alignment_->setMeanErrorThreshold(0.1f);
    brute_force_alignment_->setXYZStep(0.1f, 0.1f, 0.05f);
    brute_force_alignment_->setXYZRange(0.5f, 0.5f, 0.1f);
    brute_force_alignment_->setRotationStep((float)(M_PI / 18.0f));
    brute_force_alignment_->setRotationRange((float)(M_PI / 6.0f));
    // ---- callback, :290-305 ---------------------------------------------------------------------------------
    applyUniformSubsample(scan_cloud, 2);                                                      // :292
    pcl::PointCloud<PointT>::Ptr cropped_scan_cloud = std::make_shared<pcl::PointCloud<PointT>>();
    cropPointCloudThroughRadius(Eigen::Matrix4f::Identity(), 10.0, scan_cloud, cropped_scan_cloud);  // :295-296
    out.push_back((float)cropped_scan_cloud->size());
    out.push_back(checksum(*cropped_scan_cloud));
    pcl::PointCloud<PointT>::Ptr ref_cropped_map_cloud_ = std::make_shared<pcl::PointCloud<PointT>>();
    cropPointCloudThroughRadius(map_T_sensor_, 10.0, map_cloud_, ref_cropped_map_cloud_);      // :302
    icp_->setTargetPointCloud(ref_cropped_map_cloud_);                                         // :303
    out.push_back((float)ref_cropped_map_cloud_->size());
    out.push_back(checksum(*ref_cropped_map_cloud_));
    // ---- performCoarseAlignment, :200-261 -------------------------------------------------------------------------
    {
        auto scan_copy = std::make_shared<pcl::PointCloud<PointT>>(*cropped_scan_cloud);
        auto map_copy = std::make_shared<pcl::PointCloud<PointT>>(*ref_cropped_map_cloud_);
        applyUniformSubsample(map_copy, 15);   // :210
        removeFloor(scan_copy);                // :211-213
        removeFloor(map_copy);
        out.push_back((float)scan_copy->size());
        out.push_back((float)map_copy->size());
        brute_force_alignment_->setSourceCloud(scan_copy);       // :216-219
        brute_force_alignment_->setTargetCloud(map_copy);
        brute_force_alignment_->setInitialGuess(map_T_sensor_);
        const bool ok = brute_force_alignment_->alignClouds();
        out.push_back(ok ? 1.f : 0.f);
        out.push_back(brute_force_alignment_->firstAlignmentCompleted() ? 1.f : 0.f);
        put(out, brute_force_alignment_->getBestTransformation());
    }
    // ---- fine alignment, :335-338 ------------------------------------------------------------------------------------
    icp_->setSourcePointCloud(cropped_scan_cloud);
    icp_->setInitialTransformation(map_T_sensor_);
    const ICPResult host_path = icp_->calculateAlignment();
    put(out, host_path);
    // the same re-crop as a window change of the resident map (replaces :302-303): identical result
    std::size_t n_target = 0;
    resident.cropToTarget(map_T_sensor_, 10.0, icp_->handle(), &n_target);
    out.push_back((float)n_target);
    put(out, icp_->calculateAlignment());
    pcl::PointCloud<PointT>::Ptr again = std::make_shared<pcl::PointCloud<PointT>>();
    resident.crop(map_T_sensor_, 10.0, again);
    out.push_back((float)again->size());
    out.push_back(checksum(*again));
    FILE *f = std::fopen(argv[4], "wb");
    std::fwrite(out.data(), sizeof(float), out.size(), f);
    std::fclose(f);
    return 0;
}
