// shim_driver.cpp -- drives the C++ drop-in class the way localization_node.cpp does
// (reference localization/src/localization_node.cpp:24-29, 223-236, 303, 335-338) and writes
// the results to a file that tests/test_cpp_shim.py compares with the oracle.
//   shim_driver <map.f32> <scan.f32> <T0.f32 (16, column-major)> <out.f32>
#define SSF_SHIM_STANDALONE
#include <localization/icp_point_to_point.h>

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

static void put(std::vector<float> &out, const ICPResult &r)
{
    for (int i = 0; i < 16; ++i) out.push_back(r.transformation.data()[i]);
    out.push_back(r.error);
    out.push_back((float)r.iterations);
    out.push_back(r.has_converged ? 1.f : 0.f);
}

int main(int argc, char **argv)
{
    if (argc < 5) { std::fprintf(stderr, "usage: %s map scan T0 out\n", argv[0]); return argc == 2 && std::string(argv[1]) == "--help" ? 0 : 2; }
    auto map = load_cloud(argv[1]);
    auto scan = load_cloud(argv[2]);
    Eigen::Matrix4f T0 = Eigen::Matrix4f::Identity();
    {
        FILE *f = std::fopen(argv[3], "rb");
        if (!f || std::fread(T0.data(), sizeof(float), 16, f) != 16) { std::perror(argv[3]); return 2; }
        std::fclose(f);
    }
    // localization_node.cpp:24-29
    auto icp = std::make_shared<ICPPointToPoint>(0.5f, 10, 0.05f, 1e-5f);
    icp->setDebugMode(false);
    icp->setTargetPointCloud(map);  // :303
    std::vector<float> out;
    // fine alignment, :335-337
    icp->setSourcePointCloud(scan);
    icp->setInitialTransformation(T0);
    put(out, icp->calculateAlignment());
    // the "strong" re-parameterisation of performCoarseAlignment, :223-236, then restore
    icp->setMaxCorrespondenceDist(5.0f);
    icp->setTransformationEpsilon(1e-2f);
    icp->setAcceptableMeanError(0.4f);
    icp->setNumIterations(80);
    icp->setSourcePointCloud(scan);
    icp->setInitialTransformation(T0);
    put(out, icp->calculateAlignment());
    icp->setMaxCorrespondenceDist(0.5f);
    icp->setNumIterations(10);
    icp->setTransformationEpsilon(1e-5f);
    icp->setAcceptableMeanError(0.05f);
    put(out, icp->calculateAlignment());  // same source and initial transform, fine parameters again
    // too few correspondences: sentinel result, caller's has_converged check (:231) sees false
    Eigen::Matrix4f far = Eigen::Matrix4f::Identity();
    far(0, 3) = 1e4f;
    icp->setInitialTransformation(far);
    put(out, icp->calculateAlignment());
    FILE *f = std::fopen(argv[4], "wb");
    std::fwrite(out.data(), sizeof(float), out.size(), f);
    std::fclose(f);
    return 0;
}
