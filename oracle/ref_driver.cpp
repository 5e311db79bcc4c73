// ref_driver.cpp -- extern "C" entry points around the reference's OWN sources, compiled
// unmodified from /root/reference into oracle/_ref/libssf_ref.so (recipe: oracle/Makefile,
// target `ref`).  TEST INFRASTRUCTURE ONLY: used by tests/ to pin ssf_oracle.c (and through it
// the GPU path) to the reference's text, and by bench.py as the "reference" CPU baseline.
//
// Compiled together with, and calling only the public API of:
//   localization/src/icp_point_to_point.cpp        (ICPPointToPoint, icp_point_to_point.h:41-85)
//   localization/src/brute_force_alignment.cpp     (BruteForceAlignment, brute_force_alignment.h:22-112)
//   localization/include/localization/point_cloud_processing.hpp:31-92 (three free functions)
// against the stand-in Eigen / PCL headers in oracle/ref_stubs/ (PCL, FLANN and Eigen are not
// installed in this image; see the header comments there for what is restated and how).
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>

#include <localization/icp_point_to_point.h>
#include <localization/brute_force_alignment.h>
#include <localization/point_cloud_processing.hpp>

namespace {
pcl::PointCloud<PointT>::Ptr make_cloud(const float *xyz, int64_t n, int stride)
{
    pcl::PointCloud<PointT>::Ptr c(new pcl::PointCloud<PointT>());
    c->points.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) c->points[(size_t)i] = PointT(xyz[i * stride], xyz[i * stride + 1], xyz[i * stride + 2]);
    c->width = (std::uint32_t)n;
    c->height = 1;
    return c;
}
Eigen::Matrix4f make_T(const float *colmajor)
{
    Eigen::Matrix4f T;
    std::memcpy(T.data(), colmajor, sizeof(float) * 16);
    return T;
}
int64_t write_cloud(const pcl::PointCloud<PointT> &c, float *out)
{
    for (size_t i = 0; i < c.points.size(); ++i) { out[4 * i] = c.points[i].x; out[4 * i + 1] = c.points[i].y; out[4 * i + 2] = c.points[i].z; out[4 * i + 3] = 1.f; }
    return (int64_t)c.points.size();
}
// captures std::cout / std::cerr while alive
struct Capture {
    std::ostringstream out, err;
    std::streambuf *o, *e;
    Capture() : o(std::cout.rdbuf(out.rdbuf())), e(std::cerr.rdbuf(err.rdbuf())) {}
    ~Capture() { std::cout.rdbuf(o); std::cerr.rdbuf(e); }
};
std::string g_stdout, g_stderr;
} // namespace

struct ssf_ref_icp { ICPPointToPoint *icp; };

extern "C" {

struct ssf_ref_result {
    float transformation[16]; // column-major
    float error;
    int32_t iterations;
    int32_t has_converged;
};

// ICPPointToPoint(max_correspondence_dist, num_iterations, acceptable_mean_error, transformation_epsilon)
void *ssf_ref_icp_create(float max_corr, int num_iterations, float acceptable_mean_error, float transformation_epsilon)
{
    ssf_ref_icp *h = new ssf_ref_icp;
    h->icp = new ICPPointToPoint(max_corr, num_iterations, acceptable_mean_error, transformation_epsilon);
    h->icp->setDebugMode(false);
    return h;
}
void ssf_ref_icp_destroy(void *hv)
{
    ssf_ref_icp *h = (ssf_ref_icp *)hv;
    if (!h) return;
    delete h->icp;
    delete h;
}
void ssf_ref_icp_set_params(void *hv, float max_corr, int num_iterations, float acceptable_mean_error, float transformation_epsilon)
{
    ICPPointToPoint *icp = ((ssf_ref_icp *)hv)->icp;
    icp->setMaxCorrespondenceDist(max_corr);
    icp->setNumIterations(num_iterations);
    icp->setAcceptableMeanError(acceptable_mean_error);
    icp->setTransformationEpsilon(transformation_epsilon);
}
void ssf_ref_icp_set_debug(void *hv, int on) { ((ssf_ref_icp *)hv)->icp->setDebugMode(on != 0); }
void ssf_ref_icp_set_target(void *hv, const float *xyz, int64_t n, int stride)
{
    auto c = make_cloud(xyz, n, stride);
    ((ssf_ref_icp *)hv)->icp->setTargetPointCloud(c);
}
void ssf_ref_icp_set_source(void *hv, const float *xyz, int64_t n, int stride)
{
    auto c = make_cloud(xyz, n, stride);
    ((ssf_ref_icp *)hv)->icp->setSourcePointCloud(c);
}
void ssf_ref_icp_set_initial(void *hv, const float *T_colmajor) { ((ssf_ref_icp *)hv)->icp->setInitialTransformation(make_T(T_colmajor)); }

// calculateAlignment(); stdout / stderr of the call are captured (ssf_ref_last_stdout / _stderr)
void ssf_ref_icp_align(void *hv, ssf_ref_result *res)
{
    ICPResult r;
    {
        Capture cap;
        r = ((ssf_ref_icp *)hv)->icp->calculateAlignment();
        g_stdout = cap.out.str();
        g_stderr = cap.err.str();
    }
    std::memcpy(res->transformation, r.transformation.data(), sizeof(float) * 16);
    res->error = r.error;
    res->iterations = r.iterations;
    res->has_converged = r.has_converged ? 1 : 0;
}
const char *ssf_ref_last_stdout() { return g_stdout.c_str(); }
const char *ssf_ref_last_stderr() { return g_stderr.c_str(); }

// default-constructed ICPResult (icp_point_to_point.h:28-39)
void ssf_ref_default_result(ssf_ref_result *res)
{
    ICPResult r;
    std::memcpy(res->transformation, r.transformation.data(), sizeof(float) * 16);
    res->error = r.error;
    res->iterations = r.iterations;
    res->has_converged = r.has_converged ? 1 : 0;
}

struct ssf_ref_bfa_params {
    float x_step, y_step, z_step;
    float x_range, y_range, z_range;
    float yaw_step, yaw_range;
    float mean_error_threshold;
};

// One BruteForceAlignment object: setters as at localization_node.cpp:38-43, setSourceCloud /
// setTargetCloud / setInitialGuess, n_calls consecutive alignClouds() (the node retries every
// callback; best-so-far is carried over inside the object).  success_out / T_out: per call.
void ssf_ref_bfa_align(const float *tgt, int64_t n_tgt, int tgt_stride, const float *src, int64_t n_src, int src_stride,
                       const float *T_guess, const ssf_ref_bfa_params *p, int n_calls, int32_t *success_out, float *T_out)
{
    BruteForceAlignment bfa;
    bfa.setXYZStep(p->x_step, p->y_step, p->z_step);
    bfa.setXYZRange(p->x_range, p->y_range, p->z_range);
    bfa.setRotationStep(p->yaw_step);
    bfa.setRotationRange(p->yaw_range);
    bfa.setMeanErrorThreshold(p->mean_error_threshold);
    auto s = make_cloud(src, n_src, src_stride), t = make_cloud(tgt, n_tgt, tgt_stride);
    bfa.setSourceCloud(s);
    bfa.setTargetCloud(t);
    bfa.setInitialGuess(make_T(T_guess));
    for (int c = 0; c < n_calls; ++c) {
        const bool ok = bfa.alignClouds();
        success_out[c] = ok ? 1 : 0;
        const Eigen::Matrix4f T = bfa.getBestTransformation();
        std::memcpy(T_out + 16 * c, T.data(), sizeof(float) * 16);
    }
}

// point_cloud_processing.hpp; outputs are float4 rows, return = number of rows
int64_t ssf_ref_crop_radius(const float *xyz, int64_t n, int stride, const float *T_colmajor, double radius, float *out)
{
    auto c = make_cloud(xyz, n, stride);
    pcl::PointCloud<PointT>::Ptr cropped(new pcl::PointCloud<PointT>());
    cropPointCloudThroughRadius(make_T(T_colmajor), radius, c, cropped);
    return write_cloud(*cropped, out);
}
int64_t ssf_ref_subsample(const float *xyz, int64_t n, int stride, int64_t step, float *out)
{
    auto c = make_cloud(xyz, n, stride);
    applyUniformSubsample(c, (std::size_t)step);
    return write_cloud(*c, out);
}
int64_t ssf_ref_remove_floor(const float *xyz, int64_t n, int stride, float *out)
{
    auto c = make_cloud(xyz, n, stride);
    removeFloor(c);
    return write_cloud(*c, out);
}

} // extern "C"
