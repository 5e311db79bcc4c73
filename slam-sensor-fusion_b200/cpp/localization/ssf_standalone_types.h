// ssf_standalone_types.h -- minimal stand-ins for pcl::PointXYZ / pcl::PointCloud / Eigen::Matrix4f with the
// same memory layout (16-byte xyz_ points, column-major 4x4 float), used by the drop-in headers of this
// directory when SSF_SHIM_STANDALONE is defined (this repository's tests: PCL, Eigen and ROS 2 are not
// installed in the build image).  In a ROS 2 workspace the real types are used instead.
#ifndef SSF_STANDALONE_TYPES_H
#define SSF_STANDALONE_TYPES_H
#ifdef SSF_SHIM_STANDALONE
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

namespace pcl {
struct alignas(16) PointXYZ {
    float x, y, z, data_pad;
    PointXYZ() : x(0), y(0), z(0), data_pad(1.f) {}
    PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_), data_pad(1.f) {}
};
template <class P>
struct PointCloud {
    using Ptr = std::shared_ptr<PointCloud<P>>;
    std::vector<P> points;
    std::uint32_t width = 0, height = 0;
    bool is_dense = true;
    std::size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = height = 0; }
    P &operator[](std::size_t i) { return points[i]; }
    const P &operator[](std::size_t i) const { return points[i]; }
};
}  // namespace pcl
namespace Eigen {
struct Matrix4f {  // column-major like Eigen
    float m[16];
    static Matrix4f Identity()
    {
        Matrix4f r;
        for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.f : 0.f;
        return r;
    }
    float &operator()(int r, int c) { return m[c * 4 + r]; }
    float operator()(int r, int c) const { return m[c * 4 + r]; }
    const float *data() const { return m; }
    float *data() { return m; }
    float trace() const { return (m[0] + m[5]) + (m[10] + m[15]); }
};
}  // namespace Eigen
#endif  // SSF_SHIM_STANDALONE
#endif  // SSF_STANDALONE_TYPES_H
