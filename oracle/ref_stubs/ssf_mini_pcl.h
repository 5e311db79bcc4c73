// ssf_mini_pcl.h -- minimal stand-in for the parts of PCL the reference's hot-path sources use
// (PointXYZ, PointCloud, KdTreeFLANN::nearestKSearch, search::KdTree::radiusSearch,
// PointIndices, ExtractIndices), so that the reference's own .cpp/.hpp files compile UNMODIFIED
// into oracle/_ref.  TEST INFRASTRUCTURE ONLY; never included by the product.
//
// This is not PCL.  Semantics restated from PCL/FLANN's published behaviour  [ext]:
//   * KdTreeFLANN::setInputCloud indexes the finite points' (x, y, z); nearestKSearch(k = 1) is
//     an exact search returning the SQUARED L2 distance accumulated left to right in float
//     (flann::L2_Simple).  Equal-distance ties: lowest point index (the project's contract;
//     FLANN itself keeps the first candidate its traversal reaches).
//   * search::KdTree::radiusSearch keeps dist2 < float(radius * radius) (strict), sorted by
//     ascending distance, equal distances by index (FLANN's DistanceIndex ordering).
//   * ExtractIndices::filter copies the indexed points in index-list order and may write into
//     its own input cloud.
// The k = 1 search below is an exact KD-tree written for this file (median split on the widest
// axis, pruning on a lower bound kept strictly conservative), independent of ssf_oracle.c's.
#ifndef SSF_MINI_PCL_H
#define SSF_MINI_PCL_H
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <memory>
#include <numeric>
#include <vector>

namespace pcl {

struct alignas(16) PointXYZ {
    float x, y, z, pad;
    PointXYZ() : x(0.f), y(0.f), z(0.f), pad(1.f) {}
    PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_), pad(1.f) {}
};

template <class PointT> class PointCloud
{
public:
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
    std::vector<PointT> points;
    std::uint32_t width = 0, height = 0;
    bool is_dense = true;
    std::size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = height = 0; }
    void push_back(const PointT &p) { points.push_back(p); width = (std::uint32_t)points.size(); height = 1; }
    PointT &operator[](std::size_t i) { return points[i]; }
    const PointT &operator[](std::size_t i) const { return points[i]; }
    PointCloud &operator+=(const PointCloud &o)
    {
        points.insert(points.end(), o.points.begin(), o.points.end());
        width = (std::uint32_t)points.size();
        height = 1;
        return *this;
    }
};

struct PointIndices {
    typedef std::shared_ptr<PointIndices> Ptr;
    std::vector<int> indices;
};

namespace mini {
inline float sqdist(const float *a, const float *b)
{
    const float d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    float r = d0 * d0;
    r += d1 * d1;
    r += d2 * d2;
    return r;
}

class ExactKd
{
public:
    void build(const std::vector<PointXYZ> &pts)
    {
        xyz_.clear(); ids_.clear(); nodes_.clear();
        for (std::size_t i = 0; i < pts.size(); ++i)
            if (std::isfinite(pts[i].x) && std::isfinite(pts[i].y) && std::isfinite(pts[i].z)) ids_.push_back((int)i);
        src_ = &pts;
        if (!ids_.empty()) split(0, (int)ids_.size());
        xyz_.resize(3 * ids_.size());
        for (std::size_t k = 0; k < ids_.size(); ++k) { const PointXYZ &p = pts[(std::size_t)ids_[k]]; xyz_[3 * k] = p.x; xyz_[3 * k + 1] = p.y; xyz_[3 * k + 2] = p.z; }
        src_ = nullptr;
    }
    bool empty() const { return ids_.empty(); }
    void nearest(const float q[3], int &best_id, float &best_d2) const
    {
        best_id = -1; best_d2 = FLT_MAX;
        if (ids_.empty()) return;
        double off[3] = {0.0, 0.0, 0.0};
        descend(0, q, off, 0.0, best_id, best_d2);
    }

private:
    struct Node { int lo, hi, left, right, axis; float lmax, rmin; };
    int split(int lo, int hi)
    {
        const int id = (int)nodes_.size();
        nodes_.push_back(Node{lo, hi, -1, -1, 0, 0.f, 0.f});
        if (hi - lo <= 12) return id;
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int k = lo; k < hi; ++k) {
            const PointXYZ &p = (*src_)[(std::size_t)ids_[(std::size_t)k]];
            const float c[3] = {p.x, p.y, p.z};
            for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], c[a]); mx[a] = std::max(mx[a], c[a]); }
        }
        int axis = 0;
        for (int a = 1; a < 3; ++a) if (mx[a] - mn[a] > mx[axis] - mn[axis]) axis = a;
        const int mid = lo + (hi - lo) / 2;
        auto coord = [&](int i) { const PointXYZ &p = (*src_)[(std::size_t)i]; return axis == 0 ? p.x : (axis == 1 ? p.y : p.z); };
        std::nth_element(ids_.begin() + lo, ids_.begin() + mid, ids_.begin() + hi,
                         [&](int a, int b) { const float ca = coord(a), cb = coord(b); return ca < cb || (ca == cb && a < b); });
        float lmax = -FLT_MAX, rmin = FLT_MAX;
        for (int k = lo; k < mid; ++k) lmax = std::max(lmax, coord(ids_[(std::size_t)k]));
        for (int k = mid; k < hi; ++k) rmin = std::min(rmin, coord(ids_[(std::size_t)k]));
        const int l = split(lo, mid), r = split(mid, hi);
        Node &n = nodes_[(std::size_t)id];
        n.left = l; n.right = r; n.axis = axis; n.lmax = lmax; n.rmin = rmin;
        return id;
    }
    // off[a] = squared gap already known along axis a; bound = their sum (a lower bound, in double,
    // of the real squared distance to anything in the subtree)
    void descend(int id, const float q[3], double off[3], double bound, int &best_id, float &best_d2) const
    {
        const Node &n = nodes_[(std::size_t)id];
        if (n.left < 0) {
            for (int k = n.lo; k < n.hi; ++k) {
                const float d = sqdist(q, &xyz_[3 * (std::size_t)k]);
                const int pid = ids_[(std::size_t)k];
                if (d < best_d2 || (d == best_d2 && pid < best_id)) { best_d2 = d; best_id = pid; }
            }
            return;
        }
        const double v = q[n.axis];
        const double gl = v - (double)n.lmax, gr = (double)n.rmin - v; // > 0: outside that child along the axis
        const bool left_first = gl + (-gr) < 0.0;
        const int first = left_first ? n.left : n.right, second = left_first ? n.right : n.left;
        const double gap = left_first ? gr : gl;
        descend(first, q, off, bound, best_id, best_d2);
        const double cut = gap > 0.0 ? gap * gap : 0.0;
        const double saved = off[n.axis];
        const double far_bound = bound - saved + std::max(saved, cut);
        // a float d2 can undershoot the real squared distance by a few ulp: keep a relative margin
        // so that equal-distance candidates are always visited (tie -> lowest index)
        if (far_bound * (1.0 - 1e-6) <= (double)best_d2) {
            off[n.axis] = std::max(saved, cut);
            descend(second, q, off, far_bound, best_id, best_d2);
            off[n.axis] = saved;
        }
    }
    const std::vector<PointXYZ> *src_ = nullptr;
    std::vector<int> ids_;
    std::vector<float> xyz_;
    std::vector<Node> nodes_;
};
} // namespace mini

template <class PointT> class KdTreeFLANN
{
public:
    typedef std::shared_ptr<KdTreeFLANN<PointT>> Ptr;
    void setInputCloud(const typename PointCloud<PointT>::Ptr &cloud) { cloud_ = cloud; tree_.build(cloud->points); }
    void setInputCloud(const typename PointCloud<PointT>::ConstPtr &cloud) { cloud_ = cloud; tree_.build(cloud->points); }
    int nearestKSearch(const PointT &p, int k, std::vector<int> &k_indices, std::vector<float> &k_sqr_distances) const
    {
        if (k != 1 || tree_.empty()) { k_indices.clear(); k_sqr_distances.clear(); return 0; }
        const float q[3] = {p.x, p.y, p.z};
        int id; float d2;
        tree_.nearest(q, id, d2);
        k_indices.resize(1); k_sqr_distances.resize(1);
        k_indices[0] = id; k_sqr_distances[0] = d2;
        return 1;
    }
    int radiusSearch(const PointT &p, double radius, std::vector<int> &k_indices, std::vector<float> &k_sqr_distances,
                     unsigned int max_nn = 0) const
    {
        k_indices.clear(); k_sqr_distances.clear();
        if (!cloud_) return 0;
        const float r2 = static_cast<float>(radius * radius);
        const float q[3] = {p.x, p.y, p.z};
        std::vector<std::pair<float, int>> hits;
        for (std::size_t i = 0; i < cloud_->points.size(); ++i) {
            const PointT &c = cloud_->points[i];
            if (!(std::isfinite(c.x) && std::isfinite(c.y) && std::isfinite(c.z))) continue;
            const float cc[3] = {c.x, c.y, c.z};
            const float d = mini::sqdist(q, cc);
            if (d < r2) hits.emplace_back(d, (int)i);
        }
        std::sort(hits.begin(), hits.end());
        if (max_nn > 0 && hits.size() > max_nn) hits.resize(max_nn);
        for (const auto &h : hits) { k_sqr_distances.push_back(h.first); k_indices.push_back(h.second); }
        return (int)hits.size();
    }

private:
    typename PointCloud<PointT>::ConstPtr cloud_;
    mini::ExactKd tree_;
};

namespace search {
template <class PointT> class KdTree : public pcl::KdTreeFLANN<PointT>
{
public:
    typedef std::shared_ptr<KdTree<PointT>> Ptr;
};
} // namespace search

template <class PointT> class ExtractIndices
{
public:
    void setInputCloud(const typename PointCloud<PointT>::Ptr &cloud) { input_ = cloud; }
    void setIndices(const PointIndices::Ptr &indices) { indices_ = indices; }
    void setNegative(bool negative) { negative_ = negative; }
    void filter(PointCloud<PointT> &output)
    {
        std::vector<PointT> out;
        if (!negative_) {
            out.reserve(indices_->indices.size());
            for (int i : indices_->indices) out.push_back(input_->points[(std::size_t)i]);
        } else {
            std::vector<char> drop(input_->points.size(), 0);
            for (int i : indices_->indices) drop[(std::size_t)i] = 1;
            for (std::size_t i = 0; i < input_->points.size(); ++i) if (!drop[i]) out.push_back(input_->points[i]);
        }
        output.points.swap(out);
        output.width = (std::uint32_t)output.points.size();
        output.height = 1;
        output.is_dense = true;
    }

private:
    typename PointCloud<PointT>::Ptr input_;
    PointIndices::Ptr indices_;
    bool negative_ = false;
};

} // namespace pcl
#endif
