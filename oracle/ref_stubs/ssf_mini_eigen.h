// ssf_mini_eigen.h -- minimal stand-in for the parts of Eigen 3 that the reference's hot-path
// sources use, so that /root/reference/localization/src/{icp_point_to_point,brute_force_alignment}.cpp
// and include/localization/point_cloud_processing.hpp compile UNMODIFIED into oracle/_ref.
//
// TEST INFRASTRUCTURE ONLY (same rule as ssf_oracle.c): never included by the product.
//
// This is not Eigen.  Every operation keeps the float evaluation order Eigen 3.3/3.4 uses for
// the same expression on x86-64 with SSE2 and no FMA (the reference's build, CMakeLists.txt:5-11):
//   * fixed 4x4 * 4x4 and 4x4 * 4-vector: packet path, result column = ((c0*b0 + c1*b1) + c2*b2) + c3*b3
//   * fixed 3x3 * 3x3, 3x3 * 3-vector, squaredNorm of a 3-vector: coefficient path with the
//     unrolled reduction  a0 + (a1 + a2)            (redux_novec_unroller, Length 3 -> 1 + 2)
//   * determinant of a 3x3: bruteforce_det3_helper order
//   * X^T * Y with dynamic depth (the 3x3 cross-covariance, icp_point_to_point.cpp:134): Eigen
//     runs its blocked GEMM whose depth blocking depends on the host's cache sizes, so the
//     summation order is not a property of the source; this stand-in adds row by row
//     (sequential), the same choice as ssf_oracle.c.  [ext]
//   * JacobiSVD<Matrix3f>: the published two-sided Jacobi algorithm (real 2x2 kernel), written
//     here from its description, independent of the C restatement in ssf_oracle.c.  [ext]
//   * AngleAxisf::toRotationMatrix(): the published formula, op for op.
// What this buys: control flow, thresholds, shrink/re-search rules, composition order and every
// float expression that is spelled out in the reference's own text come from the reference's
// text, not from a restatement.
#ifndef SSF_MINI_EIGEN_H
#define SSF_MINI_EIGEN_H
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>

namespace Eigen {

typedef std::ptrdiff_t Index;
enum { ComputeFullU = 0x04, ComputeThinU = 0x08, ComputeFullV = 0x10, ComputeThinV = 0x20 };

namespace mini {
// Eigen's default IOFormat: columns aligned to the widest coefficient, " " between
// coefficients, "\n" between rows, stream precision.
template <class Get>
inline std::ostream &print(std::ostream &s, int rows, int cols, Get get)
{
    std::streamsize width = 0;
    for (int j = 0; j < cols; ++j)
        for (int i = 0; i < rows; ++i) {
            std::stringstream ss;
            ss.copyfmt(s);
            ss << get(i, j);
            width = std::max<std::streamsize>(width, (std::streamsize)ss.str().length());
        }
    for (int i = 0; i < rows; ++i) {
        if (i) s << "\n";
        if (width) s.width(width);
        s << get(i, 0);
        for (int j = 1; j < cols; ++j) {
            s << " ";
            if (width) s.width(width);
            s << get(i, j);
        }
    }
    return s;
}
inline float red3(float a0, float a1, float a2) { const float t = a1 + a2; return a0 + t; }
} // namespace mini

struct Vector3f;
struct Matrix3f;

struct RowVector3f {
    float v[3];
    RowVector3f() : v{0.f, 0.f, 0.f} {}
    RowVector3f(float x, float y, float z) : v{x, y, z} {}
    float operator()(int i) const { return v[i]; }
    float squaredNorm() const { return mini::red3(v[0] * v[0], v[1] * v[1], v[2] * v[2]); }
    float norm() const { return std::sqrt(squaredNorm()); }
};

struct Vector3f {
    float v[3];
    Vector3f() : v{0.f, 0.f, 0.f} {}
    Vector3f(float x, float y, float z) : v{x, y, z} {}
    Vector3f(const RowVector3f &r) : v{r.v[0], r.v[1], r.v[2]} {} // Eigen transposes vectors on assignment
    static Vector3f UnitZ() { return Vector3f(0.f, 0.f, 1.f); }
    static Vector3f Zero() { return Vector3f(); }
    float &operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
    float &operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    float x() const { return v[0]; }
    float y() const { return v[1]; }
    float z() const { return v[2]; }
    RowVector3f transpose() const { return RowVector3f(v[0], v[1], v[2]); }
    Vector3f &operator+=(const RowVector3f &r) { v[0] += r.v[0]; v[1] += r.v[1]; v[2] += r.v[2]; return *this; }
    Vector3f &operator+=(const Vector3f &r) { v[0] += r.v[0]; v[1] += r.v[1]; v[2] += r.v[2]; return *this; }
    // DenseBase::operator/=(const Scalar&): the argument is converted to the scalar type, then a
    // coefficient-wise true division (not a multiplication by the reciprocal)
    Vector3f &operator/=(float s) { v[0] /= s; v[1] /= s; v[2] /= s; return *this; }
    Vector3f &operator*=(float s) { v[0] *= s; v[1] *= s; v[2] *= s; return *this; }
    float squaredNorm() const { return mini::red3(v[0] * v[0], v[1] * v[1], v[2] * v[2]); }
    float norm() const { return std::sqrt(squaredNorm()); }
};
inline Vector3f operator-(const Vector3f &a, const Vector3f &b) { return Vector3f(a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]); }
inline Vector3f operator+(const Vector3f &a, const Vector3f &b) { return Vector3f(a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]); }
inline Vector3f operator*(float s, const Vector3f &a) { return Vector3f(s * a.v[0], s * a.v[1], s * a.v[2]); }
inline RowVector3f operator-(const RowVector3f &a, const RowVector3f &b) { return RowVector3f(a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]); }

struct Vector4f {
    float v[4];
    Vector4f() : v{0.f, 0.f, 0.f, 0.f} {}
    Vector4f(float x, float y, float z, float w) : v{x, y, z, w} {}
    float &operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    float &operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
};

// column-major like Eigen
struct Matrix3f {
    float m[9];
    Matrix3f() { for (float &x : m) x = 0.f; }
    static Matrix3f Identity() { Matrix3f r; r.m[0] = r.m[4] = r.m[8] = 1.f; return r; }
    static Matrix3f Zero() { return Matrix3f(); }
    float &operator()(int r, int c) { return m[c * 3 + r]; }
    float operator()(int r, int c) const { return m[c * 3 + r]; }
    Matrix3f transpose() const { Matrix3f t; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t(c, r) = (*this)(r, c); return t; }
    float determinant() const
    {
        const Matrix3f &a = *this; // determinant_impl<Derived, 3>
        return a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) - a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) +
               a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
    }
    struct ColRef {
        float *p;
        ColRef &operator*=(float s) { p[0] *= s; p[1] *= s; p[2] *= s; return *this; }
    };
    ColRef col(int c) { return ColRef{m + 3 * c}; }
};
inline Matrix3f operator*(const Matrix3f &a, const Matrix3f &b)
{
    Matrix3f r;
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 3; ++i) r(i, c) = mini::red3(a(i, 0) * b(0, c), a(i, 1) * b(1, c), a(i, 2) * b(2, c));
    return r;
}
inline Vector3f operator*(const Matrix3f &a, const Vector3f &b)
{
    Vector3f r;
    for (int i = 0; i < 3; ++i) r.v[i] = mini::red3(a(i, 0) * b.v[0], a(i, 1) * b.v[1], a(i, 2) * b.v[2]);
    return r;
}
inline std::ostream &operator<<(std::ostream &s, const Matrix3f &a) { return mini::print(s, 3, 3, [&](int i, int j) { return a(i, j); }); }
inline std::ostream &operator<<(std::ostream &s, const Vector3f &a) { return mini::print(s, 3, 1, [&](int i, int) { return a.v[i]; }); }

struct Matrix4f {
    float m[16];
    Matrix4f() { for (float &x : m) x = 0.f; }
    static Matrix4f Identity() { Matrix4f r; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.f; return r; }
    static Matrix4f Zero() { return Matrix4f(); }
    float &operator()(int r, int c) { return m[c * 4 + r]; }
    float operator()(int r, int c) const { return m[c * 4 + r]; }
    const float *data() const { return m; }
    float *data() { return m; }
    float trace() const { return (m[0] + m[5]) + (m[10] + m[15]); } // redux unroller, Length 4 -> 2 + 2
    template <int R, int C> struct BlockRef {
        Matrix4f *o; int r0, c0;
        BlockRef &operator=(const Matrix3f &b) { static_assert(R == 3 && C == 3, "3x3 block"); for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) (*o)(r0 + r, c0 + c) = b(r, c); return *this; }
        BlockRef &operator=(const Vector3f &b) { static_assert(R == 3 && C == 1, "3x1 block"); for (int r = 0; r < 3; ++r) (*o)(r0 + r, c0) = b.v[r]; return *this; }
    };
    template <int R, int C> BlockRef<R, C> block(int r0, int c0) { return BlockRef<R, C>{this, r0, c0}; }
};
inline Matrix4f operator*(const Matrix4f &a, const Matrix4f &b)
{
    Matrix4f r;
    for (int c = 0; c < 4; ++c)
        for (int i = 0; i < 4; ++i) {
            float s = a(i, 0) * b(0, c);
            s += a(i, 1) * b(1, c);
            s += a(i, 2) * b(2, c);
            s += a(i, 3) * b(3, c);
            r(i, c) = s;
        }
    return r;
}
inline Vector4f operator*(const Matrix4f &a, const Vector4f &b)
{
    Vector4f r;
    for (int i = 0; i < 4; ++i) {
        float s = a(i, 0) * b.v[0];
        s += a(i, 1) * b.v[1];
        s += a(i, 2) * b.v[2];
        s += a(i, 3) * b.v[3];
        r.v[i] = s;
    }
    return r;
}
inline std::ostream &operator<<(std::ostream &s, const Matrix4f &a) { return mini::print(s, 4, 4, [&](int i, int j) { return a(i, j); }); }

// N x 3, column-major (three planes) like Eigen::Matrix<float, Dynamic, 3>
struct MatrixX3f {
    std::vector<float> d;
    Index n = 0;
    MatrixX3f() {}
    MatrixX3f(Index rows, Index cols) { resize(rows, cols); }
    void resize(Index rows, Index /*cols == 3*/) { n = rows; d.assign((size_t)(3 * rows), 0.f); } // Eigen leaves it uninitialised
    Index rows() const { return n; }
    Index cols() const { return 3; }
    float &operator()(Index r, Index c) { return d[(size_t)(c * n + r)]; }
    float operator()(Index r, Index c) const { return d[(size_t)(c * n + r)]; }
    struct ConstRow {
        const MatrixX3f *o; Index r;
        operator RowVector3f() const { return RowVector3f((*o)(r, 0), (*o)(r, 1), (*o)(r, 2)); }
        operator Vector3f() const { return Vector3f((*o)(r, 0), (*o)(r, 1), (*o)(r, 2)); }
    };
    struct Row {
        MatrixX3f *o; Index r;
        operator RowVector3f() const { return RowVector3f((*o)(r, 0), (*o)(r, 1), (*o)(r, 2)); }
        operator Vector3f() const { return Vector3f((*o)(r, 0), (*o)(r, 1), (*o)(r, 2)); }
        Row &operator=(const RowVector3f &v) { (*o)(r, 0) = v.v[0]; (*o)(r, 1) = v.v[1]; (*o)(r, 2) = v.v[2]; return *this; }
        Row &operator=(const Vector3f &v) { (*o)(r, 0) = v.v[0]; (*o)(r, 1) = v.v[1]; (*o)(r, 2) = v.v[2]; return *this; }
        Row &operator=(const Row &v) { return *this = (RowVector3f)v; }
        Row(const Row &) = default;
        Row(MatrixX3f *o_, Index r_) : o(o_), r(r_) {}
    };
    Row row(Index r) { return Row(this, r); }
    ConstRow row(Index r) const { return ConstRow{this, r}; }
    struct Transposed { const MatrixX3f *o; };
    Transposed transpose() const { return Transposed{this}; }
};
inline RowVector3f operator-(const MatrixX3f::ConstRow &a, const RowVector3f &b) { return (RowVector3f)a - b; }
inline RowVector3f operator-(const MatrixX3f::ConstRow &a, const MatrixX3f::ConstRow &b) { return (RowVector3f)a - (RowVector3f)b; }
inline RowVector3f operator-(const MatrixX3f::Row &a, const RowVector3f &b) { return (RowVector3f)a - b; }
inline Vector3f &operator+=(Vector3f &a, const MatrixX3f::ConstRow &b) { return a += (RowVector3f)b; }
// (3 x K) * (K x 3): see the header comment -- sequential over the rows.  [ext]
inline Matrix3f operator*(const MatrixX3f::Transposed &a, const MatrixX3f &b)
{
    Matrix3f H;
    const MatrixX3f &A = *a.o;
    for (Index i = 0; i < A.rows(); ++i)
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) H(r, c) += A(i, r) * b(i, c);
    return H;
}

struct AngleAxisf {
    float angle_;
    Vector3f axis_;
    AngleAxisf(float angle, const Vector3f &axis) : angle_(angle), axis_(axis) {}
    Matrix3f toRotationMatrix() const
    {
        Matrix3f res;
        const Vector3f sin_axis = std::sin(angle_) * axis_;
        const float c = std::cos(angle_);
        const Vector3f cos1_axis = (1.f - c) * axis_;
        float tmp;
        tmp = cos1_axis.x() * axis_.y();
        res(0, 1) = tmp - sin_axis.z();
        res(1, 0) = tmp + sin_axis.z();
        tmp = cos1_axis.x() * axis_.z();
        res(0, 2) = tmp + sin_axis.y();
        res(2, 0) = tmp - sin_axis.y();
        tmp = cos1_axis.y() * axis_.z();
        res(1, 2) = tmp - sin_axis.x();
        res(2, 1) = tmp + sin_axis.x();
        res(0, 0) = cos1_axis.x() * axis_.x() + c;
        res(1, 1) = cos1_axis.y() * axis_.y() + c;
        res(2, 2) = cos1_axis.z() * axis_.z() + c;
        return res;
    }
};

// Two-sided Jacobi SVD of a real square matrix (the algorithm Eigen::JacobiSVD documents): sweep
// over the 2x2 sub-problems (p, q), q < p, until every off-diagonal pair is below
// precision * max|diagonal|; each sub-problem is first symmetrised by a left rotation, then
// diagonalised by a Jacobi rotation; finally signs are moved into U and the singular values are
// sorted in decreasing order.
template <class M> class JacobiSVD;
template <> class JacobiSVD<Matrix3f>
{
public:
    JacobiSVD(const Matrix3f &A, unsigned /*options*/) { compute(A); }
    const Matrix3f &matrixU() const { return U_; }
    const Matrix3f &matrixV() const { return V_; }
    const Vector3f &singularValues() const { return S_; }

private:
    struct Rot { float c, s; }; // JacobiRotation: [[c, s], [-s, c]] when applied on the left
    static void applyLeft(Matrix3f &W, int p, int q, Rot j)
    { // rows p, q <- J^* [row p; row q]  with  x' = c x + s y,  y' = -s x + c y
        for (int k = 0; k < 3; ++k) { const float x = W(p, k), y = W(q, k); W(p, k) = j.c * x + j.s * y; W(q, k) = -j.s * x + j.c * y; }
    }
    static void applyRight(Matrix3f &W, int p, int q, Rot j)
    { // cols p, q <- [col p, col q] J  with  x' = c x - s y,  y' = s x + c y
        for (int k = 0; k < 3; ++k) { const float x = W(k, p), y = W(k, q); W(k, p) = j.c * x - j.s * y; W(k, q) = j.s * x + j.c * y; }
    }
    static Rot makeJacobi(float x, float y, float z)
    { // diagonalises [[x, y], [y, z]]
        const float deno = 2.f * std::fabs(y);
        if (deno < FLT_MIN) return Rot{1.f, 0.f};
        const float tau = (x - z) / deno, w = std::sqrt(tau * tau + 1.f);
        const float t = tau > 0.f ? 1.f / (tau + w) : 1.f / (tau - w);
        const float sign_t = t > 0.f ? 1.f : -1.f, n = 1.f / std::sqrt(t * t + 1.f);
        return Rot{n, -sign_t * (y / std::fabs(y)) * std::fabs(t) * n};
    }
    void compute(const Matrix3f &A)
    {
        const float precision = 2.f * FLT_EPSILON, tiny = FLT_MIN;
        float scale = 0.f;
        for (float x : A.m) scale = std::max(scale, std::fabs(x));
        if (!(scale > 0.f) || !std::isfinite(scale)) scale = 1.f;
        Matrix3f W;
        for (int i = 0; i < 9; ++i) W.m[i] = A.m[i] / scale;
        U_ = Matrix3f::Identity();
        V_ = Matrix3f::Identity();
        float maxDiag = std::max(std::fabs(W(0, 0)), std::max(std::fabs(W(1, 1)), std::fabs(W(2, 2))));
        bool finished = false;
        for (int sweep = 0; !finished && sweep < 64; ++sweep) {
            finished = true;
            for (int p = 1; p < 3; ++p)
                for (int q = 0; q < p; ++q) {
                    const float threshold = std::max(tiny, precision * maxDiag);
                    if (!(std::fabs(W(p, q)) > threshold || std::fabs(W(q, p)) > threshold)) continue;
                    finished = false;
                    // real_2x2_jacobi_svd
                    const float m00 = W(p, p), m01 = W(p, q), m10 = W(q, p), m11 = W(q, q);
                    const float t = m00 + m11, d = m10 - m01;
                    Rot rot1;
                    if (std::fabs(d) < tiny) rot1 = Rot{1.f, 0.f};
                    else { const float u = t / d, tmp = std::sqrt(1.f + u * u); rot1 = Rot{u / tmp, 1.f / tmp}; }
                    const float a00 = rot1.c * m00 + rot1.s * m10, a01 = rot1.c * m01 + rot1.s * m11, a11 = -rot1.s * m01 + rot1.c * m11;
                    const Rot jr = makeJacobi(a00, a01, a11);
                    // j_left = rot1 * j_right^T
                    const Rot jrt{jr.c, -jr.s};
                    const Rot jl{rot1.c * jrt.c - rot1.s * jrt.s, rot1.c * jrt.s + rot1.s * jrt.c};
                    applyLeft(W, p, q, jl);
                    applyRight(U_, p, q, Rot{jl.c, -jl.s});
                    applyRight(W, p, q, jr);
                    applyRight(V_, p, q, jr);
                    maxDiag = std::max(maxDiag, std::max(std::fabs(W(p, p)), std::fabs(W(q, q))));
                }
        }
        for (int i = 0; i < 3; ++i) {
            const float a = std::fabs(W(i, i));
            S_.v[i] = a;
            if (a != 0.f) { const float f = W(i, i) / a; for (int k = 0; k < 3; ++k) U_(k, i) *= f; }
        }
        for (int i = 0; i < 3; ++i) S_.v[i] *= scale;
        for (int i = 0; i < 3; ++i) {
            int pos = i;
            for (int k = i + 1; k < 3; ++k) if (S_.v[k] > S_.v[pos]) pos = k;
            if (S_.v[pos] == 0.f) break;
            if (pos != i) {
                std::swap(S_.v[i], S_.v[pos]);
                for (int k = 0; k < 3; ++k) { std::swap(U_(k, i), U_(k, pos)); std::swap(V_(k, i), V_(k, pos)); }
            }
        }
    }
    Matrix3f U_, V_;
    Vector3f S_;
};

} // namespace Eigen
#endif
