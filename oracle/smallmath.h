// TEST INFRASTRUCTURE — part of the CPU oracle (see oracle/README.md).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
// legs may build or call anything in this directory.
//
// Small fixed-size linear algebra templated on the scalar, plus a
// forward-mode dual number.  The reference tapes its Scalar-templated model
// code with CppAD (upright_core/include/upright_core/types.h:9-31 — every
// type is `Eigen::Matrix<Scalar,...>`); the oracle gets the same derivatives
// by instantiating its Scalar-templated restatement with `Dual`.
#pragma once
#include <cmath>
#include <cstring>

namespace orc {

constexpr int MAXDIR = 27;  // d/dx for nx <= 27

struct Dual {
    double v;
    double d[MAXDIR];
    Dual() : v(0) { std::memset(d, 0, sizeof(d)); }
    Dual(double c) : v(c) { std::memset(d, 0, sizeof(d)); }  // NOLINT implicit
    static Dual variable(double c, int dir) {
        Dual r(c);
        r.d[dir] = 1.0;
        return r;
    }
};
inline Dual operator+(const Dual& a, const Dual& b) {
    Dual r;
    r.v = a.v + b.v;
    for (int i = 0; i < MAXDIR; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
inline Dual operator-(const Dual& a, const Dual& b) {
    Dual r;
    r.v = a.v - b.v;
    for (int i = 0; i < MAXDIR; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
inline Dual operator-(const Dual& a) {
    Dual r;
    r.v = -a.v;
    for (int i = 0; i < MAXDIR; ++i) r.d[i] = -a.d[i];
    return r;
}
inline Dual operator*(const Dual& a, const Dual& b) {
    Dual r;
    r.v = a.v * b.v;
    for (int i = 0; i < MAXDIR; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
inline Dual operator/(const Dual& a, const Dual& b) {
    Dual r;
    const double inv = 1.0 / b.v;
    r.v = a.v * inv;
    for (int i = 0; i < MAXDIR; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
inline Dual& operator+=(Dual& a, const Dual& b) { return a = a + b; }
inline Dual& operator-=(Dual& a, const Dual& b) { return a = a - b; }
inline Dual sin(const Dual& a) {
    Dual r;
    r.v = std::sin(a.v);
    const double c = std::cos(a.v);
    for (int i = 0; i < MAXDIR; ++i) r.d[i] = c * a.d[i];
    return r;
}
inline Dual cos(const Dual& a) {
    Dual r;
    r.v = std::cos(a.v);
    const double s = -std::sin(a.v);
    for (int i = 0; i < MAXDIR; ++i) r.d[i] = s * a.d[i];
    return r;
}
inline Dual sqrt(const Dual& a) {
    Dual r;
    r.v = std::sqrt(a.v);
    const double h = 0.5 / r.v;
    for (int i = 0; i < MAXDIR; ++i) r.d[i] = h * a.d[i];
    return r;
}
inline double value(double x) { return x; }
inline double value(const Dual& x) { return x.v; }
inline double partial(double, int) { return 0.0; }
inline double partial(const Dual& x, int i) { return x.d[i]; }

using std::cos;
using std::sin;
using std::sqrt;

template <typename S>
struct Vec3 {
    S x, y, z;
    Vec3() : x(0.0), y(0.0), z(0.0) {}
    Vec3(S a, S b, S c) : x(a), y(b), z(c) {}
    S& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const S& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <typename S>
Vec3<S> operator+(const Vec3<S>& a, const Vec3<S>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename S>
Vec3<S> operator-(const Vec3<S>& a, const Vec3<S>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename S>
Vec3<S> operator-(const Vec3<S>& a) { return {-a.x, -a.y, -a.z}; }
template <typename S>
Vec3<S> operator*(const S& s, const Vec3<S>& a) { return {s * a.x, s * a.y, s * a.z}; }
template <typename S>
Vec3<S> operator*(const Vec3<S>& a, const S& s) { return {s * a.x, s * a.y, s * a.z}; }
template <typename S>
S dot(const Vec3<S>& a, const Vec3<S>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename S>
Vec3<S> cross(const Vec3<S>& a, const Vec3<S>& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename S>
Vec3<S> vec3_from(const double* p) { return {S(p[0]), S(p[1]), S(p[2])}; }

template <typename S>
struct Mat3 {
    S m[3][3];
    Mat3() {
        for (auto& r : m)
            for (auto& e : r) e = S(0.0);
    }
    static Mat3 identity() {
        Mat3 I;
        I.m[0][0] = I.m[1][1] = I.m[2][2] = S(1.0);
        return I;
    }
    static Mat3 from_rowmajor(const double* p) {
        Mat3 M;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) M.m[i][j] = S(p[3 * i + j]);
        return M;
    }
    Mat3 transpose() const {
        Mat3 T;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) T.m[i][j] = m[j][i];
        return T;
    }
};
template <typename S>
Mat3<S> operator*(const Mat3<S>& A, const Mat3<S>& B) {
    Mat3<S> C;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
    return C;
}
template <typename S>
Mat3<S> operator+(const Mat3<S>& A, const Mat3<S>& B) {
    Mat3<S> C;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C.m[i][j] = A.m[i][j] + B.m[i][j];
    return C;
}
template <typename S>
Vec3<S> operator*(const Mat3<S>& A, const Vec3<S>& v) {
    return {A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z, A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
            A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z};
}
// skew3 (upright_core/include/upright_core/util.h:25-35)
template <typename S>
Mat3<S> skew3(const Vec3<S>& v) {
    Mat3<S> M;
    M.m[0][1] = -v.z;
    M.m[0][2] = v.y;
    M.m[1][0] = v.z;
    M.m[1][2] = -v.x;
    M.m[2][0] = -v.y;
    M.m[2][1] = v.x;
    return M;
}
// Rodrigues rotation about a unit axis
template <typename S>
Mat3<S> axis_angle(const Vec3<S>& u, const S& th) {
    const S c = cos(th), s = sin(th), t = S(1.0) - c;
    Mat3<S> R;
    R.m[0][0] = c + t * u.x * u.x;
    R.m[0][1] = t * u.x * u.y - s * u.z;
    R.m[0][2] = t * u.x * u.z + s * u.y;
    R.m[1][0] = t * u.x * u.y + s * u.z;
    R.m[1][1] = c + t * u.y * u.y;
    R.m[1][2] = t * u.y * u.z - s * u.x;
    R.m[2][0] = t * u.x * u.z - s * u.y;
    R.m[2][1] = t * u.y * u.z + s * u.x;
    R.m[2][2] = c + t * u.z * u.z;
    return R;
}

}  // namespace orc
