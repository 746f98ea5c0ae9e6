// vecmath.h — small FP64 vector / affine-transform kit for the host side.
//
// The reference does all of this with vendored Eigen 3.2.2.  The host must hand the
// device *bit-identical* transforms, so the handful of Eigen operations the reference
// instantiates are restated here with the SAME floating-point association order
// (checked bit-for-bit against the reference object graph in
// tests/test_host_flatten.py):
//   Transform::translate / scale / rotate   eigen/Eigen/src/Geometry/Transform.h:838-843,784-790,882-886
//   Transform::inverse (Affine)             eigen/Eigen/src/Geometry/Transform.h:1124-1151
//   3x3 inverse by cofactors                eigen/Eigen/src/LU/Inverse.h:130-159
//   4x4 determinant (Costabel)              eigen/Eigen/src/LU/Determinant.h:18-23,86-97
//   AngleAxis::toRotationMatrix             eigen/Eigen/src/Geometry/AngleAxis.h:204-229
//   Vector4d redux (SSE2 packet order)      eigen/Eigen/src/Core/Redux.h:131-136,299-305
// Compile WITHOUT FMA contraction (-ffp-contract=off): the reference build has none.
#pragma once
#include <cmath>

namespace as2 {

struct Vec3 {
    double x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    bool isZero() const { return x == 0 && y == 0 && z == 0; }
};

inline Vec3 operator+(const Vec3& a, const Vec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(double s, const Vec3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 operator/(const Vec3& a, double s) { return {a.x / s, a.y / s, a.z / s}; }

// Eigen's unrolled scalar reduction of a 3-vector splits [0,1) + [1,3).
inline double sum3(double a, double b, double c) { return a + (b + c); }
inline double norm3(const Vec3& v) { return std::sqrt(sum3(v.x * v.x, v.y * v.y, v.z * v.z)); }
inline Vec3 normalized3(const Vec3& v) { return v / norm3(v); }
inline Vec3 cross(const Vec3& a, const Vec3& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Homogeneous 4-vector (points w=1, directions w=0), like the reference's Vector4d.
struct Vec4 {
    double x = 0, y = 0, z = 0, w = 0;
    Vec4() = default;
    Vec4(double a, double b, double c, double d) : x(a), y(b), z(c), w(d) {}
    static Vec4 point(const Vec3& p) { return {p.x, p.y, p.z, 1.0}; }
    static Vec4 dir(const Vec3& d) { return {d.x, d.y, d.z, 0.0}; }
    Vec3 head() const { return {x, y, z}; }
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    bool isZero() const { return x == 0 && y == 0 && z == 0 && w == 0; }
};
inline Vec4 operator+(const Vec4& a, const Vec4& b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline Vec4 operator-(const Vec4& a, const Vec4& b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline Vec4 operator*(double s, const Vec4& a) { return {s * a.x, s * a.y, s * a.z, s * a.w}; }
inline Vec4 operator/(const Vec4& a, double s) { return {a.x / s, a.y / s, a.z / s, a.w / s}; }
// Two 2-wide packets added lane-wise, then a horizontal add: (x0+x2)+(x1+x3).
inline double sum4(double a, double b, double c, double d) { return (a + c) + (b + d); }
inline double dot4(const Vec4& a, const Vec4& b) { return sum4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline double norm4(const Vec4& v) { return std::sqrt(dot4(v, v)); }
inline Vec4 normalized4(const Vec4& v) { return v / norm4(v); }

// Affine 3D transform: rows 0..2 of a 4x4 whose last row is 0 0 0 1.
struct Affine {
    double m[3][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}};

    static Affine identity() { return Affine(); }
    void setIdentity() { *this = Affine(); }

    // T := T * Translate(v):  translation += linear * v
    void translate(const Vec3& v) {
        for (int i = 0; i < 3; i++)
            m[i][3] += (m[i][0] * v.x + m[i][1] * v.y) + m[i][2] * v.z;
    }
    // T := T * Scale(v): column j of the linear part times v[j]
    void scale(const Vec3& v) {
        for (int i = 0; i < 3; i++) {
            m[i][0] = m[i][0] * v.x;
            m[i][1] = m[i][1] * v.y;
            m[i][2] = m[i][2] * v.z;
        }
    }
    // T := T * Rot(angle, axis) with the rotation matrix built like AngleAxis::toRotationMatrix
    void rotate(double angle, const Vec3& axis) {
        double s = std::sin(angle), c = std::cos(angle);
        Vec3 sin_axis = s * axis;
        Vec3 cos1_axis = (1.0 - c) * axis;
        double R[3][3];
        double tmp;
        tmp = cos1_axis.x * axis.y;
        R[0][1] = tmp - sin_axis.z;
        R[1][0] = tmp + sin_axis.z;
        tmp = cos1_axis.x * axis.z;
        R[0][2] = tmp + sin_axis.y;
        R[2][0] = tmp - sin_axis.y;
        tmp = cos1_axis.y * axis.z;
        R[1][2] = tmp - sin_axis.x;
        R[2][1] = tmp + sin_axis.x;
        R[0][0] = cos1_axis.x * axis.x + c;
        R[1][1] = cos1_axis.y * axis.y + c;
        R[2][2] = cos1_axis.z * axis.z + c;
        double L[3][3];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                L[i][j] = (m[i][0] * R[0][j] + m[i][1] * R[1][j]) + m[i][2] * R[2][j];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) m[i][j] = L[i][j];
    }

    // Affine inverse: cofactor inverse of the linear part, translation = (-Linv) * t.
    Affine inverse() const {
        auto cof = [&](int i, int j) {
            int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1];
        };
        double c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
        double det = sum3(c0 * m[0][0], c1 * m[1][0], c2 * m[2][0]);
        double invdet = 1.0 / det;
        Affine r;
        r.m[0][0] = c0 * invdet;
        r.m[0][1] = c1 * invdet;
        r.m[0][2] = c2 * invdet;
        r.m[1][0] = cof(0, 1) * invdet;
        r.m[1][1] = cof(1, 1) * invdet;
        r.m[1][2] = cof(2, 1) * invdet;
        r.m[2][0] = cof(0, 2) * invdet;
        r.m[2][1] = cof(1, 2) * invdet;
        r.m[2][2] = cof(2, 2) * invdet;
        for (int i = 0; i < 3; i++)
            r.m[i][3] = ((-r.m[i][0]) * m[0][3] + (-r.m[i][1]) * m[1][3]) + (-r.m[i][2]) * m[2][3];
        return r;
    }

    // determinant of the full 4x4 (last row 0 0 0 1), Costabel's 30-multiply form
    double determinant() const {
        auto M = [&](int r, int c) -> double {
            if (r < 3) return m[r][c];
            return c == 3 ? 1.0 : 0.0;
        };
        auto h = [&](int j, int k, int p, int q) {
            return (M(j, 0) * M(k, 1) - M(k, 0) * M(j, 1)) * (M(p, 2) * M(q, 3) - M(q, 2) * M(p, 3));
        };
        return h(0, 1, 2, 3) - h(0, 2, 1, 3) + h(0, 3, 1, 2) + h(1, 2, 0, 3) - h(1, 3, 0, 2) + h(2, 3, 0, 1);
    }

    // T * v for a homogeneous vector: rows 0..2 via the 3x4 block, w copied.
    Vec4 apply(const Vec4& v) const {
        Vec4 r;
        for (int i = 0; i < 3; i++)
            r[i] = ((m[i][0] * v.x + m[i][1] * v.y) + m[i][2] * v.z) + m[i][3] * v.w;
        r.w = v.w;
        return r;
    }
};

}  // namespace as2
