// ORACLE (test infrastructure only) — minimal fp64 linear algebra for the back-end restatement.
// The reference expresses all of this in Eigen (absent from the build container); these helpers are the
// self-contained replacement SURVEY.md §8c calls for.  Plain loops, no FMA tricks, row-major storage.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace ob {

struct V3 {
    double x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    double &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, double s) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalized(V3 a) { return a / norm(a); }

struct M3 {
    double m[3][3];
    M3() { std::memset(m, 0, sizeof(m)); }
    static M3 identity() {
        M3 r;
        r.m[0][0] = r.m[1][1] = r.m[2][2] = 1.0;
        return r;
    }
    double operator()(int i, int j) const { return m[i][j]; }
    double &operator()(int i, int j) { return m[i][j]; }
};
inline M3 operator*(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += a.m[i][k] * b.m[k][j];
            r.m[i][j] = s;
        }
    return r;
}
inline V3 operator*(const M3 &a, V3 v) {
    return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
inline M3 operator*(double s, const M3 &a) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[i][j] = s * a.m[i][j];
    return r;
}
inline M3 operator+(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j];
    return r;
}
inline M3 operator-(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] - b.m[i][j];
    return r;
}
inline M3 operator-(const M3 &a) { return (-1.0) * a; }
inline M3 transpose(const M3 &a) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i];
    return r;
}
// Utility::skewSymmetric, vins_estimator/src/utility/utility.h:30-38
inline M3 skew(V3 q) {
    M3 r;
    r.m[0][1] = -q.z;
    r.m[0][2] = q.y;
    r.m[1][0] = q.z;
    r.m[1][2] = -q.x;
    r.m[2][0] = -q.y;
    r.m[2][1] = q.x;
    return r;
}

// Hamilton quaternion, Eigen convention (w, x, y, z) constructor order.
struct Q {
    double w, x, y, z;
    Q() : w(1), x(0), y(0), z(0) {}
    Q(double w_, double x_, double y_, double z_) : w(w_), x(x_), y(y_), z(z_) {}
    V3 vec() const { return {x, y, z}; }
};
inline Q operator*(const Q &a, const Q &b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
// Eigen::Quaternion::inverse(): conjugate / squaredNorm
inline Q inverse(const Q &q) {
    double n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
    return {q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
}
inline Q normalized(const Q &q) {
    double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    return {q.w / n, q.x / n, q.y / n, q.z / n};
}
// Eigen::Quaternion::toRotationMatrix()
inline M3 to_rot(const Q &q) {
    M3 r;
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    r.m[0][0] = 1 - (tyy + tzz);
    r.m[0][1] = txy - twz;
    r.m[0][2] = txz + twy;
    r.m[1][0] = txy + twz;
    r.m[1][1] = 1 - (txx + tzz);
    r.m[1][2] = tyz - twx;
    r.m[2][0] = txz - twy;
    r.m[2][1] = tyz + twx;
    r.m[2][2] = 1 - (txx + tyy);
    return r;
}
// Eigen: q * v == toRotationMatrix-equivalent rotation (v + 2w (u x v) + 2 u x (u x v))
inline V3 rot(const Q &q, V3 v) {
    V3 u = q.vec();
    V3 uv = cross(u, v);
    uv = uv + uv;
    return v + q.w * uv + cross(u, uv);
}
// Utility::deltaQ (utility.h:16-28): first-order (1, theta/2), NOT normalised
inline Q delta_q(V3 theta) { return {1.0, theta.x / 2.0, theta.y / 2.0, theta.z / 2.0}; }
// Utility::Qleft / Qright bottom-right 3x3 (utility.h:47-64)
inline M3 qleft_br(const Q &q) { return q.w * M3::identity() + skew(q.vec()); }
inline M3 qright_br(const Q &q) { return q.w * M3::identity() - skew(q.vec()); }
inline void qleft4(const Q &q, double L[4][4]) {
    L[0][0] = q.w;
    L[0][1] = -q.x;
    L[0][2] = -q.y;
    L[0][3] = -q.z;
    L[1][0] = q.x;
    L[2][0] = q.y;
    L[3][0] = q.z;
    M3 b = qleft_br(q);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) L[i + 1][j + 1] = b.m[i][j];
}
inline void qright4(const Q &q, double R[4][4]) {
    R[0][0] = q.w;
    R[0][1] = -q.x;
    R[0][2] = -q.y;
    R[0][3] = -q.z;
    R[1][0] = q.x;
    R[2][0] = q.y;
    R[3][0] = q.z;
    M3 b = qright_br(q);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i + 1][j + 1] = b.m[i][j];
}

// ---- dynamic dense matrices (row-major) --------------------------------------------------------
struct Mat {
    int r = 0, c = 0;
    std::vector<double> a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
    double &operator()(int i, int j) { return a[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
};

// In-place lower Cholesky A = L L^T (A symmetric, uses lower triangle). Returns false if not PD.
inline bool cholesky_lower(Mat &A) {
    const int n = A.r;
    for (int j = 0; j < n; j++) {
        double d = A(j, j);
        for (int k = 0; k < j; k++) d -= A(j, k) * A(j, k);
        if (!(d > 0.0)) return false;
        d = std::sqrt(d);
        A(j, j) = d;
        for (int i = j + 1; i < n; i++) {
            double s = A(i, j);
            for (int k = 0; k < j; k++) s -= A(i, k) * A(j, k);
            A(i, j) = s / d;
        }
    }
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) A(i, j) = 0.0;
    return true;
}
inline void chol_solve(const Mat &L, std::vector<double> &b) {  // solves L L^T x = b in place
    const int n = L.r;
    for (int i = 0; i < n; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= L(i, k) * b[k];
        b[i] = s / L(i, i);
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = b[i];
        for (int k = i + 1; k < n; k++) s -= L(k, i) * b[k];
        b[i] = s / L(i, i);
    }
}

// Inverse by Gauss-Jordan with partial pivoting (stands in for Eigen's PartialPivLU-based inverse()).
inline bool invert(const Mat &A, Mat &inv) {
    const int n = A.r;
    Mat w(n, 2 * n);
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) w(i, j) = A(i, j);
        w(i, n + i) = 1.0;
    }
    for (int col = 0; col < n; col++) {
        int piv = col;
        double best = std::fabs(w(col, col));
        for (int i = col + 1; i < n; i++)
            if (std::fabs(w(i, col)) > best) {
                best = std::fabs(w(i, col));
                piv = i;
            }
        if (best == 0.0) return false;
        if (piv != col)
            for (int j = 0; j < 2 * n; j++) std::swap(w(col, j), w(piv, j));
        const double d = w(col, col);
        for (int j = 0; j < 2 * n; j++) w(col, j) /= d;
        for (int i = 0; i < n; i++) {
            if (i == col) continue;
            const double f = w(i, col);
            if (f == 0.0) continue;
            for (int j = 0; j < 2 * n; j++) w(i, j) -= f * w(col, j);
        }
    }
    inv = Mat(n, n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) inv(i, j) = w(i, n + j);
    return true;
}

// Symmetric eigen-decomposition A = V diag(w) V^T by Householder tridiagonalisation + implicit-shift QL (the
// textbook tred2 / tqli pair; stands in for Eigen::SelfAdjointEigenSolver, which uses the same two stages).
// Eigenvector order/sign are not unique, callers use invariants only.
inline void sym_eig(const Mat &Ain, std::vector<double> &w, Mat &V) {
    const int n = Ain.r;
    V = Ain;
    w.assign(n, 0.0);
    if (n == 0) return;
    std::vector<double> e(n, 0.0);
    Mat &a = V;
    std::vector<double> &d = w;
    // ---- tred2 ----
    for (int i = n - 1; i >= 1; i--) {
        const int l = i - 1;
        double h = 0.0, scale = 0.0;
        if (l > 0) {
            for (int k = 0; k <= l; k++) scale += std::fabs(a(i, k));
            if (scale == 0.0) {
                e[i] = a(i, l);
            } else {
                for (int k = 0; k <= l; k++) {
                    a(i, k) /= scale;
                    h += a(i, k) * a(i, k);
                }
                double f = a(i, l);
                double g = (f >= 0.0 ? -std::sqrt(h) : std::sqrt(h));
                e[i] = scale * g;
                h -= f * g;
                a(i, l) = f - g;
                f = 0.0;
                for (int j = 0; j <= l; j++) {
                    a(j, i) = a(i, j) / h;
                    g = 0.0;
                    for (int k = 0; k <= j; k++) g += a(j, k) * a(i, k);
                    for (int k = j + 1; k <= l; k++) g += a(k, j) * a(i, k);
                    e[j] = g / h;
                    f += e[j] * a(i, j);
                }
                const double hh = f / (h + h);
                for (int j = 0; j <= l; j++) {
                    f = a(i, j);
                    e[j] = g = e[j] - hh * f;
                    for (int k = 0; k <= j; k++) a(j, k) -= (f * e[k] + g * a(i, k));
                }
            }
        } else {
            e[i] = a(i, l);
        }
        d[i] = h;
    }
    d[0] = 0.0;
    e[0] = 0.0;
    for (int i = 0; i < n; i++) {
        const int l = i - 1;
        if (d[i] != 0.0) {
            for (int j = 0; j <= l; j++) {
                double g = 0.0;
                for (int k = 0; k <= l; k++) g += a(i, k) * a(k, j);
                for (int k = 0; k <= l; k++) a(k, j) -= g * a(k, i);
            }
        }
        d[i] = a(i, i);
        a(i, i) = 1.0;
        for (int j = 0; j <= l; j++) a(j, i) = a(i, j) = 0.0;
    }
    // ---- tqli ----
    for (int i = 1; i < n; i++) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    for (int l = 0; l < n; l++) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; m++) {
                const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
                if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
            }
            if (m != l) {
                if (iter++ == 200) break;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; i--) {
                    double f = s * e[i];
                    const double bb = c * e[i];
                    e[i + 1] = (r = std::hypot(f, g));
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * bb;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - bb;
                    for (int k = 0; k < n; k++) {
                        f = a(k, i + 1);
                        a(k, i + 1) = s * a(k, i) + c * f;
                        a(k, i) = c * a(k, i) - s * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
}

}  // namespace ob
