// Small dense fp64 cores of the estimator initialisation, written once as __host__ __device__
// functions on plain arrays: the kernels in initial.cu call them from single threads / warps around their parallel
// reductions, and tests/native/init_math_shim.cpp compiles THE SAME header with g++ so the arithmetic can be checked
// against the numpy oracle on a machine without a GPU (test infrastructure only — liblfvio.so never runs these on the
// host).  No Eigen: Jacobi eigen/singular value iterations and Householder least squares are spelled out.
//
// Reference code these follow (paths under vins_estimator/src/):
//   triangulate_point        initial/initial_sfm.cpp:6-23, initial/solve_5pts.cpp:363-378
//   decompose_essential      initial/solve_5pts.cpp:337-361
//   epnp_*                   pnp_solver.cpp:46-440
//   alignment_pair           initial/initial_aligment.cpp:73-123, 146-180
//   tangent_basis            initial/initial_aligment.cpp:40-54
//   gyro_bias_pair           initial/initial_aligment.cpp:13-25
//   ba_residual_jac          initial/initial_sfm.h:26-66 (+ ceres::QuaternionParameterization)
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define LFVIO_HD __host__ __device__ inline
#else
#define LFVIO_HD inline
#endif

namespace lfvio {
namespace im {

// ------------------------------------------------------------------------------------------------------------
// generic small linear algebra
// ------------------------------------------------------------------------------------------------------------
// Cyclic Jacobi on the symmetric n x n matrix A (row-major, destroyed: its diagonal ends up holding the eigenvalues);
// V (n x n row-major) receives the eigenvectors as COLUMNS.
LFVIO_HD void jacobi_sym(int n, double *A, double *V) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 40; sweep++) {
        double off = 0, dia = 0;
        for (int i = 0; i < n; i++) {
            dia += fabs(A[i * n + i]);
            for (int j = i + 1; j < n; j++) off += fabs(A[i * n + j]);
        }
        if (off <= 1e-300 || off <= 1e-18 * dia) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = A[p * n + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) {  // A <- A J
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {  // A <- J^T A
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
}

// One-sided (Hestenes) Jacobi SVD of the m x n matrix A (row-major, in place): on return the columns of A are
// u_i * sigma_i (mutually orthogonal) and V (n x n row-major) holds the right singular vectors as columns.  No A^T A.
LFVIO_HD void onesided_jacobi(int m, int n, double *A, double *V) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 40; sweep++) {
        double off = 0;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double al = 0, be = 0, ga = 0;
                for (int k = 0; k < m; k++) {
                    al += A[k * n + p] * A[k * n + p];
                    be += A[k * n + q] * A[k * n + q];
                    ga += A[k * n + p] * A[k * n + q];
                }
                if (ga == 0.0 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
                off = fmax(off, fabs(ga) / sqrt(al * be));
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < m; k++) {
                    const double ap = A[k * n + p], aq = A[k * n + q];
                    A[k * n + p] = c * ap - s * aq;
                    A[k * n + q] = s * ap + c * aq;
                }
                for (int k = 0; k < n; k++) {
                    const double vp = V[k * n + p], vq = V[k * n + q];
                    V[k * n + p] = c * vp - s * vq;
                    V[k * n + q] = s * vp + c * vq;
                }
            }
        if (off < 1e-15) break;
    }
}

LFVIO_HD void mat3_mul(const double *A, const double *B, double *C) {  // C = A B (row-major 3 x 3; C must not alias)
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
LFVIO_HD void mat3_mul_t(const double *A, const double *B, double *C) {  // C = A B^T
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
LFVIO_HD void mat3_tmul(const double *A, const double *B, double *C) {  // C = A^T B
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
LFVIO_HD void mat3_vec(const double *A, const double *v, double *o) {
    for (int i = 0; i < 3; i++) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
LFVIO_HD void mat3_tvec(const double *A, const double *v, double *o) {
    for (int i = 0; i < 3; i++) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}
LFVIO_HD double det3(const double *A) {
    return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}
LFVIO_HD void inv3(const double *A, double *I) {
    const double d = 1.0 / det3(A);
    I[0] = (A[4] * A[8] - A[5] * A[7]) * d, I[1] = (A[2] * A[7] - A[1] * A[8]) * d, I[2] = (A[1] * A[5] - A[2] * A[4]) * d;
    I[3] = (A[5] * A[6] - A[3] * A[8]) * d, I[4] = (A[0] * A[8] - A[2] * A[6]) * d, I[5] = (A[2] * A[3] - A[0] * A[5]) * d;
    I[6] = (A[3] * A[7] - A[4] * A[6]) * d, I[7] = (A[1] * A[6] - A[0] * A[7]) * d, I[8] = (A[0] * A[4] - A[1] * A[3]) * d;
}

// SVD of a 3 x 3 matrix M = U diag(s) V^T with s descending; U, V row-major with the singular vectors as columns.  The
// left vector of a (numerically) zero smallest singular value is completed as u0 x u1.
LFVIO_HD void svd3(const double *M, double *U, double *s, double *V) {
    double A[9], W[9];
    for (int i = 0; i < 9; i++) A[i] = M[i];
    onesided_jacobi(3, 3, A, W);
    double nn[3];
    int ord[3] = {0, 1, 2};
    for (int c = 0; c < 3; c++) nn[c] = sqrt(A[c] * A[c] + A[3 + c] * A[3 + c] + A[6 + c] * A[6 + c]);
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2 - a; b++)
            if (nn[ord[b]] < nn[ord[b + 1]]) {
                const int t = ord[b];
                ord[b] = ord[b + 1];
                ord[b + 1] = t;
            }
    for (int c = 0; c < 3; c++) {
        const int o = ord[c];
        s[c] = nn[o];
        for (int r = 0; r < 3; r++) {
            V[3 * r + c] = W[3 * r + o];
            U[3 * r + c] = nn[o] > 0 ? A[3 * r + o] / nn[o] : 0.0;
        }
    }
    if (!(s[2] > 1e-13 * s[0])) {
        U[2] = U[3] * U[7] - U[6] * U[4];
        U[5] = U[6] * U[1] - U[0] * U[7];
        U[8] = U[0] * U[4] - U[3] * U[1];
    }
}

// Least squares min |A x - b| for a full-rank m x n system (m >= n, n <= 5; A row-major, A and b destroyed) by
// Householder QR — what Eigen's colPivHouseholderQr().solve returns for full column rank.
LFVIO_HD void lstsq_small(int m, int n, double *A, double *b, double *x) {
    for (int k = 0; k < n; k++) {
        double nrm = 0;
        for (int i = k; i < m; i++) nrm += A[i * n + k] * A[i * n + k];
        nrm = sqrt(nrm);
        if (nrm == 0.0) continue;
        const double alpha = A[k * n + k] > 0 ? -nrm : nrm;
        double v[6];
        for (int i = k; i < m; i++) v[i] = A[i * n + k];
        v[k] -= alpha;
        double vv = 0;
        for (int i = k; i < m; i++) vv += v[i] * v[i];
        if (vv == 0.0) continue;
        for (int j = k; j < n; j++) {
            double d = 0;
            for (int i = k; i < m; i++) d += v[i] * A[i * n + j];
            d = 2.0 * d / vv;
            for (int i = k; i < m; i++) A[i * n + j] -= d * v[i];
        }
        double d = 0;
        for (int i = k; i < m; i++) d += v[i] * b[i];
        d = 2.0 * d / vv;
        for (int i = k; i < m; i++) b[i] -= d * v[i];
    }
    for (int k = n - 1; k >= 0; k--) {
        double s = b[k];
        for (int j = k + 1; j < n; j++) s -= A[k * n + j] * x[j];
        x[k] = A[k * n + k] != 0.0 ? s / A[k * n + k] : 0.0;
    }
}

// Eigen::Quaternion (w, x, y, z) helpers
LFVIO_HD void quat_mul(const double *a, const double *b, double *o) {
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
    o[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
}
LFVIO_HD void quat_to_rot(const double *q, double *R) {  // toRotationMatrix(), no normalisation
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y,
                 tzz = tz * z;
    R[0] = 1 - (tyy + tzz), R[1] = txy - twz, R[2] = txz + twy;
    R[3] = txy + twz, R[4] = 1 - (txx + tzz), R[5] = tyz - twx;
    R[6] = txz - twy, R[7] = tyz + twx, R[8] = 1 - (txx + tyy);
}
LFVIO_HD void rot_to_quat(const double *R, double *q) {  // Eigen::Quaternion(Matrix3)
    const double t = R[0] + R[4] + R[8];
    if (t > 0) {
        const double s = sqrt(t + 1.0) * 2;
        q[0] = 0.25 * s, q[1] = (R[7] - R[5]) / s, q[2] = (R[2] - R[6]) / s, q[3] = (R[3] - R[1]) / s;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        const double s = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0) * 2;
        q[0] = (R[3 * k + j] - R[3 * j + k]) / s;
        q[1 + i] = 0.25 * s;
        q[1 + j] = (R[3 * j + i] + R[3 * i + j]) / s;
        q[1 + k] = (R[3 * k + i] + R[3 * i + k]) / s;
    }
}

// ------------------------------------------------------------------------------------------------------------
// two-view geometry
// ------------------------------------------------------------------------------------------------------------
// P0, P1: 3 x 4 row-major [R | t] (camera <- world); p0, p1: bearings.  X: the triangulated point.
LFVIO_HD void triangulate_point(const double *P0, const double *P1, const double *p0, const double *p1, double *X) {
    double D[16], V[16];
    for (int c = 0; c < 4; c++) {
        D[c] = p0[0] * P0[8 + c] - p0[2] * P0[c];
        D[4 + c] = p0[1] * P0[8 + c] - p0[2] * P0[4 + c];
        D[8 + c] = p1[0] * P1[8 + c] - p1[2] * P1[c];
        D[12 + c] = p1[1] * P1[8 + c] - p1[2] * P1[4 + c];
    }
    onesided_jacobi(4, 4, D, V);
    int bi = 0;
    double best = 1e300;
    for (int c = 0; c < 4; c++) {
        const double nn = D[c] * D[c] + D[4 + c] * D[4 + c] + D[8 + c] * D[8 + c] + D[12 + c] * D[12 + c];
        if (nn < best) best = nn, bi = c;
    }
    const double w = V[12 + bi];
    X[0] = V[bi] / w, X[1] = V[4 + bi] / w, X[2] = V[8 + bi] / w;
}

LFVIO_HD void decompose_essential(const double *E, double *R1, double *R2, double *t) {
    double U[9], s[3], V[9], T[9];
    svd3(E, U, s, V);
    if (det3(U) < 0)
        for (int i = 0; i < 9; i++) U[i] = -U[i];
    if (det3(V) < 0)
        for (int i = 0; i < 9; i++) V[i] = -V[i];
    const double W[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1}, Wt[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
    mat3_mul(U, W, T);
    mat3_mul_t(T, V, R1);
    mat3_mul(U, Wt, T);
    mat3_mul_t(T, V, R2);
    t[0] = U[2], t[1] = U[5], t[2] = U[8];
}

// One correspondence of one recoverPose candidate (solve_5pts.cpp:430-448): 1 if both bearings look towards the point.
LFVIO_HD int cheirality(const double *R, const double *t, const double *b1, const double *b2) {
    const double P0[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    double P1[12], Q[3], tq[3];
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) P1[4 * r + c] = R[3 * r + c];
        P1[4 * r + 3] = t[r];
    }
    triangulate_point(P0, P1, b1, b2, Q);
    const double l_dot = (b1[0] * Q[0] + b1[1] * Q[1] + b1[2] * Q[2]) / sqrt(Q[0] * Q[0] + Q[1] * Q[1] + Q[2] * Q[2]);
    mat3_vec(R, Q, tq);
    tq[0] += t[0], tq[1] += t[1], tq[2] += t[2];
    const double r_dot = (b2[0] * tq[0] + b2[1] * tq[1] + b2[2] * tq[2]) / sqrt(tq[0] * tq[0] + tq[1] * tq[1] + tq[2] * tq[2]);
    return (l_dot > 0 && r_dot > 0) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------------------
// EPnP (pnp_solver.cpp) — the O(1) cores; the sums over the correspondences are the caller's (parallel) job.
// ------------------------------------------------------------------------------------------------------------
// choose_control_points (:46-78): mean = centroid of the world points, cov = sum (pw - mean)(pw - mean)^T.  cws: 4 x 3.
LFVIO_HD void epnp_control_points(int n, const double *mean, const double *cov, double *cws) {
    double A[9], V[9];
    for (int i = 0; i < 9; i++) A[i] = cov[i];
    jacobi_sym(3, A, V);
    int ord[3] = {0, 1, 2};
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2 - a; b++)
            if (A[4 * ord[b]] < A[4 * ord[b + 1]]) {
                const int t = ord[b];
                ord[b] = ord[b + 1];
                ord[b + 1] = t;
            }
    for (int k = 0; k < 3; k++) cws[k] = mean[k];
    for (int i = 1; i < 4; i++) {
        const int o = ord[i - 1];
        const double ev = A[4 * o] > 0 ? A[4 * o] : 0.0;
        double k = sqrt(ev / n);
        // sign convention for the principal direction (an SVD implementation's free choice that the noisy-data result
        // depends on): its component of largest magnitude is positive — shared with oracle/initial.py
        int big = 0;
        for (int c = 1; c < 3; c++)
            if (fabs(V[3 * c + o]) > fabs(V[3 * big + o])) big = c;
        if (V[3 * big + o] < 0) k = -k;
        for (int c = 0; c < 3; c++) cws[3 * i + c] = mean[c] + k * V[3 * c + o];
    }
}
// compute_barycentric_coordinates (:79-101): CC_inv from the control points, then per point
LFVIO_HD void epnp_cc_inv(const double *cws, double *CCi) {
    double CC[9];
    for (int i = 0; i < 3; i++)
        for (int j = 1; j < 4; j++) CC[3 * i + j - 1] = cws[3 * j + i] - cws[i];
    inv3(CC, CCi);
}
LFVIO_HD void epnp_alphas(const double *cws, const double *CCi, const double *pw, double *al) {
    const double d[3] = {pw[0] - cws[0], pw[1] - cws[1], pw[2] - cws[2]};
    for (int j = 0; j < 3; j++) al[1 + j] = CCi[3 * j] * d[0] + CCi[3 * j + 1] * d[1] + CCi[3 * j + 2] * d[2];
    al[0] = 1.0 - al[1] - al[2] - al[3];
}
// the two rows of M of one correspondence (:280-293, fx = fy = 1, cx = cy = 0) contribute al_j al_k * G to block (j, k)
// of M^T M with G = [[1, 0, -x], [0, 1, -y], [-x, -y, x^2 + y^2]], x = u / w, y = v / w.
LFVIO_HD double epnp_mtm_entry(const double *al, const double *us, int r, int c) {
    const double x = (0.0 - us[0]) / us[2], y = (0.0 - us[1]) / us[2];
    const int j = r / 3, a = r % 3, k = c / 3, b = c % 3;
    double g;
    if (a == 2 && b == 2) g = x * x + y * y;
    else if (a == 2) g = (b == 0) ? x : y;
    else if (b == 2) g = (a == 0) ? x : y;
    else g = (a == b) ? 1.0 : 0.0;
    return al[j] * al[k] * g;
}

struct EpnpCore {
    double Ut[144];     // rows = eigenvectors of M^T M, descending eigenvalue (svd.matrixU().transpose())
    double betas[3][4];  // after gauss_newton, for find_betas_0 / _1 / _2
};

LFVIO_HD void epnp_gauss_newton(const double *L, const double *rho, double *b) {  // :384-434
    for (int it = 0; it < 15; it++) {
        double A[24], r[6], x[4];
        for (int i = 0; i < 6; i++) {
            const double *l = L + 10 * i;
            A[4 * i + 0] = 2 * l[0] * b[0] + l[1] * b[1] + l[3] * b[2] + l[6] * b[3];
            A[4 * i + 1] = l[1] * b[0] + 2 * l[2] * b[1] + l[4] * b[2] + l[7] * b[3];
            A[4 * i + 2] = l[3] * b[0] + l[4] * b[1] + 2 * l[5] * b[2] + l[8] * b[3];
            A[4 * i + 3] = l[6] * b[0] + l[7] * b[1] + l[8] * b[2] + 2 * l[9] * b[3];
            r[i] = rho[i] - (l[0] * b[0] * b[0] + l[1] * b[0] * b[1] + l[2] * b[1] * b[1] + l[3] * b[0] * b[2] +
                             l[4] * b[1] * b[2] + l[5] * b[2] * b[2] + l[6] * b[0] * b[3] + l[7] * b[1] * b[3] +
                             l[8] * b[2] * b[3] + l[9] * b[3] * b[3]);
        }
        lstsq_small(6, 4, A, r, x);
        for (int i = 0; i < 4; i++) b[i] += x[i];
    }
}

// After the eigen-decomposition of M^T M (MtM: its diagonal holds the eigenvalues, V: eigenvectors as columns): Ut, the
// three beta initialisations and their Gauss-Newton refinements.  cws: 4 x 3.
LFVIO_HD void epnp_core_after_eig(const double *MtM, const double *V, const double *cws, EpnpCore *out) {
    int ord[12];
    for (int i = 0; i < 12; i++) ord[i] = i;
    for (int a = 0; a < 11; a++)
        for (int b = 0; b < 11 - a; b++)
            if (MtM[13 * ord[b]] < MtM[13 * ord[b + 1]]) {
                const int t = ord[b];
                ord[b] = ord[b + 1];
                ord[b + 1] = t;
            }
    for (int r = 0; r < 12; r++)
        for (int c = 0; c < 12; c++) out->Ut[12 * r + c] = V[12 * c + ord[r]];
    const double *Ut = out->Ut;
    // compute_L_6x10 (:98-136)
    double dv[4][6][3], L[60], rho[6];
    for (int i = 0; i < 4; i++) {
        int a = 0, b = 1;
        for (int j = 0; j < 6; j++) {
            for (int k = 0; k < 3; k++) dv[i][j][k] = Ut[12 * (11 - i) + 3 * a + k] - Ut[12 * (11 - i) + 3 * b + k];
            if (++b > 3) {
                a++;
                b = a + 1;
            }
        }
    }
    for (int i = 0; i < 6; i++) {
        auto dot = [&](int p, int q) { return dv[p][i][0] * dv[q][i][0] + dv[p][i][1] * dv[q][i][1] + dv[p][i][2] * dv[q][i][2]; };
        L[10 * i + 0] = dot(0, 0), L[10 * i + 1] = 2.0 * dot(0, 1), L[10 * i + 2] = dot(1, 1), L[10 * i + 3] = 2.0 * dot(0, 2);
        L[10 * i + 4] = 2.0 * dot(1, 2), L[10 * i + 5] = dot(2, 2), L[10 * i + 6] = 2.0 * dot(0, 3), L[10 * i + 7] = 2.0 * dot(1, 3);
        L[10 * i + 8] = 2.0 * dot(2, 3), L[10 * i + 9] = dot(3, 3);
    }
    {  // compute_rho (:137-146)
        const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
        for (int i = 0; i < 6; i++) {
            double s = 0;
            for (int k = 0; k < 3; k++) {
                const double d = cws[3 * pa[i] + k] - cws[3 * pb[i] + k];
                s += d * d;
            }
            rho[i] = s;
        }
    }
    double A[30], r[6], x[5];
    {  // find_betas_0 (:147-175): columns 0, 1, 3, 6
        const int cs[4] = {0, 1, 3, 6};
        for (int i = 0; i < 6; i++) {
            for (int c = 0; c < 4; c++) A[4 * i + c] = L[10 * i + cs[c]];
            r[i] = rho[i];
        }
        lstsq_small(6, 4, A, r, x);
        double *b = out->betas[0];
        if (x[0] < 0) {
            b[0] = sqrt(-x[0]);
            b[1] = -x[1] / b[0], b[2] = -x[2] / b[0], b[3] = -x[3] / b[0];
        } else {
            b[0] = sqrt(x[0]);
            b[1] = x[1] / b[0], b[2] = x[2] / b[0], b[3] = x[3] / b[0];
        }
    }
    {  // find_betas_1 (:177-203): columns 0, 1, 2
        for (int i = 0; i < 6; i++) {
            for (int c = 0; c < 3; c++) A[3 * i + c] = L[10 * i + c];
            r[i] = rho[i];
        }
        lstsq_small(6, 3, A, r, x);
        double *b = out->betas[1];
        if (x[0] < 0) {
            b[0] = sqrt(-x[0]);
            b[1] = (x[2] < 0) ? sqrt(-x[2]) : 0.0;
        } else {
            b[0] = sqrt(x[0]);
            b[1] = (x[2] > 0) ? sqrt(x[2]) : 0.0;
        }
        if (x[1] < 0) b[0] = -b[0];
        b[2] = 0.0, b[3] = 0.0;
    }
    {  // find_betas_2 (:204-231): columns 0..4 (the sign test at :217 writes to B5, i.e. has no effect)
        for (int i = 0; i < 6; i++) {
            for (int c = 0; c < 5; c++) A[5 * i + c] = L[10 * i + c];
            r[i] = rho[i];
        }
        lstsq_small(6, 5, A, r, x);
        double *b = out->betas[2];
        if (x[0] < 0) {
            b[0] = sqrt(-x[0]);
            b[1] = (x[2] < 0) ? sqrt(-x[2]) : 0.0;
        } else {
            b[0] = sqrt(x[0]);
            b[1] = (x[2] > 0) ? sqrt(x[2]) : 0.0;
        }
        b[2] = x[3] / b[0];
        b[3] = 0.0;
    }
    for (int k = 0; k < 3; k++) epnp_gauss_newton(L, rho, out->betas[k]);
}
// MtM: 12 x 12 (destroyed), V: 144 scratch.  Serial form (host shim); the kernels run the Jacobi sweeps on a warp.
LFVIO_HD void epnp_core(double *MtM, const double *cws, EpnpCore *out, double *V) {
    jacobi_sym(12, MtM, V);
    epnp_core_after_eig(MtM, V, cws, out);
}
// compute_ccs (:232-243): ccs 4 x 3
LFVIO_HD void epnp_ccs(const EpnpCore *core, int which, double *ccs) {
    for (int q = 0; q < 12; q++) ccs[q] = 0;
    for (int i = 0; i < 4; i++)
        for (int q = 0; q < 12; q++) ccs[q] += core->betas[which][i] * core->Ut[12 * (11 - i) + q];
}
// estimate_R_and_t (:253-280) from W = sum (pc - pc0)(pw - pw0)^T and the two centroids
LFVIO_HD void epnp_rt_from_w(const double *W, const double *pc0, const double *pw0, double *R, double *T) {
    double U[9], s[3], V[9], t[3];
    svd3(W, U, s, V);
    mat3_mul_t(U, V, R);
    mat3_vec(R, pw0, t);
    for (int k = 0; k < 3; k++) T[k] = pc0[k] - t[k];
}

// ------------------------------------------------------------------------------------------------------------
// visual-inertial alignment
// ------------------------------------------------------------------------------------------------------------
LFVIO_HD void tangent_basis(const double *g0, double *bc /* 3 x 2 row-major */) {
    const double n = sqrt(g0[0] * g0[0] + g0[1] * g0[1] + g0[2] * g0[2]);
    const double a[3] = {g0[0] / n, g0[1] / n, g0[2] / n};
    double tmp[3] = {0, 0, 1};
    if (a[0] == 0.0 && a[1] == 0.0 && a[2] == 1.0) tmp[0] = 1, tmp[2] = 0;
    const double d = a[0] * tmp[0] + a[1] * tmp[1] + a[2] * tmp[2];
    double b[3] = {tmp[0] - a[0] * d, tmp[1] - a[1] * d, tmp[2] - a[2] * d};
    const double bn = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
    for (int k = 0; k < 3; k++) b[k] /= bn;
    const double c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    for (int k = 0; k < 3; k++) bc[2 * k] = b[k], bc[2 * k + 1] = c[k];
}

// tmp_A (6 x (6 + k + 1), row-major with ld = 10) and tmp_b of one consecutive frame pair.  gcols: 3 x k row-major
// (identity for LinearAlignment, lxly for RefineGravity), gfixed: 0 or g0.  Ri, Rj row-major ImageFrame::R.
LFVIO_HD void alignment_pair(const double *Ri, const double *Rj, const double *Ti, const double *Tj, const double *dp,
                             const double *dv, double dt, const double *tic, const double *gcols, int k, const double *gfixed,
                             double *tA, double *tb) {
    for (int i = 0; i < 60; i++) tA[i] = 0;
    double RiT_Rj[9], v[3], w[3];
    mat3_tmul(Ri, Rj, RiT_Rj);
    for (int r = 0; r < 3; r++) {
        tA[10 * r + r] = -dt;
        tA[10 * (3 + r) + r] = -1.0;
        for (int c = 0; c < 3; c++) tA[10 * (3 + r) + 3 + c] = RiT_Rj[3 * r + c];
        for (int c = 0; c < k; c++) {
            double s = 0;  // (Ri^T gcols)(r, c)
            for (int m = 0; m < 3; m++) s += Ri[3 * m + r] * gcols[k * m + c];
            tA[10 * r + 6 + c] = s * dt * dt / 2;
            tA[10 * (3 + r) + 6 + c] = s * dt;
        }
    }
    const double dT[3] = {Tj[0] - Ti[0], Tj[1] - Ti[1], Tj[2] - Ti[2]};
    mat3_tvec(Ri, dT, v);
    for (int r = 0; r < 3; r++) tA[10 * r + 6 + k] = v[r] / 100.0;
    mat3_vec(RiT_Rj, tic, v);
    mat3_tvec(Ri, gfixed, w);
    for (int r = 0; r < 3; r++) {
        tb[r] = dp[r] + v[r] - tic[r] - w[r] * dt * dt / 2;
        tb[3 + r] = dv[r] - w[r] * dt;
    }
}

// tmp_A^T tmp_A and tmp_A^T tmp_b contribution of one IMU interval to solveGyroscopeBias: A (9) += J^T J, b (3) += J^T r
LFVIO_HD void gyro_bias_pair(const double *Ri, const double *Rj, const double *dq, const double *Jrg, double *A, double *b) {
    double Rij[9], qij[4], qi[4], q[4];
    mat3_tmul(Ri, Rj, Rij);
    rot_to_quat(Rij, qij);
    const double n2 = dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2] + dq[3] * dq[3];
    qi[0] = dq[0] / n2, qi[1] = -dq[1] / n2, qi[2] = -dq[2] / n2, qi[3] = -dq[3] / n2;
    quat_mul(qi, qij, q);
    const double r[3] = {2 * q[1], 2 * q[2], 2 * q[3]};
    mat3_tmul(Jrg, Jrg, A);
    mat3_tvec(Jrg, r, b);
}

// ------------------------------------------------------------------------------------------------------------
// extrinsic rotation calibration (InitialEXRotation::CalibrationExRotation, initial/initial_ex_rotation.cpp:13-68)
// ------------------------------------------------------------------------------------------------------------
// Rc / Rimu / Rc_g: N x 9 row-major rotation matrices of the pairs i = 1 .. N; A: 4N x 4 work space.  ric: the estimated
// camera-to-IMU rotation (row-major), sv: the four singular values of the stacked system, descending.
LFVIO_HD void ex_rotation_solve(int N, const double *Rc, const double *Rimu, const double *Rc_g, double *A, double *ric, double *sv) {
    for (int i = 0; i < N; i++) {
        double r1[4], r2[4], ri[4], d[4];
        rot_to_quat(Rc + 9 * i, r1);
        rot_to_quat(Rc_g + 9 * i, r2);
        rot_to_quat(Rimu + 9 * i, ri);
        const double r2c[4] = {r2[0], -r2[1], -r2[2], -r2[3]};
        quat_mul(r1, r2c, d);  // Eigen::Quaternion::angularDistance
        const double ang = 180.0 / 3.14159265358979323846 * 2.0 * atan2(sqrt(d[1] * d[1] + d[2] * d[2] + d[3] * d[3]), fabs(d[0]));
        const double huber = ang > 5.0 ? 5.0 / ang : 1.0;
        // L (left-multiplication matrix of Rc, :35-41) - R (right-multiplication matrix of Rimu, :43-50), coefficients (x, y, z, w)
        const double lw = r1[0], lx = r1[1], ly = r1[2], lz = r1[3], rw = ri[0], rx = ri[1], ry = ri[2], rz = ri[3];
        const double Lm[16] = {lw, -lz, ly, lx, lz, lw, -lx, ly, -ly, lx, lw, lz, -lx, -ly, -lz, lw};
        const double Rm[16] = {rw, rz, -ry, rx, -rz, rw, rx, ry, ry, -rx, rw, rz, -rx, -ry, -rz, rw};
        for (int k = 0; k < 16; k++) A[16 * i + k] = huber * (Lm[k] - Rm[k]);
    }
    double V[16];
    onesided_jacobi(4 * N, 4, A, V);
    double nn[4];
    int ord[4] = {0, 1, 2, 3};
    for (int c = 0; c < 4; c++) {
        double t = 0;
        for (int r = 0; r < 4 * N; r++) t += A[4 * r + c] * A[4 * r + c];
        nn[c] = sqrt(t);
    }
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3 - a; b++)
            if (nn[ord[b]] < nn[ord[b + 1]]) {
                const int t = ord[b];
                ord[b] = ord[b + 1];
                ord[b + 1] = t;
            }
    for (int c = 0; c < 4; c++) sv[c] = nn[ord[c]];
    const int o = ord[3];  // svd.matrixV().col(3); Quaterniond(Vector4d) reads (x, y, z, w)
    const double x[4] = {V[o], V[4 + o], V[8 + o], V[12 + o]};
    const double n = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
    const double q[4] = {x[3] / n, x[0] / n, x[1] / n, x[2] / n};
    double R[9];
    quat_to_rot(q, R);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) ric[3 * r + c] = R[3 * c + r];  // toRotationMatrix().inverse()
}

// ------------------------------------------------------------------------------------------------------------
// bundle adjustment residual
// ------------------------------------------------------------------------------------------------------------
// r = n(R(q) p + t) - obs; Jr (rotation, local: ceres::QuaternionParameterization, x_plus = [cos|d|, sin|d| d/|d|] (x) x),
// Jt (translation), Jp (point): each 3 x 3 row-major, may be skipped with want_jac = false.
LFVIO_HD void ba_residual_jac(const double *q, const double *t, const double *p, const double *obs, bool want_jac, double *r,
                              double *Jr, double *Jt, double *Jp) {
    const double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double qq[4] = {q[0] / qn, q[1] / qn, q[2] / qn, q[3] / qn};
    double R[9], rp[3];
    quat_to_rot(qq, R);
    mat3_vec(R, p, rp);
    const double pc[3] = {rp[0] + t[0], rp[1] + t[1], rp[2] + t[2]};
    const double n2 = pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2], nrm = sqrt(n2);
    for (int k = 0; k < 3; k++) r[k] = pc[k] / nrm - obs[k];
    if (!want_jac) return;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Jt[3 * i + j] = ((i == j ? 1.0 : 0.0) - pc[i] * pc[j] / n2) / nrm;
    const double S[9] = {0, 2 * rp[2], -2 * rp[1], -2 * rp[2], 0, 2 * rp[0], 2 * rp[1], -2 * rp[0], 0};  // -2 [rp]x
    mat3_mul(Jt, S, Jr);
    mat3_mul(Jt, R, Jp);
}
LFVIO_HD void quat_plus(const double *q, const double *d, double *o) {  // QuaternionParameterization::Plus
    const double nd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (nd == 0.0) {
        for (int k = 0; k < 4; k++) o[k] = q[k];
        return;
    }
    const double s = sin(nd) / nd;
    const double dq[4] = {cos(nd), s * d[0], s * d[1], s * d[2]};
    quat_mul(dq, q, o);
}

}  // namespace im
}  // namespace lfvio
