// TEST INFRASTRUCTURE ONLY.  Compiles lf-vio_b200/csrc/init_math.cuh (the __host__ __device__ cores the initialisation
// kernels call) with g++ and wraps them with serial loops where the kernels use parallel reductions, so that
// tests/test_init_math_host.py can check the arithmetic against oracle/initial.py without a GPU.  Never linked into
// liblfvio.so or liblfvio_host.so; the product path runs these cores on the device only.
#include "../../lf-vio_b200/csrc/ba_core.cuh"
#include <vector>

using namespace lfvio::im;

extern "C" {

void shim_triangulate_point(const double *P0, const double *P1, const double *p0, const double *p1, double *X) {
    triangulate_point(P0, P1, p0, p1, X);
}

void shim_recover_pose(const double *E, const double *b1, const double *b2, int n, double *R, double *t, int *counts) {
    double R1[9], R2[9], tt[3], nt[3];
    decompose_essential(E, R1, R2, tt);
    for (int k = 0; k < 3; k++) nt[k] = -tt[k];
    const double *Rs[4] = {R1, R2, R1, R2};
    const double *ts[4] = {tt, tt, nt, nt};
    for (int c = 0; c < 4; c++) {
        counts[c] = 0;
        for (int i = 0; i < n; i++) counts[c] += cheirality(Rs[c], ts[c], b1 + 3 * i, b2 + 3 * i);
    }
    const int g1 = counts[0], g2 = counts[1], g3 = counts[2], g4 = counts[3];
    int k;
    if (g1 >= g2 && g1 >= g3 && g1 >= g4) k = 0;
    else if (g2 >= g1 && g2 >= g3 && g2 >= g4) k = 1;
    else if (g3 >= g1 && g3 >= g2 && g3 >= g4) k = 2;
    else k = 3;
    for (int i = 0; i < 9; i++) R[i] = Rs[k][i];
    for (int i = 0; i < 3; i++) t[i] = ts[k][i];
}

double shim_epnp(const double *pws, const double *us, int n, double *R, double *T) {
    double mean[3] = {0, 0, 0}, cov[9] = {0}, cws[12], CCi[9];
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) mean[k] += pws[3 * i + k];
    for (int k = 0; k < 3; k++) mean[k] /= n;
    for (int i = 0; i < n; i++)
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) cov[3 * a + b] += (pws[3 * i + a] - mean[a]) * (pws[3 * i + b] - mean[b]);
    epnp_control_points(n, mean, cov, cws);
    epnp_cc_inv(cws, CCi);
    double *al = new double[4 * n];
    double MtM[144] = {0}, V[144];
    for (int i = 0; i < n; i++) {
        epnp_alphas(cws, CCi, pws + 3 * i, al + 4 * i);
        for (int r = 0; r < 12; r++)
            for (int c = 0; c < 12; c++) MtM[12 * r + c] += epnp_mtm_entry(al + 4 * i, us + 3 * i, r, c);
    }
    EpnpCore core;
    epnp_core(MtM, cws, &core, V);
    double best = 0;
    for (int which = 0; which < 3; which++) {
        double ccs[12], pc0[3] = {0, 0, 0}, W[9] = {0}, Rw[9], Tw[3];
        epnp_ccs(&core, which, ccs);
        double sgn = 1.0;
        {
            double z = 0;
            for (int j = 0; j < 4; j++) z += al[j] * ccs[3 * j + 2];
            const double s0 = us[2] > 0 ? 1.0 : -1.0;
            if ((z < 0.0 && s0 > 0) || (z > 0.0 && s0 < 0)) sgn = -1.0;
        }
        double *pcs = new double[3 * n];
        for (int i = 0; i < n; i++)
            for (int k = 0; k < 3; k++) {
                double s = 0;
                for (int j = 0; j < 4; j++) s += al[4 * i + j] * ccs[3 * j + k];
                pcs[3 * i + k] = sgn * s;
                pc0[k] += sgn * s;
            }
        for (int k = 0; k < 3; k++) pc0[k] /= n;
        for (int i = 0; i < n; i++)
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) W[3 * a + b] += (pcs[3 * i + a] - pc0[a]) * (pws[3 * i + b] - mean[b]);
        epnp_rt_from_w(W, pc0, mean, Rw, Tw);
        double err = 0;
        for (int i = 0; i < n; i++) {
            double p[3];
            mat3_vec(Rw, pws + 3 * i, p);
            for (int k = 0; k < 3; k++) {
                const double d = us[3 * i + k] - (p[k] + Tw[k]);
                err += d * d;
            }
        }
        err /= n;
        delete[] pcs;
        if (which == 0 || err < best) {
            best = err;
            for (int i = 0; i < 9; i++) R[i] = Rw[i];
            for (int i = 0; i < 3; i++) T[i] = Tw[i];
        }
    }
    delete[] al;
    return best;
}

// A, b of LinearAlignment (k = 3) / one RefineGravity pass (k = 2): ns = 3 n + k + 1, A row-major ns x ns (x 1000)
void shim_alignment_system(const double *Rs, const double *Ts, const double *dp, const double *dv, const double *dts, int n,
                           const double *tic, const double *gcols, int k, const double *gfixed, double *A, double *b) {
    const int ns = 3 * n + k + 1, kk = k + 1;
    for (int i = 0; i < ns * ns; i++) A[i] = 0;
    for (int i = 0; i < ns; i++) b[i] = 0;
    for (int i = 0; i < n - 1; i++) {
        double tA[60], tb[6];
        alignment_pair(Rs + 9 * i, Rs + 9 * (i + 1), Ts + 3 * i, Ts + 3 * (i + 1), dp + 3 * i, dv + 3 * i, dts[i], tic, gcols, k, gfixed,
                       tA, tb);
        const int w = 6 + kk;
        for (int a = 0; a < w; a++) {
            const int ga = a < 6 ? 3 * i + a : ns - kk + (a - 6);
            double rb = 0;
            for (int r = 0; r < 6; r++) rb += tA[10 * r + a] * tb[r];
            b[ga] += rb * 1000.0;
            for (int c = 0; c < w; c++) {
                const int gc = c < 6 ? 3 * i + c : ns - kk + (c - 6);
                double s = 0;
                for (int r = 0; r < 6; r++) s += tA[10 * r + a] * tA[10 * r + c];
                A[ga * ns + gc] += s * 1000.0;
            }
        }
    }
}

void shim_tangent_basis(const double *g0, double *bc) { tangent_basis(g0, bc); }

void shim_gyro_bias(const double *Rs, const double *dq, const double *Jrg, int n, double *A, double *b) {
    for (int i = 0; i < 9; i++) A[i] = 0;
    for (int i = 0; i < 3; i++) b[i] = 0;
    for (int i = 0; i < n - 1; i++) {
        double a[9], bb[3];
        gyro_bias_pair(Rs + 9 * i, Rs + 9 * (i + 1), dq + 4 * i, Jrg + 9 * i, a, bb);
        for (int k = 0; k < 9; k++) A[k] += a[k];
        for (int k = 0; k < 3; k++) b[k] += bb[k];
    }
}

void shim_ba_residual_jac(const double *q, const double *t, const double *p, const double *obs, double *r, double *Jr, double *Jt,
                          double *Jp) {
    ba_residual_jac(q, t, p, obs, true, r, Jr, Jt, Jp);
}
void shim_quat_plus(const double *q, const double *d, double *o) { quat_plus(q, d, o); }
void shim_svd3(const double *M, double *U, double *s, double *V) { svd3(M, U, s, V); }
void shim_ex_rotation_solve(int N, const double *Rc, const double *Rimu, const double *Rc_g, double *ric, double *sv) {
    std::vector<double> A(16 * (size_t)N);
    ex_rotation_solve(N, Rc, Rimu, Rc_g, A.data(), ric, sv);
}

// The LM loop of sfm_ba_kernel, serial: same per-point work items and decisions (ba_core.cuh), plain += instead of atomics
// and a serial Cholesky.  Returns the number of iterations.
int shim_ba(int F, int L, int l, const int *obs_first, const int *obs_frame, const double *obs_pts, const unsigned char *state,
            double *cq, double *ct, double *pts, double *costs, int *n_success, int *converged) {
    const BaProblem P{F, L, l, obs_first, obs_frame, obs_pts, state};
    const int D = 6 * F;
    std::vector<double> S(D * D), gc(D), scale_c(D, 1.0), dxc(D), V(9 * L), gp(3 * L), sp(3 * L, 1.0), Vinv(9 * L), cand(3 * L),
        cand_q(4 * F), cand_t(3 * F);
    BaCtl c;
    ba_init(&c);
    auto add = [](double *p, double v) { *p += v; };
    bool first = true;
    while (true) {
        std::fill(S.begin(), S.end(), 0.0);
        std::fill(gc.begin(), gc.end(), 0.0);
        c.cost = 0, c.gmax = 0;
        for (int j = 0; j < L; j++)
            if (state[j]) c.cost += ba_point_linearize(P, cq, ct, pts, j, S.data(), gc.data(), &V[9 * j], &gp[3 * j], &c.gmax, add);
        for (int d = 0; d < D; d++) c.gmax = fmax(c.gmax, fabs(gc[d]));
        if (!ba_begin_iteration(&c)) break;
        if (first) {
            for (int d = 0; d < D; d++) scale_c[d] = ba_jacobi_scale(S[d * D + d]);
            for (int j = 0; j < L; j++)
                if (state[j])
                    for (int a = 0; a < 3; a++) sp[3 * j + a] = ba_jacobi_scale(V[9 * j + 4 * a]);
            first = false;
        }
        for (int a = 0; a < D; a++) {
            for (int b = 0; b < D; b++) S[a * D + b] *= scale_c[a] * scale_c[b];
            gc[a] *= scale_c[a];
        }
        for (int d = 0; d < D; d++) S[d * D + d] = S[d * D + d] == 0.0 ? 1.0 : S[d * D + d] + ba_lm_diag(S[d * D + d], c.radius);
        for (int j = 0; j < L; j++)
            if (state[j])
                ba_point_schur(P, cq, ct, pts, j, &V[9 * j], &gp[3 * j], scale_c.data(), &sp[3 * j], c.radius, S.data(), gc.data(),
                               &Vinv[9 * j], add);
        // Cholesky (lower) and solve S x = gc, dxc = -x
        c.chol_ok = 1;
        for (int j = 0; j < D && c.chol_ok; j++) {
            double d = S[j * D + j];
            for (int k = 0; k < j; k++) d -= S[j * D + k] * S[j * D + k];
            if (!(d > 0.0)) {
                c.chol_ok = 0;
                break;
            }
            const double lj = sqrt(d);
            S[j * D + j] = lj;
            for (int i = j + 1; i < D; i++) {
                double v = S[i * D + j];
                for (int k = 0; k < j; k++) v -= S[i * D + k] * S[j * D + k];
                S[i * D + j] = v / lj;
            }
        }
        c.model_change = c.step2 = c.xnorm2 = 0;
        if (c.chol_ok) {
            for (int i = 0; i < D; i++) {
                double v = gc[i];
                for (int k = 0; k < i; k++) v -= S[i * D + k] * dxc[k];
                dxc[i] = v / S[i * D + i];
            }
            for (int i = D - 1; i >= 0; i--) {
                double v = dxc[i];
                for (int k = i + 1; k < D; k++) v -= S[k * D + i] * dxc[k];
                dxc[i] = v / S[i * D + i];
            }
            for (int i = 0; i < D; i++) dxc[i] = -dxc[i];
            for (int j = 0; j < L; j++)
                if (state[j])
                    ba_point_backsub(P, cq, ct, pts, j, &gp[3 * j], scale_c.data(), &sp[3 * j], &Vinv[9 * j], dxc.data(), cand.data(),
                                     &c.model_change, &c.step2, &c.xnorm2);
            for (int f = 0; f < F; f++)
                ba_cam_candidate(P, cq, ct, f, scale_c.data(), dxc.data(), cand_q.data(), cand_t.data(), &c.step2, &c.xnorm2);
            c.cand_cost = 0;
            for (int j = 0; j < L; j++)
                if (state[j]) c.cand_cost += ba_point_cost(P, cand_q.data(), cand_t.data(), cand.data(), j);
        }
        if (ba_decide(&c)) {
            for (int j = 0; j < L; j++)
                if (state[j])
                    for (int a = 0; a < 3; a++) pts[3 * j + a] = cand[3 * j + a];
            for (int k = 0; k < 4 * F; k++) cq[k] = cand_q[k];
            for (int k = 0; k < 3 * F; k++) ct[k] = cand_t[k];
        }
        if (c.done) break;
    }
    for (int k = 0; k <= c.n_success; k++) costs[k] = c.costs[k];
    *n_success = c.n_success;
    *converged = c.converged;
    return c.iter;
}
}
