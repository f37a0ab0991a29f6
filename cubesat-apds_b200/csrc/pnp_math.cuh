// Scalar f64 building blocks of the PnP stage (replaces cv::solvePnPRansac as called at
// homographier/src/homographier/mod.rs:347-361): OpenCV's one-sided Jacobi SVD (the SIGNS of the
// singular vectors of PW0^T PW0 change EPnP's answer under noise, so the rotation order is kept),
// cv::solve(DECOMP_SVD), cv::Rodrigues in both directions, and EPnP itself.
//
// Everything here is `__host__ __device__` so the same source is exercised on the host by
// tests/hostcheck (a test-only harness; the product never runs it on the CPU) and on the device
// by ransac_pnp.cu.  EPnP is written against an "executor" that owns the loop over the used
// points: SerialExec (one thread = one minimal sample) or the CTA-wide BlockExec of ransac_pnp.cu
// (every thread runs the scalar part redundantly, the sums over points are block reductions).
#pragma once
#include <cfloat>
#include <cmath>

#ifdef __CUDACC__
#define DUNK_HD __host__ __device__
#else
#define DUNK_HD
#endif
// phase stamps exist only in the `make timing` build (ctx.h); the host check compiles this header alone
#ifndef DUNK_PHASE
#define DUNK_PHASE(k) \
    do {              \
    } while (0)
#endif

namespace dunk {
namespace pnp {

struct Camera {
    double fu, fv, uc, vc;
};

struct Point {
    double X, Y, Z, u, v;
};

// cv::undistortPoints with zero distortion ((u - cx) * (1/fx)), optionally rounded to f32 (f32 image
// points give f32 normalised points), then epnp::init_points maps back: x * fu + uc.
DUNK_HD inline Point load_point(const float* obj, const float* img, int i, const Camera& c, bool f32_normalised) {
    Point p;
    p.X = obj[3 * i];
    p.Y = obj[3 * i + 1];
    p.Z = obj[3 * i + 2];
    double xn = ((double)img[2 * i] - c.uc) * (1.0 / c.fu);
    double yn = ((double)img[2 * i + 1] - c.vc) * (1.0 / c.fv);
    if (f32_normalised) {
        xn = (double)(float)xn;
        yn = (double)(float)yn;
    }
    p.u = xn * c.fu + c.uc;
    p.v = yn * c.fv + c.vc;
    return p;
}

// JacobiSVDImpl_ (core/src/lapack.cpp).  At: N rows of length M = the columns of A.  On return the
// rows of At are the left singular vectors (unit length), W the singular values in decreasing
// order, Vt (N x N, only if WantV) the right singular vectors as rows.
template <int M, int N, bool WantV>
DUNK_HD void jacobi_svd(double* At, double* W, double* Vt) {
    const double eps = DBL_EPSILON * 10;
    for (int i = 0; i < N; ++i) {
        double sd = 0;
        for (int k = 0; k < M; ++k) sd += At[i * M + k] * At[i * M + k];
        W[i] = sd;
        if (WantV) {
            for (int k = 0; k < N; ++k) Vt[i * N + k] = 0;
            Vt[i * N + i] = 1;
        }
    }
    const int max_iter = M > 30 ? M : 30;
    for (int iter = 0; iter < max_iter; ++iter) {
        bool changed = false;
        for (int i = 0; i < N - 1; ++i)
            for (int j = i + 1; j < N; ++j) {
                double* Ai = At + i * M;
                double* Aj = At + j * M;
                double a = W[i], p = 0, b = W[j];
                for (int k = 0; k < M; ++k) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = hypot(p, beta);
                double c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (int k = 0; k < M; ++k) {
                    const double t0 = c * Ai[k] + s * Aj[k];
                    const double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0;
                    Aj[k] = t1;
                    a += t0 * t0;
                    b += t1 * t1;
                }
                W[i] = a;
                W[j] = b;
                changed = true;
                if (WantV) {
                    double* Vi = Vt + i * N;
                    double* Vj = Vt + j * N;
                    for (int k = 0; k < N; ++k) {
                        const double t0 = c * Vi[k] + s * Vj[k];
                        const double t1 = -s * Vi[k] + c * Vj[k];
                        Vi[k] = t0;
                        Vj[k] = t1;
                    }
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < N; ++i) {
        double sd = 0;
        for (int k = 0; k < M; ++k) sd += At[i * M + k] * At[i * M + k];
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < N - 1; ++i) {
        int j = i;
        for (int k = i + 1; k < N; ++k)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            double t = W[i]; W[i] = W[j]; W[j] = t;
            for (int k = 0; k < M; ++k) { t = At[i * M + k]; At[i * M + k] = At[j * M + k]; At[j * M + k] = t; }
            if (WantV)
                for (int k = 0; k < N; ++k) { t = Vt[i * N + k]; Vt[i * N + k] = Vt[j * N + k]; Vt[j * N + k] = t; }
        }
    }
    for (int i = 0; i < N; ++i) {
        const double s = W[i] > DBL_MIN ? 1.0 / W[i] : 0.0;   // (an exactly-zero singular value gets a random vector in OpenCV)
        for (int k = 0; k < M; ++k) At[i * M + k] *= s;
    }
}

// cv::solve(A, b, x, DECOMP_SVD) for an M x N system (row-major A): least squares through the SVD
template <int M, int N>
DUNK_HD void svd_solve(const double* A, const double* b, double* x) {
    double At[N * M], W[N], Vt[N * N];
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) At[j * M + i] = A[i * N + j];
    jacobi_svd<M, N, true>(At, W, Vt);
    double thr = 0;
    for (int j = 0; j < N; ++j) thr += W[j];
    thr *= DBL_EPSILON * 2;
    for (int j = 0; j < N; ++j) x[j] = 0;
    for (int j = 0; j < N; ++j) {
        if (!(W[j] > thr)) continue;
        double y = 0;
        for (int i = 0; i < M; ++i) y += At[j * M + i] * b[i];
        y /= W[j];
        for (int k = 0; k < N; ++k) x[k] += y * Vt[j * N + k];
    }
}

// cvInvert(A, Ai, CV_SVD) for a 3 x 3 matrix
DUNK_HD inline void svd_inverse3(const double* A, double* Ai) {
    double At[9], W[3], Vt[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) At[j * 3 + i] = A[i * 3 + j];
    jacobi_svd<3, 3, true>(At, W, Vt);
    const double thr = (W[0] + W[1] + W[2]) * DBL_EPSILON * 2;
    for (int i = 0; i < 9; ++i) Ai[i] = 0;
    for (int k = 0; k < 3; ++k) {
        if (!(W[k] > thr)) continue;
        const double wi = 1.0 / W[k];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Ai[r * 3 + c] += Vt[k * 3 + r] * wi * At[k * 3 + c];
    }
}

// R = U * Vt of the SVD of a 3 x 3 matrix A (row-major)
DUNK_HD inline void svd_rotation3(const double* A, double* R) {
    double At[9], W[3], Vt[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) At[j * 3 + i] = A[i * 3 + j];
    jacobi_svd<3, 3, true>(At, W, Vt);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R[r * 3 + c] = At[0 * 3 + r] * Vt[0 * 3 + c] + At[1 * 3 + r] * Vt[1 * 3 + c] + At[2 * 3 + r] * Vt[2 * 3 + c];
}

// Householder least squares of a 6 x 4 system (epnp::qr_solve); A and b are destroyed
DUNK_HD inline bool householder_ls_6x4(double* A, double* b, double* x) {
    const int M = 6, N = 4;
    for (int k = 0; k < N; ++k) {
        double nrm = 0;
        for (int i = k; i < M; ++i) nrm += A[i * N + k] * A[i * N + k];
        nrm = sqrt(nrm);
        if (nrm == 0) return false;
        const double alpha = A[k * N + k] > 0 ? -nrm : nrm;
        double v[6];
        for (int i = k; i < M; ++i) v[i] = A[i * N + k];
        v[k] -= alpha;
        double vn = 0;
        for (int i = k; i < M; ++i) vn += v[i] * v[i];
        if (vn == 0) continue;
        const double f = 2.0 / vn;
        for (int j = k; j < N; ++j) {
            double d = 0;
            for (int i = k; i < M; ++i) d += v[i] * A[i * N + j];
            d *= f;
            for (int i = k; i < M; ++i) A[i * N + j] -= d * v[i];
        }
        double d = 0;
        for (int i = k; i < M; ++i) d += v[i] * b[i];
        d *= f;
        for (int i = k; i < M; ++i) b[i] -= d * v[i];
    }
    for (int k = N - 1; k >= 0; --k) {
        double s = b[k];
        for (int j = k + 1; j < N; ++j) s -= A[k * N + j] * x[j];
        x[k] = s / A[k * N + k];
    }
    return true;
}

// cv::Rodrigues, vector -> matrix
DUNK_HD inline void rodrigues_to_matrix(const double* r, double* R) {
    const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < DBL_EPSILON) {
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0);
        return;
    }
    const double c = cos(theta), s = sin(theta), c1 = 1.0 - c, it = 1.0 / theta;
    const double x = r[0] * it, y = r[1] * it, z = r[2] * it;
    R[0] = c + c1 * x * x;      R[1] = c1 * x * y - s * z;  R[2] = c1 * x * z + s * y;
    R[3] = c1 * x * y + s * z;  R[4] = c + c1 * y * y;      R[5] = c1 * y * z - s * x;
    R[6] = c1 * x * z - s * y;  R[7] = c1 * y * z + s * x;  R[8] = c + c1 * z * z;
}

// cv::Rodrigues, matrix -> vector (the matrix is first projected onto SO(3) through its SVD)
DUNK_HD inline void rodrigues_to_vector(const double* Rin, double* r) {
    double R[9];
    svd_rotation3(Rin, R);
    r[0] = R[7] - R[5];
    r[1] = R[2] - R[6];
    r[2] = R[3] - R[1];
    const double s = sqrt((r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) {
            r[0] = r[1] = r[2] = 0;
            return;
        }
        double t = (R[0] + 1) * 0.5;
        r[0] = sqrt(t > 0 ? t : 0.);
        t = (R[4] + 1) * 0.5;
        r[1] = sqrt(t > 0 ? t : 0.) * (R[1] < 0 ? -1. : 1.);
        t = (R[8] + 1) * 0.5;
        r[2] = sqrt(t > 0 ? t : 0.) * (R[2] < 0 ? -1. : 1.);
        if (fabs(r[0]) < fabs(r[1]) && fabs(r[0]) < fabs(r[2]) && ((R[5] > 0) != (r[1] * r[2] > 0))) r[2] = -r[2];
        theta /= sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        r[0] *= theta; r[1] *= theta; r[2] *= theta;
        return;
    }
    const double vth = theta / (2 * s);
    r[0] *= vth; r[1] *= vth; r[2] *= vth;
}

// One thread owns the whole point set: `idx` (may be null = identity) selects `n` points.
struct SerialExec {
    static constexpr bool kBlock = false;
    const float* obj;
    const float* img;
    const int* idx;
    int n;
    Camera cam;
    bool f32_normalised;
    DUNK_HD double count() const { return (double)n; }
    DUNK_HD Point first() const { return load_point(obj, img, idx ? idx[0] : 0, cam, f32_normalised); }
    template <int K, class F>
    DUNK_HD void sum(F f, double (&out)[K]) const {
        for (int k = 0; k < K; ++k) out[k] = 0;
        for (int i = 0; i < n; ++i) f(load_point(obj, img, idx ? idx[i] : i, cam, f32_normalised), out);
    }
};

// epnp::compute_pose (calib3d/src/epnp.cpp).  Returns the mean reprojection error of the chosen
// candidate; R (row-major) and t receive the pose.
template <class Exec>
DUNK_HD double epnp_solve(const Exec& ex, const Camera& cam, double (&R)[9], double (&t)[3]) {
    const double n = ex.count();
    const double fu = cam.fu, fv = cam.fv, uc = cam.uc, vc = cam.vc;
    // ---- choose_control_points
    double cws[4][3];
    {
        double s[3];
        ex.template sum<3>([](const Point& p, double (&a)[3]) { a[0] += p.X; a[1] += p.Y; a[2] += p.Z; }, s);
        for (int j = 0; j < 3; ++j) cws[0][j] = s[j] / n;
    }
    const double c0x = cws[0][0], c0y = cws[0][1], c0z = cws[0][2];
    {
        double s[6];
        ex.template sum<6>(
            [=](const Point& p, double (&a)[6]) {
                const double x = p.X - c0x, y = p.Y - c0y, z = p.Z - c0z;
                a[0] += x * x; a[1] += x * y; a[2] += x * z; a[3] += y * y; a[4] += y * z; a[5] += z * z;
            },
            s);
        double At[9] = {s[0], s[1], s[2], s[1], s[3], s[4], s[2], s[4], s[5]}, dc[3];
        jacobi_svd<3, 3, false>(At, dc, nullptr);
        for (int i = 1; i < 4; ++i) {
            const double k = sqrt(dc[i - 1] / n);
            for (int j = 0; j < 3; ++j) cws[i][j] = cws[0][j] + k * At[3 * (i - 1) + j];
        }
    }
    if (Exec::kBlock) DUNK_PHASE(6);
    // ---- compute_barycentric_coordinates: alphas[1..3] = CC^-1 (pw - cw0)
    double ci[9];
    {
        double cc[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 1; j < 4; ++j) cc[3 * i + j - 1] = cws[j][i] - cws[0][i];
        svd_inverse3(cc, ci);
    }
    auto alphas = [=](const Point& p, double (&a)[4]) {
        const double x = p.X - c0x, y = p.Y - c0y, z = p.Z - c0z;
        a[1] = ci[0] * x + ci[1] * y + ci[2] * z;
        a[2] = ci[3] * x + ci[4] * y + ci[5] * z;
        a[3] = ci[6] * x + ci[7] * y + ci[8] * z;
        a[0] = 1.0 - a[1] - a[2] - a[3];
    };
    // ---- fill_M, M^T M (upper triangle, 78 sums), its 4 smallest singular vectors
    double ut[144];
    {
        double s[78];
        ex.template sum<78>(
            [=](const Point& p, double (&acc)[78]) {
                double a[4];
                alphas(p, a);
                double m1[12], m2[12];
                for (int j = 0; j < 4; ++j) {
                    m1[3 * j] = a[j] * fu; m1[3 * j + 1] = 0.0;       m1[3 * j + 2] = a[j] * (uc - p.u);
                    m2[3 * j] = 0.0;       m2[3 * j + 1] = a[j] * fv; m2[3 * j + 2] = a[j] * (vc - p.v);
                }
                int k = 0;
                for (int r = 0; r < 12; ++r)
                    for (int c = r; c < 12; ++c) acc[k++] += m1[r] * m1[c] + m2[r] * m2[c];
            },
            s);
        int k = 0;
        for (int r = 0; r < 12; ++r)
            for (int c = r; c < 12; ++c) {
                ut[r * 12 + c] = s[k];
                ut[c * 12 + r] = s[k];
                ++k;
            }
        double d[12];
        if (Exec::kBlock) DUNK_PHASE(7);
        jacobi_svd<12, 12, false>(ut, d, nullptr);
        if (Exec::kBlock) DUNK_PHASE(8);
    }
    const double* v[4] = {ut + 12 * 11, ut + 12 * 10, ut + 12 * 9, ut + 12 * 8};
    // ---- compute_L_6x10, compute_rho
    const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
    double L[60], rho[6];
    for (int p = 0; p < 6; ++p) {
        double dv[4][3];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 3; ++j) dv[i][j] = v[i][3 * pa[p] + j] - v[i][3 * pb[p] + j];
        auto dot = [&](int a, int b) { return dv[a][0] * dv[b][0] + dv[a][1] * dv[b][1] + dv[a][2] * dv[b][2]; };
        double* row = L + 10 * p;
        row[0] = dot(0, 0); row[1] = 2.0 * dot(0, 1); row[2] = dot(1, 1); row[3] = 2.0 * dot(0, 2); row[4] = 2.0 * dot(1, 2);
        row[5] = dot(2, 2); row[6] = 2.0 * dot(0, 3); row[7] = 2.0 * dot(1, 3); row[8] = 2.0 * dot(2, 3); row[9] = dot(3, 3);
        double d2 = 0;
        for (int j = 0; j < 3; ++j) d2 += (cws[pa[p]][j] - cws[pb[p]][j]) * (cws[pa[p]][j] - cws[pb[p]][j]);
        rho[p] = d2;
    }
    if (Exec::kBlock) DUNK_PHASE(9);
    // ---- betas: three approximations, each refined by 5 Gauss-Newton steps
    double betas[3][4];
    {   // find_betas_approx_1: columns 0 1 3 6 (B11 B12 B13 B14)
        double A[24], b4[4];
        for (int i = 0; i < 6; ++i) { A[4 * i] = L[10 * i]; A[4 * i + 1] = L[10 * i + 1]; A[4 * i + 2] = L[10 * i + 3]; A[4 * i + 3] = L[10 * i + 6]; }
        svd_solve<6, 4>(A, rho, b4);
        double* be = betas[0];
        if (b4[0] < 0) { be[0] = sqrt(-b4[0]); be[1] = -b4[1] / be[0]; be[2] = -b4[2] / be[0]; be[3] = -b4[3] / be[0]; }
        else           { be[0] = sqrt(b4[0]);  be[1] = b4[1] / be[0];  be[2] = b4[2] / be[0];  be[3] = b4[3] / be[0]; }
    }
    {   // find_betas_approx_2: columns 0 1 2 (B11 B12 B22)
        double A[18], b3[3];
        for (int i = 0; i < 6; ++i) { A[3 * i] = L[10 * i]; A[3 * i + 1] = L[10 * i + 1]; A[3 * i + 2] = L[10 * i + 2]; }
        svd_solve<6, 3>(A, rho, b3);
        double* be = betas[1];
        if (b3[0] < 0) { be[0] = sqrt(-b3[0]); be[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0; }
        else           { be[0] = sqrt(b3[0]);  be[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0; }
        if (b3[1] < 0) be[0] = -be[0];
        be[2] = 0.0; be[3] = 0.0;
    }
    {   // find_betas_approx_3: columns 0..4 (B11 B12 B22 B13 B23)
        double A[30], b5[5];
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 5; ++j) A[5 * i + j] = L[10 * i + j];
        svd_solve<6, 5>(A, rho, b5);
        double* be = betas[2];
        if (b5[0] < 0) { be[0] = sqrt(-b5[0]); be[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0; }
        else           { be[0] = sqrt(b5[0]);  be[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0; }
        if (b5[1] < 0) be[0] = -be[0];
        be[2] = b5[3] / be[0];
        be[3] = 0.0;
    }
    if (Exec::kBlock) DUNK_PHASE(10);
    double ccs[3][4][3];   // candidate, control point, xyz
    for (int c = 0; c < 3; ++c) {
        double* be = betas[c];
        for (int it = 0; it < 5; ++it) {   // gauss_newton
            double A[24], b[6], x[4];
            for (int i = 0; i < 6; ++i) {
                const double* l = L + 10 * i;
                A[4 * i + 0] = 2 * l[0] * be[0] + l[1] * be[1] + l[3] * be[2] + l[6] * be[3];
                A[4 * i + 1] = l[1] * be[0] + 2 * l[2] * be[1] + l[4] * be[2] + l[7] * be[3];
                A[4 * i + 2] = l[3] * be[0] + l[4] * be[1] + 2 * l[5] * be[2] + l[8] * be[3];
                A[4 * i + 3] = l[6] * be[0] + l[7] * be[1] + l[8] * be[2] + 2 * l[9] * be[3];
                b[i] = rho[i] - (l[0] * be[0] * be[0] + l[1] * be[0] * be[1] + l[2] * be[1] * be[1] + l[3] * be[0] * be[2] +
                                 l[4] * be[1] * be[2] + l[5] * be[2] * be[2] + l[6] * be[0] * be[3] + l[7] * be[1] * be[3] +
                                 l[8] * be[2] * be[3] + l[9] * be[3] * be[3]);
            }
            if (!householder_ls_6x4(A, b, x)) break;
            for (int k = 0; k < 4; ++k) be[k] += x[k];
        }
        for (int j = 0; j < 4; ++j)        // compute_ccs
            for (int k = 0; k < 3; ++k)
                ccs[c][j][k] = be[0] * v[0][3 * j + k] + be[1] * v[1][3 * j + k] + be[2] * v[2][3 * j + k] + be[3] * v[3][3 * j + k];
    }
    if (Exec::kBlock) DUNK_PHASE(11);
    // ---- solve_for_sign: the first point must lie in front of the camera
    {
        double a[4];
        alphas(ex.first(), a);
        for (int c = 0; c < 3; ++c) {
            const double z = a[0] * ccs[c][0][2] + a[1] * ccs[c][1][2] + a[2] * ccs[c][2][2] + a[3] * ccs[c][3][2];
            if (z < 0.0)
                for (int j = 0; j < 4; ++j)
                    for (int k = 0; k < 3; ++k) ccs[c][j][k] = -ccs[c][j][k];
        }
    }
    // ---- estimate_R_and_t for the three candidates (Procrustes on camera/world point sets)
    auto pcs = [&](const double (&a)[4], int c, double (&pc)[3]) {
        for (int k = 0; k < 3; ++k) pc[k] = a[0] * ccs[c][0][k] + a[1] * ccs[c][1][k] + a[2] * ccs[c][2][k] + a[3] * ccs[c][3][k];
    };
    double pc0[3][3];
    {
        double s[9];
        ex.template sum<9>(
            [&](const Point& p, double (&acc)[9]) {
                double a[4], pc[3];
                alphas(p, a);
                for (int c = 0; c < 3; ++c) {
                    pcs(a, c, pc);
                    acc[3 * c] += pc[0]; acc[3 * c + 1] += pc[1]; acc[3 * c + 2] += pc[2];
                }
            },
            s);
        for (int c = 0; c < 3; ++c)
            for (int k = 0; k < 3; ++k) pc0[c][k] = s[3 * c + k] / n;
    }
    double Rs[3][9], ts[3][3];
    {
        double s[27];
        ex.template sum<27>(
            [&](const Point& p, double (&acc)[27]) {
                double a[4], pc[3];
                alphas(p, a);
                const double w[3] = {p.X - c0x, p.Y - c0y, p.Z - c0z};
                for (int c = 0; c < 3; ++c) {
                    pcs(a, c, pc);
                    for (int j = 0; j < 3; ++j)
                        for (int k = 0; k < 3; ++k) acc[9 * c + 3 * j + k] += (pc[j] - pc0[c][j]) * w[k];
                }
            },
            s);
        for (int c = 0; c < 3; ++c) {
            double* Rc = Rs[c];
            svd_rotation3(s + 9 * c, Rc);
            const double det = Rc[0] * Rc[4] * Rc[8] + Rc[1] * Rc[5] * Rc[6] + Rc[2] * Rc[3] * Rc[7] - Rc[2] * Rc[4] * Rc[6] -
                               Rc[1] * Rc[3] * Rc[8] - Rc[0] * Rc[5] * Rc[7];
            if (det < 0) { Rc[6] = -Rc[6]; Rc[7] = -Rc[7]; Rc[8] = -Rc[8]; }
            for (int k = 0; k < 3; ++k) ts[c][k] = pc0[c][k] - (Rc[3 * k] * c0x + Rc[3 * k + 1] * c0y + Rc[3 * k + 2] * c0z);
        }
    }
    if (Exec::kBlock) DUNK_PHASE(12);
    // ---- reprojection_error, pick the best candidate
    double err[3];
    ex.template sum<3>(
        [&](const Point& p, double (&acc)[3]) {
            for (int c = 0; c < 3; ++c) {
                const double* Rc = Rs[c];
                const double Xc = Rc[0] * p.X + Rc[1] * p.Y + Rc[2] * p.Z + ts[c][0];
                const double Yc = Rc[3] * p.X + Rc[4] * p.Y + Rc[5] * p.Z + ts[c][1];
                const double iz = 1.0 / (Rc[6] * p.X + Rc[7] * p.Y + Rc[8] * p.Z + ts[c][2]);
                const double ue = uc + fu * Xc * iz, ve = vc + fv * Yc * iz;
                acc[c] += sqrt((p.u - ue) * (p.u - ue) + (p.v - ve) * (p.v - ve));
            }
        },
        err);
    int N = 0;
    if (err[1] < err[0]) N = 1;
    if (err[2] < err[N]) N = 2;
    for (int i = 0; i < 9; ++i) R[i] = Rs[N][i];
    for (int i = 0; i < 3; ++i) t[i] = ts[N][i];
    return err[N] / n;
}

// ---- P3P minimal solver (SOLVEPNP_P3P, and OpenCV's kernel whenever exactly 4 points are given) ----
// Unknown depths s0, s1 = u s0, s2 = v s0 along the unit bearings of three image points; eliminating s0 and v
// from the three law-of-cosines equations leaves a quartic in u (coefficients derived symbolically,
// tools/derive_p3p.py).  Roots with OpenCV's Ferrari / Cardano routines (p3p.cpp solve_deg4 / solve_deg3),
// triad alignment of the camera-frame triangle with the world triangle (they are congruent), and the fourth
// point selects the solution with the smallest squared reprojection error in normalised coordinates.
DUNK_HD inline int solve_deg2(double a, double b, double c, double* x) {
    const double delta = b * b - 4 * a * c;
    if (delta < 0) return 0;
    const double inv_2a = 0.5 / a;
    if (delta == 0) { x[0] = -b * inv_2a; return 1; }
    const double s = sqrt(delta);
    x[0] = (-b + s) * inv_2a;
    x[1] = (-b - s) * inv_2a;
    return 2;
}

DUNK_HD inline int solve_deg3(double a, double b, double c, double d, double* x) {
    if (a == 0) {
        if (b == 0) {
            if (c == 0) return 0;
            x[0] = -d / c;
            return 1;
        }
        return solve_deg2(b, c, d, x);
    }
    const double inv_a = 1.0 / a;
    const double b_a = inv_a * b, c_a = inv_a * c, d_a = inv_a * d, b_a2 = b_a * b_a;
    const double Q = (3 * c_a - b_a2) / 9;
    const double R = (9 * b_a * c_a - 27 * d_a - 2 * b_a * b_a2) / 54;
    const double Q3 = Q * Q * Q, D = Q3 + R * R, b_a_3 = (1.0 / 3.0) * b_a;
    if (Q == 0) {
        if (R == 0) { x[0] = x[1] = x[2] = -b_a_3; return 3; }
        const double cr = pow(fabs(2 * R), 1.0 / 3.0);
        x[0] = (R < 0 ? -cr : cr) - b_a_3;
        return 1;
    }
    if (D <= 0) {
        double arg = R / sqrt(-Q3);
        arg = arg > 1. ? 1. : arg < -1. ? -1. : arg;
        const double theta = acos(arg), sq = sqrt(-Q);
        x[0] = 2 * sq * cos(theta / 3.0) - b_a_3;
        x[1] = 2 * sq * cos((theta + 2 * 3.14159265358979323846) / 3.0) - b_a_3;
        x[2] = 2 * sq * cos((theta + 4 * 3.14159265358979323846) / 3.0) - b_a_3;
        return 3;
    }
    double AD = 0, BD = 0;
    const double R_abs = fabs(R);
    if (R_abs > DBL_EPSILON) {
        AD = pow(R_abs + sqrt(D), 1.0 / 3.0);
        AD = R >= 0 ? AD : -AD;
        BD = -Q / AD;
    }
    x[0] = AD + BD - b_a_3;
    return 1;
}

DUNK_HD inline int solve_deg4(double a, double b, double c, double d, double e, double* x) {
    if (a == 0) return solve_deg3(b, c, d, e, x);
    const double inv_a = 1.0 / a;
    b *= inv_a; c *= inv_a; d *= inv_a; e *= inv_a;
    const double b2 = b * b, bc = b * c, b3 = b2 * b;
    double r[3];
    if (solve_deg3(1, -c, d * b - 4 * e, 4 * c * e - d * d - b2 * e, r) == 0) return 0;
    const double r0 = r[0];
    const double R2 = 0.25 * b2 - c + r0;
    if (R2 < 0) return 0;
    const double R = sqrt(R2);
    double D2, E2;
    if (R < 10e-12) {
        const double temp = r0 * r0 - 4 * e;
        if (temp < 0) D2 = E2 = -1;
        else {
            const double st = sqrt(temp);
            D2 = 0.75 * b2 - 2 * c + 2 * st;
            E2 = D2 - 4 * st;
        }
    } else {
        const double uu = 0.75 * b2 - 2 * c - R2, vv = 0.25 * (1.0 / R) * (4 * bc - 8 * d - b3);
        D2 = uu + vv;
        E2 = uu - vv;
    }
    const double b_4 = 0.25 * b, R_2 = 0.5 * R;
    int n = 0;
    if (D2 >= 0) {
        const double D = sqrt(D2);
        x[0] = R_2 + 0.5 * D - b_4;
        x[1] = x[0] - D;
        n = 2;
    }
    if (E2 >= 0) {
        const double E = sqrt(E2);
        x[n] = -R_2 + 0.5 * E - b_4;
        x[n + 1] = x[n] - E;
        n += 2;
    }
    return n;
}

// orthonormal frame of a triangle as the COLUMNS of M (row-major 3 x 3)
DUNK_HD inline void triad(const double* p0, const double* p1, const double* p2, double* M) {
    double e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
    const double n1 = 1.0 / sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
    e1[0] *= n1; e1[1] *= n1; e1[2] *= n1;
    const double w[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
    double e3[3] = {e1[1] * w[2] - e1[2] * w[1], e1[2] * w[0] - e1[0] * w[2], e1[0] * w[1] - e1[1] * w[0]};
    const double n3 = 1.0 / sqrt(e3[0] * e3[0] + e3[1] * e3[1] + e3[2] * e3[2]);
    e3[0] *= n3; e3[1] *= n3; e3[2] *= n3;
    const double e2[3] = {e3[1] * e1[2] - e3[2] * e1[1], e3[2] * e1[0] - e3[0] * e1[2], e3[0] * e1[1] - e3[1] * e1[0]};
    for (int r = 0; r < 3; ++r) { M[3 * r] = e1[r]; M[3 * r + 1] = e2[r]; M[3 * r + 2] = e3[r]; }
}

// cv::solvePnP(4 points, SOLVEPNP_P3P); idx selects the 4 points (null = 0..3).  false: no real solution
DUNK_HD inline bool p3p_solve4(const float* obj, const float* img, const int* idx, const Camera& cam, bool f32_normalised,
                               double (&Rout)[9], double (&tout)[3]) {
    double P[4][3], xy[4][2], f[3][3];
    for (int i = 0; i < 4; ++i) {
        const int j = idx ? idx[i] : i;
        P[i][0] = obj[3 * j]; P[i][1] = obj[3 * j + 1]; P[i][2] = obj[3 * j + 2];
        double xn = ((double)img[2 * j] - cam.uc) * (1.0 / cam.fu), yn = ((double)img[2 * j + 1] - cam.vc) * (1.0 / cam.fv);
        if (f32_normalised) { xn = (double)(float)xn; yn = (double)(float)yn; }
        xy[i][0] = xn; xy[i][1] = yn;
    }
    for (int i = 0; i < 3; ++i) {
        const double nr = 1.0 / sqrt(xy[i][0] * xy[i][0] + xy[i][1] * xy[i][1] + 1.0);
        f[i][0] = xy[i][0] * nr; f[i][1] = xy[i][1] * nr; f[i][2] = nr;
    }
    auto d2 = [&](int i, int j) {
        return (P[i][0] - P[j][0]) * (P[i][0] - P[j][0]) + (P[i][1] - P[j][1]) * (P[i][1] - P[j][1]) + (P[i][2] - P[j][2]) * (P[i][2] - P[j][2]);
    };
    auto dot = [&](int i, int j) { return f[i][0] * f[j][0] + f[i][1] * f[j][1] + f[i][2] * f[j][2]; };
    const double a = d2(0, 1), b = d2(0, 2), c = d2(1, 2), p = dot(0, 1), q = dot(0, 2), r = dot(1, 2);
    if (a == 0 || b == 0 || c == 0) return false;
    const double k4 = -a * a + 4 * a * b * r * r - 2 * a * b + 2 * a * c - b * b + 2 * b * c - c * c;
    const double k3 = -4 * (-a * a * q * r + 2 * a * b * p * r * r - a * b * p + a * b * q * r + a * c * p + a * c * q * r - b * b * p + 2 * b * c * p - c * c * p);
    const double k2 = -2 * (2 * a * a * q * q + 2 * a * a * r * r - a * a - 4 * a * b * p * q * r - 2 * a * b * r * r - 4 * a * c * p * q * r - 2 * a * c * q * q +
                            2 * b * b * p * p + b * b - 4 * b * c * p * p - 2 * b * c + 2 * c * c * p * p + c * c);
    const double k1 = -4 * (-a * a * q * r + a * b * p + a * b * q * r + 2 * a * c * p * q * q - a * c * p + a * c * q * r - b * b * p + 2 * b * c * p - c * c * p);
    const double k0 = -a * a + 2 * a * b + 4 * a * c * q * q - 2 * a * c - b * b + 2 * b * c - c * c;
    double roots[4];
    const int nr = solve_deg4(k4, k3, k2, k1, k0, roots);
    double Mw[9];
    triad(P[0], P[1], P[2], Mw);
    bool found = false;
    double best = 0;
    for (int s = 0; s < nr; ++s) {
        const double u = roots[s];
        if (!(u > 0)) continue;
        const double den = 2 * a * (q - r * u);
        if (den == 0) continue;
        const double v = (-a * u * u + a + 2 * b * p * u - b * u * u - b - 2 * c * p * u + c * u * u + c) / den;
        const double w = 1 + u * u - 2 * u * p;
        if (!(v > 0 && w > 0)) continue;
        const double s0 = sqrt(a / w), s1 = u * s0, s2 = v * s0;
        const double C0[3] = {s0 * f[0][0], s0 * f[0][1], s0 * f[0][2]};
        const double C1[3] = {s1 * f[1][0], s1 * f[1][1], s1 * f[1][2]};
        const double C2[3] = {s2 * f[2][0], s2 * f[2][1], s2 * f[2][2]};
        double Mc[9], R[9], t[3];
        triad(C0, C1, C2, Mc);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[3 * i + j] = Mc[3 * i] * Mw[3 * j] + Mc[3 * i + 1] * Mw[3 * j + 1] + Mc[3 * i + 2] * Mw[3 * j + 2];
        for (int i = 0; i < 3; ++i) t[i] = C0[i] - (R[3 * i] * P[0][0] + R[3 * i + 1] * P[0][1] + R[3 * i + 2] * P[0][2]);
        const double X = R[0] * P[3][0] + R[1] * P[3][1] + R[2] * P[3][2] + t[0];
        const double Y = R[3] * P[3][0] + R[4] * P[3][1] + R[5] * P[3][2] + t[1];
        const double Z = R[6] * P[3][0] + R[7] * P[3][1] + R[8] * P[3][2] + t[2];
        const double ex = X / Z - xy[3][0], ey = Y / Z - xy[3][1];
        const double e = ex * ex + ey * ey;
        bool fin = e == e;
        for (int i = 0; i < 9; ++i) fin = fin && (R[i] - R[i] == 0);
        if (!fin) continue;
        if (!found || e < best) {
            found = true;
            best = e;
            for (int i = 0; i < 9; ++i) Rout[i] = R[i];
            for (int i = 0; i < 3; ++i) tout[i] = t[i];
        }
    }
    return found;
}

// cv::projectPoints (zero distortion) + PnPRansacCallback::computeError for one point: the f64
// projection is rounded to f32, the squared distance is f32.  R2 = Rodrigues(rvec) of the model.
// (separate multiplies and adds: OpenCV's build does not contract them into FMAs)
#ifdef __CUDA_ARCH__
#define DUNK_MUL(a, b) __dmul_rn((a), (b))
#define DUNK_ADD(a, b) __dadd_rn((a), (b))
#define DUNK_FMUL(a, b) __fmul_rn((a), (b))
#define DUNK_FADD(a, b) __fadd_rn((a), (b))
#define DUNK_FSUB(a, b) __fsub_rn((a), (b))
#else
#define DUNK_MUL(a, b) ((a) * (b))
#define DUNK_ADD(a, b) ((a) + (b))
#define DUNK_FMUL(a, b) ((a) * (b))
#define DUNK_FADD(a, b) ((a) + (b))
#define DUNK_FSUB(a, b) ((a) - (b))
#endif
DUNK_HD inline float reproj_err_f32(const double* R2, const double* t, const Camera& cam, const float* obj, const float* img, int i) {
    const double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2];
    double x = DUNK_ADD(DUNK_ADD(DUNK_ADD(DUNK_MUL(R2[0], X), DUNK_MUL(R2[1], Y)), DUNK_MUL(R2[2], Z)), t[0]);
    double y = DUNK_ADD(DUNK_ADD(DUNK_ADD(DUNK_MUL(R2[3], X), DUNK_MUL(R2[4], Y)), DUNK_MUL(R2[5], Z)), t[1]);
    double z = DUNK_ADD(DUNK_ADD(DUNK_ADD(DUNK_MUL(R2[6], X), DUNK_MUL(R2[7], Y)), DUNK_MUL(R2[8], Z)), t[2]);
    z = z ? 1.0 / z : 1.0;
    x = DUNK_MUL(x, z);
    y = DUNK_MUL(y, z);
    const float px = (float)DUNK_ADD(DUNK_MUL(x, cam.fu), cam.uc);
    const float py = (float)DUNK_ADD(DUNK_MUL(y, cam.fv), cam.vc);
    const float dx = DUNK_FSUB(img[2 * i], px), dy = DUNK_FSUB(img[2 * i + 1], py);
    return DUNK_FADD(DUNK_FMUL(dx, dx), DUNK_FMUL(dy, dy));
}

// SOLVEPNP_ITERATIVE's final answer: the minimum of the squared reprojection error over the used points
// (calib3d solvePnP -> Levenberg-Marquardt on (rvec, tvec); cv2 4.13.0's result agrees with the minimiser to
// ~1e-8).  Levenberg-Marquardt from the given pose with a local rotation update R <- exp([dw]x) R, so the
// Jacobian needs no Rodrigues derivative: d(RX + t)/dw = -[RX]x, d/dt = I.  One pass over the points per
// iteration accumulates J^T J (21), J^T r (6) and the cost; the 6 x 6 solve runs redundantly in every thread of
// the executor (all see the same sums), exactly like epnp_solve's scalar part.  Returns the final RMS error.
template <class Exec>
DUNK_HD double pnp_refine(const Exec& ex, const Camera& cam, double (&R)[9], double (&t)[3], int max_iters = 50) {
    double lambda = 1e-3, prev_cost = -1.0;
    double Rp[9], tp[3];          // last accepted pose
    for (int i = 0; i < 9; ++i) Rp[i] = R[i];
    for (int i = 0; i < 3; ++i) tp[i] = t[i];
    double JtJp[21], Jtrp[6];     // normal equations at the last accepted pose
    for (int i = 0; i < 21; ++i) JtJp[i] = 0;
    for (int i = 0; i < 6; ++i) Jtrp[i] = 0;
    const double n = ex.count();
    for (int it = 0; it < max_iters; ++it) {
        double s[28];
        const double r0 = R[0], r1 = R[1], r2 = R[2], r3 = R[3], r4 = R[4], r5 = R[5], r6 = R[6], r7 = R[7], r8 = R[8];
        const double t0 = t[0], t1 = t[1], t2 = t[2];
        const double fu = cam.fu, fv = cam.fv, uc = cam.uc, vc = cam.vc;
        ex.template sum<28>(
            [=](const Point& p, double (&a)[28]) {
                const double ax = r0 * p.X + r1 * p.Y + r2 * p.Z, ay = r3 * p.X + r4 * p.Y + r5 * p.Z,
                             az = r6 * p.X + r7 * p.Y + r8 * p.Z;          // R X
                const double x = ax + t0, y = ay + t1, z = az + t2, iz = 1.0 / z;
                const double ru = fu * x * iz + uc - p.u, rv = fv * y * iz + vc - p.v;
                // d(u, v)/d(x, y, z)
                const double ux = fu * iz, uz = -fu * x * iz * iz, vy = fv * iz, vz = -fv * y * iz * iz;
                // columns: dw (x' = x + dw x a: dx/dw = (0, az, -ay), dy/dw = (-az, 0, ax), dz/dw = (ay, -ax, 0)), dt
                double ju[6], jv[6];
                ju[0] = uz * ay;            ju[1] = ux * az - uz * ax;  ju[2] = -ux * ay;
                jv[0] = -vy * az + vz * ay; jv[1] = -vz * ax;           jv[2] = vy * ax;
                ju[3] = ux; ju[4] = 0.0; ju[5] = uz;
                jv[3] = 0.0; jv[4] = vy; jv[5] = vz;
                int k = 0;
                for (int r = 0; r < 6; ++r)
                    for (int c = r; c < 6; ++c) a[k++] += ju[r] * ju[c] + jv[r] * jv[c];
                for (int r = 0; r < 6; ++r) a[21 + r] += ju[r] * ru + jv[r] * rv;
                a[27] += ru * ru + rv * rv;
            },
            s);
        const double cost = s[27];
        bool accepted = true;
        if (prev_cost >= 0.0 && !(cost <= prev_cost)) {
            // worse (or not finite): back to the last accepted pose, more damping
            accepted = false;
            lambda *= 10.0;
            for (int i = 0; i < 9; ++i) R[i] = Rp[i];
            for (int i = 0; i < 3; ++i) t[i] = tp[i];
        } else {
            if (prev_cost >= 0.0) lambda = lambda * 0.1 > 1e-12 ? lambda * 0.1 : 1e-12;
            prev_cost = cost;
            for (int i = 0; i < 9; ++i) Rp[i] = R[i];
            for (int i = 0; i < 3; ++i) tp[i] = t[i];
            for (int i = 0; i < 21; ++i) JtJp[i] = s[i];
            for (int i = 0; i < 6; ++i) Jtrp[i] = s[21 + i];
        }
        // (J^T J + lambda diag) d = -J^T r by Cholesky
        double A[36], b[6], d[6];
        {
            int k = 0;
            for (int r = 0; r < 6; ++r)
                for (int c = r; c < 6; ++c) { A[r * 6 + c] = A[c * 6 + r] = JtJp[k++]; }
            for (int r = 0; r < 6; ++r) { A[r * 6 + r] *= 1.0 + lambda; b[r] = -Jtrp[r]; }
        }
        bool spd = true;
        for (int c = 0; c < 6 && spd; ++c) {
            double diag = A[c * 6 + c];
            for (int k = 0; k < c; ++k) diag -= A[c * 6 + k] * A[c * 6 + k];
            if (!(diag > 0.0)) { spd = false; break; }
            const double l = sqrt(diag);
            A[c * 6 + c] = l;
            for (int r = c + 1; r < 6; ++r) {
                double v = A[r * 6 + c];
                for (int k = 0; k < c; ++k) v -= A[r * 6 + k] * A[c * 6 + k];
                A[r * 6 + c] = v / l;
            }
        }
        if (!spd) break;
        for (int r = 0; r < 6; ++r) {
            double v = b[r];
            for (int k = 0; k < r; ++k) v -= A[r * 6 + k] * d[k];
            d[r] = v / A[r * 6 + r];
        }
        for (int r = 5; r >= 0; --r) {
            double v = d[r];
            for (int k = r + 1; k < 6; ++k) v -= A[k * 6 + r] * d[k];
            d[r] = v / A[r * 6 + r];
        }
        // R <- exp([dw]x) R, t <- t + dt (from the last accepted pose)
        double dR[9], Rn[9];
        rodrigues_to_matrix(d, dR);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Rn[r * 3 + c] = dR[r * 3] * Rp[c] + dR[r * 3 + 1] * Rp[3 + c] + dR[r * 3 + 2] * Rp[6 + c];
        for (int i = 0; i < 9; ++i) R[i] = Rn[i];
        for (int i = 0; i < 3; ++i) t[i] = tp[i] + d[3 + i];
        double step = 0.0, scale = 1.0;
        for (int i = 0; i < 3; ++i) {
            step = fmax(step, fabs(d[i]));
            scale = fmax(scale, fabs(tp[i]));
        }
        for (int i = 3; i < 6; ++i) step = fmax(step, fabs(d[i]) / scale);
        if (accepted && step < 1e-13) break;
    }
    // the loop may end on a trial pose: keep the last accepted one unless the trial is at least as good
    {
        double s[1];
        const double r0 = R[0], r1 = R[1], r2 = R[2], r3 = R[3], r4 = R[4], r5 = R[5], r6 = R[6], r7 = R[7], r8 = R[8];
        const double t0 = t[0], t1 = t[1], t2 = t[2];
        const double fu = cam.fu, fv = cam.fv, uc = cam.uc, vc = cam.vc;
        ex.template sum<1>(
            [=](const Point& p, double (&a)[1]) {
                const double x = r0 * p.X + r1 * p.Y + r2 * p.Z + t0, y = r3 * p.X + r4 * p.Y + r5 * p.Z + t1,
                             z = r6 * p.X + r7 * p.Y + r8 * p.Z + t2;
                const double ru = fu * x / z + uc - p.u, rv = fv * y / z + vc - p.v;
                a[0] += ru * ru + rv * rv;
            },
            s);
        if (!(s[0] <= prev_cost)) {
            for (int i = 0; i < 9; ++i) R[i] = Rp[i];
            for (int i = 0; i < 3; ++i) t[i] = tp[i];
        } else {
            prev_cost = s[0];
        }
    }
    return sqrt(prev_cost / (n > 0 ? n : 1.0));
}

}  // namespace pnp
}  // namespace dunk
