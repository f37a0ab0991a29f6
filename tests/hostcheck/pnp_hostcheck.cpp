// TEST-ONLY harness: compiles the __host__ __device__ PnP math (cubesat-apds_b200/csrc/pnp_math.cuh)
// for the host so tests/test_pnp_math_host.py can compare it with the oracle without a GPU.
// Nothing in the product loads this.
#include "../../cubesat-apds_b200/csrc/pnp_math.cuh"

using namespace dunk::pnp;

extern "C" {

void hc_jacobi_svd3(const double* A, double* w, double* U, double* Vt) {
    double At[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) At[j * 3 + i] = A[i * 3 + j];
    jacobi_svd<3, 3, true>(At, w, Vt);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) U[i * 3 + j] = At[j * 3 + i];
}

void hc_svd_solve_6x4(const double* A, const double* b, double* x) { svd_solve<6, 4>(A, b, x); }

void hc_rodrigues_to_matrix(const double* r, double* R) { rodrigues_to_matrix(r, R); }
void hc_rodrigues_to_vector(const double* R, double* r) { rodrigues_to_vector(R, r); }

double hc_epnp(const float* obj, const float* img, const int* idx, int n, const double* K, int f32n, double* rvec, double* tvec) {
    Camera cam{K[0], K[4], K[2], K[5]};
    SerialExec ex{obj, img, idx, n, cam, f32n != 0};
    double R[9], t[3];
    const double e = epnp_solve(ex, cam, R, t);
    rodrigues_to_vector(R, rvec);
    for (int i = 0; i < 3; ++i) tvec[i] = t[i];
    return e;
}

double hc_refine(const float* obj, const float* img, const int* idx, int n, const double* K, double* rvec, double* tvec) {
    Camera cam{K[0], K[4], K[2], K[5]};
    SerialExec ex{obj, img, idx, n, cam, false};
    double R[9], t[3] = {tvec[0], tvec[1], tvec[2]};
    rodrigues_to_matrix(rvec, R);
    const double e = pnp_refine(ex, cam, R, t);
    rodrigues_to_vector(R, rvec);
    for (int i = 0; i < 3; ++i) tvec[i] = t[i];
    return e;
}

int hc_p3p(const float* obj, const float* img, const int* idx, const double* K, int f32n, double* rvec, double* tvec) {
    Camera cam{K[0], K[4], K[2], K[5]};
    double R[9], t[3];
    if (!p3p_solve4(obj, img, idx, cam, f32n != 0, R, t)) return 0;
    rodrigues_to_vector(R, rvec);
    for (int i = 0; i < 3; ++i) tvec[i] = t[i];
    return 1;
}

void hc_reproj_err(const double* rvec, const double* tvec, const double* K, const float* obj, const float* img, int n, float* err) {
    Camera cam{K[0], K[4], K[2], K[5]};
    double R2[9];
    rodrigues_to_matrix(rvec, R2);
    for (int i = 0; i < n; ++i) err[i] = reproj_err_f32(R2, tvec, cam, obj, img, i);
}

}
