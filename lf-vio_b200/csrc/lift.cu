// K5 — batched Scaramuzza/OCamCalib lift onto the (negative-half-plane capable) bearing sphere.
//
// Reference: OCAMCamera::liftProjective, camera_model/src/camera_models/ScaramuzzaCamera.cc:622-645
// (m_inv_scale: :197; liftSphere: :609-614).  The reference is compiled for baseline x86-64, i.e.
// plain IEEE mul/add in source order with NO fused multiply-add, and it accumulates the polynomial
// low-order-first with a running power (not Horner).  To be bit-exact in f64 the kernel uses the
// never-contracted __dmul_rn/__dadd_rn intrinsics in exactly that order; sqrt and division are
// IEEE-correct in CUDA f64.
//
// HBM-bound streaming op: 16 B in, 24 B out per feature, ~20 flops.  One thread per feature.
#include "lift.cuh"

namespace lfvio {

__global__ void __launch_bounds__(256) lift_kernel(OcamDev cam, const double2 *__restrict__ uv,
                                                   double *__restrict__ xyz, int n, int normalize) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 p = uv[i];
    double3 P = ocam_lift(cam, p.x, p.y, normalize != 0);
    xyz[3 * i + 0] = P.x;
    xyz[3 * i + 1] = P.y;
    xyz[3 * i + 2] = P.z;
}

OcamDev make_ocam_dev(const lfvio_ocam &c) {
    OcamDev d;
    for (int i = 0; i < 5; i++) d.poly[i] = c.poly[i];
    d.C = c.C;
    d.D = c.D;
    d.E = c.E;
    d.cx = c.cx;
    d.cy = c.cy;
    d.inv_scale = 1.0 / (c.C - c.D * c.E);  // ScaramuzzaCamera.cc:197 (host f64, no FMA on x86-64 baseline)
    return d;
}

int launch_lift(const lfvio_ocam *cam, const double *uv_dev, double *xyz_dev, int n, int normalize,
                cudaStream_t stream) {
    if (n == 0) return LFVIO_OK;
    OcamDev d = make_ocam_dev(*cam);
    lift_kernel<<<cdiv(n, 256), 256, 0, stream>>>(d, reinterpret_cast<const double2 *>(uv_dev), xyz_dev,
                                                  n, normalize);
    LFVIO_LAUNCH_CHECK();
    return LFVIO_OK;
}

}  // namespace lfvio

extern "C" int lfvio_lift_projective_dev(const lfvio_ocam *cam, const double *uv_dev, double *xyz_dev,
                                         int n, int normalize, void *stream) {
    lfvio::clear_error();
    LFVIO_REQUIRE(cam && n >= 0, "lfvio_lift_projective_dev: bad arguments");
    LFVIO_REQUIRE(n == 0 || (uv_dev && xyz_dev), "lfvio_lift_projective_dev: null buffer");
    return lfvio::launch_lift(cam, uv_dev, xyz_dev, n, normalize, static_cast<cudaStream_t>(stream));
}

extern "C" int lfvio_lift_projective(const lfvio_ocam *cam, const double *uv, double *xyz, int n,
                                     int normalize, int device) {
    lfvio::clear_error();
    LFVIO_REQUIRE(cam && n >= 0, "lfvio_lift_projective: bad arguments");
    if (n == 0) return LFVIO_OK;
    LFVIO_REQUIRE(uv && xyz, "lfvio_lift_projective: null buffer");
    LFVIO_CUDA_TRY(cudaSetDevice(device));
    double *d_uv = nullptr, *d_xyz = nullptr;
    LFVIO_CUDA_TRY(cudaMalloc(&d_uv, sizeof(double) * 2 * n));
    cudaError_t e = cudaMalloc(&d_xyz, sizeof(double) * 3 * n);
    if (e != cudaSuccess) {
        cudaFree(d_uv);
        lfvio::set_error("cudaMalloc: %s", cudaGetErrorString(e));
        return LFVIO_ERR_CUDA;
    }
    int st = LFVIO_OK;
    auto body = [&]() -> int {
        LFVIO_CUDA_TRY(cudaMemcpy(d_uv, uv, sizeof(double) * 2 * n, cudaMemcpyHostToDevice));
        LFVIO_TRY(lfvio::launch_lift(cam, d_uv, d_xyz, n, normalize, 0));
        LFVIO_CUDA_TRY(cudaMemcpy(xyz, d_xyz, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
        return LFVIO_OK;
    };
    st = body();
    cudaFree(d_uv);
    cudaFree(d_xyz);
    return st;
}
