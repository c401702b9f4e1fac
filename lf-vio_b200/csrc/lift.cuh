// Device-side Scaramuzza lift shared by the standalone kernel (lift.cu) and the tracker (tracker.cu).
#pragma once
#include "common.cuh"

namespace lfvio {

struct OcamDev {
    double poly[5];
    double C, D, E, cx, cy, inv_scale;
};

OcamDev make_ocam_dev(const lfvio_ocam &c);
int launch_lift(const lfvio_ocam *cam, const double *uv_dev, double *xyz_dev, int n, int normalize,
                cudaStream_t stream);

#ifdef __CUDACC__
// ScaramuzzaCamera.cc:622-645, un-fused f64 in source order.  normalize: P / sqrt((x^2 + y^2) + z^2)
// (the callers' `tmp_p / tmp_p.norm()`, feature_tracker.cpp:342,458).
__device__ __forceinline__ double3 ocam_lift(const OcamDev &cam, double u, double v, bool normalize) {
    double xc0 = __dsub_rn(u, cam.cx);
    double xc1 = __dsub_rn(v, cam.cy);
    double xa = __dmul_rn(cam.inv_scale, __dsub_rn(xc0, __dmul_rn(cam.D, xc1)));
    double ya = __dmul_rn(cam.inv_scale, __dadd_rn(__dmul_rn(-cam.E, xc0), __dmul_rn(cam.C, xc1)));
    double phi = sqrt(__dadd_rn(__dmul_rn(xa, xa), __dmul_rn(ya, ya)));
    double phi_i = 1.0, z = 0.0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        z = __dadd_rn(z, __dmul_rn(phi_i, cam.poly[i]));
        phi_i = __dmul_rn(phi_i, phi);
    }
    double3 P = make_double3(xa, ya, -z);
    if (normalize) {
        double nrm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(P.x, P.x), __dmul_rn(P.y, P.y)), __dmul_rn(P.z, P.z)));
        P.x = P.x / nrm;
        P.y = P.y / nrm;
        P.z = P.z / nrm;
    }
    return P;
}
#endif

}  // namespace lfvio
