// TEST INFRASTRUCTURE ONLY (oracle/): C entry points over OCAMCamera::liftSphere / liftProjective / spaceToPlane, whose
// method bodies are extracted BY LINE RANGE from /root/reference/camera_model/src/camera_models/ScaramuzzaCamera.cc
// (:609-674, see oracle/Makefile -> _ref/gen/ocam_methods.inc) because the rest of that file needs OpenCV, Boost and
// camodocal/gpl. Only the class skeleton below (parameter accessors of ScaramuzzaCamera.h:24-65, m_inv_scale of
// ScaramuzzaCamera.cc:194-198) is restated.
#include <cmath>

#include <eigen3/Eigen/Dense>

#define SCARAMUZZA_POLY_SIZE 5
#define SCARAMUZZA_INV_POLY_SIZE 20

namespace camodocal {
class OCAMCamera {
  public:
    class Parameters {
      public:
        double C() const { return m_C; }
        double D() const { return m_D; }
        double E() const { return m_E; }
        double center_x() const { return m_center_x; }
        double center_y() const { return m_center_y; }
        double poly(int i) const { return m_poly[i]; }
        double inv_poly(int i) const { return m_inv_poly[i]; }
        double m_poly[SCARAMUZZA_POLY_SIZE], m_inv_poly[SCARAMUZZA_INV_POLY_SIZE], m_C, m_D, m_E, m_center_x, m_center_y;
    };
    explicit OCAMCamera(const Parameters &params) : mParameters(params) {
        m_inv_scale = 1.0 / (params.C() - params.D() * params.E());
    }
    void liftSphere(const Eigen::Vector2d &p, Eigen::Vector3d &P) const;
    void liftProjective(const Eigen::Vector2d &p, Eigen::Vector3d &P) const;
    void spaceToPlane(const Eigen::Vector3d &P, Eigen::Vector2d &p) const;

  private:
    Parameters mParameters;
    double m_inv_scale;
};
#include "ocam_methods.inc"
}  // namespace camodocal

extern "C" {
// cam: poly[5], inv_poly[20], C, D, E, cx, cy (30 doubles)
static camodocal::OCAMCamera make_cam(const double *cam) {
    camodocal::OCAMCamera::Parameters p;
    for (int i = 0; i < 5; i++) p.m_poly[i] = cam[i];
    for (int i = 0; i < 20; i++) p.m_inv_poly[i] = cam[5 + i];
    p.m_C = cam[25];
    p.m_D = cam[26];
    p.m_E = cam[27];
    p.m_center_x = cam[28];
    p.m_center_y = cam[29];
    return camodocal::OCAMCamera(p);
}
void ref_lift_projective(const double *cam, const double *uv, double *xyz, int n, int normalize) {
    camodocal::OCAMCamera c = make_cam(cam);
    for (int i = 0; i < n; i++) {
        Eigen::Vector3d P;
        if (normalize)
            c.liftSphere(Eigen::Vector2d(uv[2 * i], uv[2 * i + 1]), P);
        else
            c.liftProjective(Eigen::Vector2d(uv[2 * i], uv[2 * i + 1]), P);
        xyz[3 * i] = P(0), xyz[3 * i + 1] = P(1), xyz[3 * i + 2] = P(2);
    }
}
void ref_space_to_plane(const double *cam, const double *xyz, double *uv, int n) {
    camodocal::OCAMCamera c = make_cam(cam);
    for (int i = 0; i < n; i++) {
        Eigen::Vector2d p;
        c.spaceToPlane(Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), p);
        uv[2 * i] = p(0), uv[2 * i + 1] = p(1);
    }
}
}
