// Implementation of the FeatureTracker / OCAMCamera host mirrors: pure plumbing over the C-ABI (no image math here).
#include "feature_tracker.h"

#include <cmath>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

namespace lfvio_host {

int ROW = 960, COL = 1280, MAX_CNT = 150, MIN_DIST = 20, EQUALIZE = 1, FISHEYE = 0, FREQ = 10;
bool PUB_THIS_FRAME = false;
double CENTER_X = 640.991301, CENTER_Y = 490.937512;
int MAX_R = 500, MIN_R = 160;
int REJECT_WITH_F = 1;
unsigned long long RANSAC_SEED = 0x5eed5eed5eedull;
int FeatureTracker::n_id = 0;

namespace {

// The reference reads its YAML through cv::FileStorage.  The files are flat "key: value" maps with one level of
// nesting (poly_parameters: / p0: ...), which this reader flattens to "section.key".
bool read_flat_yaml(const std::string &file, std::map<std::string, std::string> &kv) {
    std::ifstream in(file);
    if (!in) return false;
    std::string line, section;
    while (std::getline(in, line)) {
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        if (line.find_first_not_of(" \t\r") == std::string::npos || line[0] == '%') continue;
        const size_t colon = line.find(':');
        if (colon == std::string::npos) continue;
        const bool indented = line[0] == ' ' || line[0] == '\t';
        auto trim = [](std::string s) {
            const size_t a = s.find_first_not_of(" \t\r\""), b = s.find_last_not_of(" \t\r\"");
            return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
        };
        const std::string key = trim(line.substr(0, colon)), val = trim(line.substr(colon + 1));
        if (!indented) {
            section = val.empty() ? key : std::string();
            if (!val.empty()) kv[key] = val;
        } else if (!section.empty()) {
            kv[section + "." + key] = val;
        }
    }
    return true;
}
double num(const std::map<std::string, std::string> &kv, const std::string &k, double dflt) {
    auto it = kv.find(k);
    return it == kv.end() ? dflt : std::atof(it->second.c_str());
}

}  // namespace

bool readParameters(const std::string &config_file) {
    std::map<std::string, std::string> kv;
    if (!read_flat_yaml(config_file, kv)) return false;
    ROW = (int)num(kv, "image_height", ROW);
    COL = (int)num(kv, "image_width", COL);
    MAX_CNT = (int)num(kv, "max_cnt", MAX_CNT);
    MIN_DIST = (int)num(kv, "min_dist", MIN_DIST);
    FREQ = (int)num(kv, "freq", FREQ);
    EQUALIZE = (int)num(kv, "equalize", EQUALIZE);
    FISHEYE = (int)num(kv, "fisheye", FISHEYE);
    CENTER_X = num(kv, "center_x", CENTER_X);
    CENTER_Y = num(kv, "center_y", CENTER_Y);
    MAX_R = (int)num(kv, "max_r", MAX_R);
    MIN_R = (int)num(kv, "min_r", MIN_R);
    if (FREQ == 0) FREQ = 100;  // parameters.cpp:71-72
    return true;
}

namespace camodocal {

void OCAMCamera::liftProjective(const Vector2d &p, Vector3d &P) const {
    lfvio_lift_projective(&mParameters, p.data(), P.data(), 1, 0, m_device);
}
void OCAMCamera::liftSphere(const Vector2d &p, Vector3d &P) const {
    lfvio_lift_projective(&mParameters, p.data(), P.data(), 1, 1, m_device);
}
void OCAMCamera::spaceToPlane(const Vector3d &P, Vector2d &p) const {
    const double norm = std::sqrt(P[0] * P[0] + P[1] * P[1]);
    const double theta = std::atan2(-P[2], norm);
    double rho = 0.0, theta_i = 1.0;
    for (int i = 0; i < 20; i++) {
        rho += theta_i * mParameters.inv_poly[i];
        theta_i *= theta;
    }
    const double invNorm = 1.0 / norm;
    const double xn0 = P[0] * invNorm * rho, xn1 = P[1] * invNorm * rho;
    p[0] = xn0 * mParameters.C + xn1 * mParameters.D + mParameters.cx;
    p[1] = xn0 * mParameters.E + xn1 + mParameters.cy;
}
bool OCAMCamera::liftProjective(const std::vector<Vector2d> &p, std::vector<Vector3d> &P, bool normalize) const {
    P.resize(p.size());
    if (p.empty()) return true;
    return lfvio_lift_projective(&mParameters, p[0].data(), P[0].data(), (int)p.size(), normalize ? 1 : 0, m_device) == LFVIO_OK;
}

CameraPtr generateCameraFromYamlFile(const std::string &filename, int device) {
    std::map<std::string, std::string> kv;
    if (!read_flat_yaml(filename, kv)) return nullptr;
    auto it = kv.find("model_type");
    if (it == kv.end() || (it->second != "scaramuzza" && it->second != "SCARAMUZZA")) return nullptr;
    lfvio_ocam c;
    memset(&c, 0, sizeof(c));
    c.image_width = (int)num(kv, "image_width", 0);
    c.image_height = (int)num(kv, "image_height", 0);
    char key[64];
    for (int i = 0; i < 5; i++) {
        snprintf(key, sizeof(key), "poly_parameters.p%d", i);
        c.poly[i] = num(kv, key, 0.0);
    }
    for (int i = 0; i < 20; i++) {
        snprintf(key, sizeof(key), "inv_poly_parameters.p%d", i);
        c.inv_poly[i] = num(kv, key, 0.0);
    }
    c.C = num(kv, "affine_parameters.ac", 1.0);
    c.D = num(kv, "affine_parameters.ad", 0.0);
    c.E = num(kv, "affine_parameters.ae", 0.0);
    c.cx = num(kv, "affine_parameters.cx", 0.0);
    c.cy = num(kv, "affine_parameters.cy", 0.0);
    return std::make_shared<OCAMCamera>(c, device);
}

}  // namespace camodocal

FeatureTracker::FeatureTracker() {}
FeatureTracker::~FeatureTracker() {
    if (m_handle) lfvio_tracker_destroy(m_handle);
}
const char *FeatureTracker::lastError() const { return lfvio_last_error(); }

void FeatureTracker::readIntrinsicParameter(const std::string &calib_file) {
    m_camera = camodocal::generateCameraFromYamlFile(calib_file, device);
    if (auto oc = std::dynamic_pointer_cast<camodocal::OCAMCamera>(m_camera)) {
        m_params = oc->getParameters();
        m_have_params = true;
    }
}
void FeatureTracker::setCamera(const camodocal::CameraPtr &cam, const lfvio_ocam &params) {
    m_camera = cam;
    m_params = params;
    m_have_params = true;
}

bool FeatureTracker::ensureHandle() {
    if (m_handle) return true;
    if (!m_have_params) return false;
    lfvio_tracker_cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.rows = ROW;
    cfg.cols = COL;
    cfg.max_cnt = MAX_CNT;
    cfg.min_dist = MIN_DIST;
    cfg.equalize = EQUALIZE;
    cfg.center_x = CENTER_X;
    cfg.center_y = CENTER_Y;
    cfg.max_r = MAX_R;
    cfg.min_r = MIN_R;
    cfg.capacity = 0;
    cfg.ransac = REJECT_WITH_F;
    cfg.ransac_seed = RANSAC_SEED;
    cfg.cam = m_params;
    return lfvio_tracker_create(&cfg, device, &m_handle) == LFVIO_OK;
}

void FeatureTracker::pull() {
    const int n = lfvio_tracker_count(m_handle);
    cur_pts.resize(n);
    ids.resize(n);
    track_cnt.resize(n);
    cur_un_pts_3d.resize(n);
    pts_velocity_3d.resize(n);
    lfvio_tracks t;
    t.capacity = n;
    t.n = 0;
    t.cur_pts = n ? &cur_pts[0].x : nullptr;
    t.ids = ids.data();
    t.track_cnt = track_cnt.data();
    t.cur_un_pts_3d = n ? &cur_un_pts_3d[0].x : nullptr;
    t.pts_velocity_3d = n ? &pts_velocity_3d[0].x : nullptr;
    lfvio_tracker_get(m_handle, &t);
}

void FeatureTracker::flushIds() {
    if (!m_ids_dirty) return;
    // the caller ran the reference's loop `for (i = 0;; i++) if (!updateID(i)) break;`: every -1 got n_id++ in index
    // order, which is exactly what the device pass does when started from the same base
    int base = m_ids_base;
    lfvio_tracker_update_ids(m_handle, &base);
    m_ids_dirty = false;
}

void FeatureTracker::readImage(const ImageView &_img, double _cur_time) {
    if (!ensureHandle()) return;  // the reference has no error return either; lastError() tells why
    flushIds();
    cur_time = _cur_time;
    if (lfvio_tracker_read_image(m_handle, _img.data, _img.step, _cur_time, PUB_THIS_FRAME ? 1 : 0) != LFVIO_OK) return;
    pull();
    prev_time = cur_time;
    m_ids_base = n_id;
}

bool FeatureTracker::updateID(unsigned int i) {
    if (i < ids.size()) {
        if (ids[i] == -1) {
            ids[i] = n_id++;
            m_ids_dirty = true;
        }
        return true;
    }
    flushIds();
    return false;
}

bool FeatureTracker::fetchMask() {
    if (!m_handle) return false;
    mask.resize((size_t)ROW * COL);
    return lfvio_tracker_get_mask(m_handle, mask.data(), COL) == LFVIO_OK;
}

}  // namespace lfvio_host

// ---- C entry point used by the tests to drive the C++ class from Python (ctypes) --------------------------
// Runs FeatureTracker over `n` frames exactly as feature_tracker_node.cpp does (readImage, then the updateID loop)
// and returns the members the node would publish after the last frame.  Test hook of the host mirror only.
extern "C" int lfvio_host_tracker_run(const lfvio_ocam *cam, const uint8_t *frames, int n, int rows, int cols, const int *pub,
                                      double t0, double dt, int max_cnt, int min_dist, float *cur_pts, int *ids, int *track_cnt,
                                      float *bearing, float *velocity, int capacity, int *n_out, int *n_id_out) {
    using namespace lfvio_host;
    ROW = rows;
    COL = cols;
    MAX_CNT = max_cnt;
    MIN_DIST = min_dist;
    EQUALIZE = 1;
    REJECT_WITH_F = 0;  // the golden sequences of tests/golden were generated without RANSAC
    FeatureTracker::n_id = 0;
    FeatureTracker tr;
    tr.setCamera(std::make_shared<camodocal::OCAMCamera>(*cam), *cam);
    for (int k = 0; k < n; k++) {
        PUB_THIS_FRAME = pub[k] != 0;
        ImageView img{frames + (size_t)k * rows * cols, rows, cols, cols};
        tr.readImage(img, t0 + k * dt);
        for (unsigned int i = 0;; i++)
            if (!tr.updateID(i)) break;
    }
    const int m = (int)tr.ids.size();
    if (m > capacity) return LFVIO_ERR_CAPACITY;
    for (int i = 0; i < m; i++) {
        cur_pts[2 * i] = tr.cur_pts[i].x;
        cur_pts[2 * i + 1] = tr.cur_pts[i].y;
        ids[i] = tr.ids[i];
        track_cnt[i] = tr.track_cnt[i];
        bearing[3 * i] = tr.cur_un_pts_3d[i].x;
        bearing[3 * i + 1] = tr.cur_un_pts_3d[i].y;
        bearing[3 * i + 2] = tr.cur_un_pts_3d[i].z;
        velocity[3 * i] = tr.pts_velocity_3d[i].x;
        velocity[3 * i + 1] = tr.pts_velocity_3d[i].y;
        velocity[3 * i + 2] = tr.pts_velocity_3d[i].z;
    }
    *n_out = m;
    *n_id_out = FeatureTracker::n_id;
    return m > 0 ? LFVIO_OK : LFVIO_ERR_CUDA;
}

// test hook (tests/test_signature_dropin.py): OCAMCamera::spaceToPlane of the host mirror, batched
extern "C" int lfvio_host_ocam_space_to_plane(const lfvio_ocam *cam, const double *xyz, double *uv, int n) {
    if (!cam || !xyz || !uv) return LFVIO_ERR_ARG;
    lfvio_host::camodocal::OCAMCamera c(*cam);
    for (int i = 0; i < n; i++) {
        lfvio_host::Vector2d p;
        c.spaceToPlane(lfvio_host::Vector3d{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, p);
        uv[2 * i] = p[0];
        uv[2 * i + 1] = p[1];
    }
    return LFVIO_OK;
}
