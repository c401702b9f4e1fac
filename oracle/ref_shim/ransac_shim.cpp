// TEST INFRASTRUCTURE ONLY (oracle/): the reference's own spherical 8-point RANSAC — compute_E_21, check_inliers and
// myfindFundamentalMat, extracted BY LINE RANGE from /root/reference/feature_tracker/src/feature_tracker.cpp (:186-327, see
// oracle/Makefile -> _ref/gen/ransac_fns.inc; the rest of that file needs OpenCV) and compiled unmodified against
// oracle/ref_stubs.  The one thing replaced is the random source: util::create_random_array (random_array.cc) seeds a fresh
// mt19937 from std::random_device on every call, so the reference is not reproducible against itself; the definition below
// feeds the same sampling procedure (draw 9, sort, unique, keep 8, repeat) from the injected counter-based generator that
// oracle/ransac.py and the CUDA path use (the final std::shuffle only permutes the rows of the constraint matrix).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include <eigen3/Eigen/Dense>
#include <ros/ros.h>

#include "random_array.h"
#include "tic_toc.h"

typedef unsigned char uchar;
using namespace std;

namespace {
unsigned long long g_seed = 0;
unsigned g_frame = 0, g_hyp = 0;
inline unsigned long long splitmix64(unsigned long long x) {
    unsigned long long z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline unsigned rng_word(unsigned long long seed, unsigned frame, unsigned hyp, unsigned k) {
    return (unsigned)(splitmix64(seed + frame * 0xD1B54A32D192ED03ull + hyp * 0x8CB92BA72F3D8DD7ull + k * 0xA24BAED4963EE407ull) >> 32);
}
}  // namespace

namespace util {
template <>
std::vector<unsigned int> create_random_array<unsigned int>(const size_t size, const unsigned int rand_min, const unsigned int rand_max) {
    const unsigned n = rand_max - rand_min + 1;
    const size_t make_size = static_cast<size_t>(size * 1.2);
    std::vector<unsigned int> v;
    unsigned k = 0;
    const unsigned hyp = g_hyp++;
    while (v.size() != size) {
        while (v.size() < make_size) v.push_back(rand_min + (unsigned)(((unsigned long long)rng_word(g_seed, g_frame, hyp, k++) * n) >> 32));
        std::sort(v.begin(), v.end());
        auto unique_end = std::unique(v.begin(), v.end());
        if (size < static_cast<size_t>(std::distance(v.begin(), unique_end))) unique_end = std::next(v.begin(), size);
        v.erase(unique_end, v.end());
    }
    return v;
}
}  // namespace util

#include "ransac_fns.inc"

extern "C" {
// myfindFundamentalMat on unit bearings b1 / b2 (n x 3) with the injected random source; status: n bytes, E: 9 (row-major)
int ref_find_fundamental(const double *b1, const double *b2, int n, unsigned long long seed, unsigned frame, unsigned char *status,
                         double *E_out) {
    g_seed = seed, g_frame = frame, g_hyp = 0;
    vector<Eigen::Vector3d> c1(n), c2(n);
    for (int i = 0; i < n; i++) {
        c1[i] = Eigen::Vector3d(b1[3 * i], b1[3 * i + 1], b1[3 * i + 2]);
        c2[i] = Eigen::Vector3d(b2[3 * i], b2[3 * i + 1], b2[3 * i + 2]);
    }
    vector<uchar> st;
    const Eigen::Matrix3d E = myfindFundamentalMat(c1, c2, 8 /* cv::FM_RANSAC */, 0.3 / 460, 0.99, st);
    for (int i = 0; i < n; i++) status[i] = st[i];
    if (E_out)
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) E_out[3 * r + c] = E(r, c);
    return 0;
}
}
