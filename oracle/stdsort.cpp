// ORACLE helper (test infrastructure only): the permutation libstdc++'s std::sort produces for the
// reference's unstable sort of tracks by track_cnt (feature_tracker/src/feature_tracker.cpp:66-67:
// sort(cnt_pts_id.begin(), cnt_pts_id.end(), [](a, b){ return a.first > b.first; })).
// The permutation depends only on the keys and on libstdc++'s introsort, not on the payload type.
#include <algorithm>
#include <utility>
#include <vector>

extern "C" void sort_desc_perm(const int *keys, int *perm, int n) {
    std::vector<std::pair<int, int>> v(n);
    for (int i = 0; i < n; i++) v[i] = {keys[i], i};
    std::sort(v.begin(), v.end(),
              [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first > b.first; });
    for (int i = 0; i < n; i++) perm[i] = v[i].second;
}
