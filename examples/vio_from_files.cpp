// ROS-free driver of the two node shims (lf-vio_b200/host/nodes.h): replays a recorded stream from two text files through
// EstimatorNode exactly as estimator_node.cpp would receive it from its topics, and writes the trajectory in the reference's
// TUM format (visualization.cpp:173-179).  The front end can be fed either with feature messages already produced by a
// tracker (this file) or with images through FeatureTrackerNode::img_callback.
//
//   imu.txt       one line per sample:   stamp ax ay az gx gy gz
//   features.txt  one block per image:   "F stamp n" followed by n lines "id x y z u v vx vy vz"
//                 (the fields of sensor_msgs/PointCloud on /feature_tracker/feature: unit bearing, pixel, bearing velocity)
//
// Build (from the repo root, after `make -C lf-vio_b200`):
//   g++ -O2 -std=c++17 -Iinclude -Ilf-vio_b200/host examples/vio_from_files.cpp -Llf-vio_b200 -llfvio_host -llfvio
//       -Wl,-rpath,$PWD/lf-vio_b200 -o vio_from_files            (one command line)
// Run on a machine with a B200:  ./vio_from_files imu.txt features.txt result_tum.txt [calib.txt] [td] [ransac_seed]
//   calib.txt (optional, whitespace separated): ric (9, row-major: camera -> IMU rotation) tic (3) acc_n gyr_n acc_w gyr_w g_norm
// tools/make_example_stream.py writes a matching set of files from the synthetic world used by the tests.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>

#include "nodes.h"

using namespace lfvio_host;

static void usage() {
    std::puts("usage: vio_from_files imu.txt features.txt result_tum.txt [calib.txt] [td] [ransac_seed] [host_list]\n"
              "  imu.txt:      stamp ax ay az gx gy gz\n"
              "  features.txt: blocks \"F stamp n\" + n lines \"id x y z u v vx vy vz\"\n"
              "  calib.txt:    ric (9, row-major) tic (3) acc_n gyr_n acc_w gyr_w g_norm\n"
              "  host_list:    1 keeps the FeatureManager on the host (default 0: on the device once the estimator is initialised)\n"
              "  writes \"stamp px py pz qx qy qz qw\" per optimised frame (TUM)");
}

int main(int argc, char **argv) {
    if (argc >= 2 && std::string(argv[1]) == "--help") {
        usage();
        return 0;
    }
    if (argc < 4) {
        usage();
        return 2;
    }
    std::ifstream imu(argv[1]), feat(argv[2]);
    if (!imu || !feat) {
        std::fprintf(stderr, "cannot open the input files\n");
        return 2;
    }
    EstimatorNode node(/*window_size=*/10, /*max_features=*/1000, /*device=*/0);
    if (node.estimator.last_status != LFVIO_OK) {
        std::fprintf(stderr, "no estimator: %s\n", node.estimator.lastError());
        return 1;
    }
    node.VINS_RESULT_PATH = argv[3];
    if (argc > 4) {  // readParameters(): extrinsicRotation / extrinsicTranslation / acc_n .. g_norm of the reference's YAML
        std::ifstream cal(argv[4]);
        double v[17];
        for (double &x : v)
            if (!(cal >> x)) {
                std::fprintf(stderr, "calib.txt needs 17 numbers: ric (9) tic (3) acc_n gyr_n acc_w gyr_w g_norm\n");
                return 2;
            }
        for (int k = 0; k < 9; k++) node.estimator.RIC0[k] = v[k];
        node.estimator.TIC0 = Vec3{v[9], v[10], v[11]};
        node.estimator.imu_noise = lfvio_imu_noise{v[12], v[13], v[14], v[15]};
        G = Vec3{0.0, 0.0, v[16]};
    }
    if (argc > 5) TD = std::atof(argv[5]);
    node.estimator.setParameter();            // tic / ric / td from TIC0 / RIC0 / TD
    node.estimator.run_initial_structure = true;
    if (argc > 6) node.estimator.ransac_seed = std::strtoull(argv[6], nullptr, 10);
    node.estimator.use_device_store = !(argc > 7 && std::atoi(argv[7]) == 1);  // track list on the device (lfvio_fm_*)

    // merge the two streams by time stamp, as the ROS callbacks would arrive
    ImuMsg im;
    bool have_imu = bool(imu >> im.stamp >> im.linear_acceleration[0] >> im.linear_acceleration[1] >> im.linear_acceleration[2] >>
                         im.angular_velocity[0] >> im.angular_velocity[1] >> im.angular_velocity[2]);
    std::string tag;
    PointCloudMsg pc;
    auto read_block = [&]() -> bool {
        int n = 0;
        if (!(feat >> tag >> pc.stamp >> n) || tag != "F") return false;
        pc.points.clear();
        for (auto &c : pc.channels) c.clear();
        for (int i = 0; i < n; i++) {
            double id, x, y, z, u, v, vx, vy, vz;
            if (!(feat >> id >> x >> y >> z >> u >> v >> vx >> vy >> vz)) return false;
            Point3f p;
            p.x = (float)x, p.y = (float)y, p.z = (float)z;
            pc.points.push_back(p);
            const float ch[6] = {(float)id, (float)u, (float)v, (float)vx, (float)vy, (float)vz};
            for (int c = 0; c < 6; c++) pc.channels[c].push_back(ch[c]);
        }
        return true;
    };
    bool have_feat = read_block();
    int images = 0;
    while (have_imu || have_feat) {
        if (have_imu && (!have_feat || im.stamp <= pc.stamp)) {
            node.imu_callback(im);
            have_imu = bool(imu >> im.stamp >> im.linear_acceleration[0] >> im.linear_acceleration[1] >> im.linear_acceleration[2] >>
                            im.angular_velocity[0] >> im.angular_velocity[1] >> im.angular_velocity[2]);
        } else {
            node.feature_callback(pc);
            have_feat = read_block();
        }
        images += node.process();
        if (node.estimator.last_status != LFVIO_OK) {
            std::fprintf(stderr, "estimator error: %s\n", node.estimator.lastError());
            return 1;
        }
    }
    std::printf("%d images processed, %zu poses written to %s, %d reboot(s)\n", images, node.odometry.size(), argv[3], node.estimator.reboot_count);
    return 0;
}
