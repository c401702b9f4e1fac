// Stand-in for <std_msgs/Header.h> (tests/test_signature_dropin.py): header.stamp.toSec() is all processImage() reads.
#pragma once
#include <string>
namespace ros {
struct Time {
    double t = 0;
    double toSec() const { return t; }
};
}  // namespace ros
namespace std_msgs {
struct Header {
    unsigned seq = 0;
    ros::Time stamp;
    std::string frame_id;
};
}  // namespace std_msgs
