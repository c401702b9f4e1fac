// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the ROS1 headers the reference's factor sources include; logging macros
// count warnings instead of printing, assertions abort. Nothing under lf-vio_b200/ includes this.
#pragma once
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <string>
namespace ros {
class NodeHandle;
}
extern "C" int lfvio_ref_warn_count;
#define ROS_INFO(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
#define ROS_DEBUG_STREAM(x) ((void)0)
#define ROS_WARN_STREAM(x) (++lfvio_ref_warn_count)
#define ROS_WARN(...) (++lfvio_ref_warn_count)
#define ROS_ERROR(...) (++lfvio_ref_warn_count)
#define ROS_BREAK() std::abort()
#define ROS_ASSERT(c) do { if (!(c)) { std::fprintf(stderr, "ROS_ASSERT failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); std::abort(); } } while (0)
#define ROS_ASSERT_MSG(c, ...) ROS_ASSERT(c)
