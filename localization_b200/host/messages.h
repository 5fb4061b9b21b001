/*
 * messages.h — plain-struct stand-ins for the ROS messages the reference's callbacks read.
 * Only the fields Localization touches are kept, with their ROS names:
 *   uwb_driver/UwbRange                      localization.cpp:306-346
 *   sensor_msgs/Imu                          localization.cpp:501-518
 *   geometry_msgs/PoseWithCovarianceStamped  localization.cpp:258-277, 464-476
 *   geometry_msgs/TwistWithCovarianceStamped localization.cpp:442-446, 564-580
 */
#ifndef UWBGO_HOST_MESSAGES_H
#define UWBGO_HOST_MESSAGES_H

#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace uwbgo {
namespace host {

struct Time {
    uint32_t sec = 0, nsec = 0;
    double toSec() const { return (double)sec + 1e-9 * (double)nsec; } /* ros::Time::toSec */
};
struct Header {
    uint32_t seq = 0;
    Time stamp;
    std::string frame_id;
};
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { Header header; Pose pose; };
struct Path { Header header; std::vector<PoseStamped> poses; };

struct UwbRange {
    Header header;
    int requester_id = 0, responder_id = 0;
    float distance = 0.f, distance_err = 0.f; /* float32 on the wire */
    int antenna = 0;
};
struct Imu {
    Header header;
    Quaternion orientation;
    std::array<double, 9> orientation_covariance{};
};
struct PoseWithCovarianceStamped {
    Header header;
    Pose pose;
    std::array<double, 36> covariance{};
};
struct Twist { Point linear, angular; };
struct TwistWithCovarianceStamped {
    Header header;
    Twist twist;
    std::array<double, 36> covariance{};
};

}  // namespace host
}  // namespace uwbgo
#endif
