// TEST SHIM (not ROS): the generated dvs_msgs/Event message as roscpp lays it out — uint16 x, uint16 y, ros::Time ts
// {uint32 sec, uint32 nsec}, uint8 polarity — so that dvs_mcemvs_b200/host/ros_adapters.hpp can be compiled and its
// conversions run in an image without ROS (tests/test_ros_adapters.py).
#pragma once
#include <cstdint>
#include <vector>
namespace ros {
struct Time {
  uint32_t sec = 0, nsec = 0;
};
}  // namespace ros
namespace dvs_msgs {
struct Event {
  uint16_t x = 0, y = 0;
  ros::Time ts;
  uint8_t polarity = 0;
};
}  // namespace dvs_msgs
