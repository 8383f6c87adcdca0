// ros_adapters.hpp — conversions from the reference's ROS / minkindr / image_geometry types to the
// PODs of the host mirror.  Compiled only where those headers exist (a catkin workspace); this
// image has none of them, so the file is inert here (`__has_include` guards) and is meant for
// the maintainer-side build described in INTEGRATION.md §2.  tests/test_ros_adapters.py compiles it
// against shims with the same type and member names (tests/shims/ros/) and runs the conversions.
#pragma once

#include "mapper_emvs_stereo/mapper_emvs_stereo.hpp"

#if defined(__has_include)
#if __has_include(<dvs_msgs/Event.h>) && __has_include(<kindr/minimal/quat-transformation.h>) && \
    __has_include(<image_geometry/pinhole_camera_model.h>)
#define EMVS_HOST_HAVE_ROS 1
#include <dvs_msgs/Event.h>
#include <image_geometry/pinhole_camera_model.h>
#include <kindr/minimal/quat-transformation.h>

namespace emvs_host {

// dvs_msgs/Event = {uint16 x, uint16 y, ros::Time ts {uint32 sec, nsec}, bool polarity}: 16 bytes,
// the same layout as emvs_event, so event vectors are passed without a copy.
static_assert(sizeof(dvs_msgs::Event) == sizeof(emvs_event), "dvs_msgs::Event layout changed");
inline const emvs_event* as_pod(const std::vector<dvs_msgs::Event>& ev) { return reinterpret_cast<const emvs_event*>(ev.data()); }

inline geometry_utils::Transformation to_pod(const kindr::minimal::QuatTransformation& T)
{
  const auto q = T.getRotation().toImplementation();   // Eigen::Quaterniond
  const double qw[4] = {q.w(), q.x(), q.y(), q.z()};
  const auto p = T.getPosition();
  const double t[3] = {p[0], p[1], p[2]};
  return geometry_utils::Transformation(qw, t);
}

// What MapperEMVS reads from the camera model (mapper_emvs_stereo.cpp:34-48) plus the LUT that
// precomputeRectifiedPoints (:256-299) builds: pass the reference's own
// precomputed_rectified_points_ (2 x N, column-major == interleaved x,y).
inline geometry_utils::CameraInfo to_pod(const image_geometry::PinholeCameraModel& cam, const float* rectified_xy,
                                         size_t n_pixels)
{
  geometry_utils::CameraInfo c;
  const cv::Size s = cam.fullResolution();
  c.width = s.width; c.height = s.height;
  c.fx = cam.fx(); c.fy = cam.fy(); c.cx = cam.cx(); c.cy = cam.cy();
  if (rectified_xy) c.rectified_points.assign(rectified_xy, rectified_xy + 2 * n_pixels);
  return c;
}

inline bool evaluateDSI(EMVS::MapperEMVS& m, const std::vector<dvs_msgs::Event>& events, const LinearTrajectory& trajectory,
                        const kindr::minimal::QuatTransformation& T_rv_w)
{
  m.dsi_.touch();
  const geometry_utils::Transformation T = to_pod(T_rv_w);
  const int rc = emvs_mapper_evaluate_dsi(m.handle(), as_pod(events), events.size(), trajectory.pods().data(),
                                          trajectory.pods().size(), &T.pod());
  if (rc == EMVS_ERR_TOO_FEW) return false;
  check(rc, "evaluateDSI");
  return true;
}

}  // namespace emvs_host
#endif
#endif
