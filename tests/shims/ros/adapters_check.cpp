// Compiles dvs_mcemvs_b200/host/ros_adapters.hpp against the shims in this directory and runs the conversions that need no
// GPU: event layout, pose and camera PODs, and that the adapter's evaluateDSI type-checks against the host mirror
// (it is instantiated, and called only when a device is present: argv[1] == "gpu").
#include <cmath>
#include <cstdio>
#include <cstring>

#include "ros_adapters.hpp"

#ifndef EMVS_HOST_HAVE_ROS
#error "the ROS shims were not found: ros_adapters.hpp stayed inert"
#endif

#define CHECK(c)                                               \
  do {                                                         \
    if (!(c)) {                                                \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
      return 1;                                                \
    }                                                          \
  } while (0)

int main(int argc, char** argv)
{
  // events: same 16 bytes, no copy
  std::vector<dvs_msgs::Event> ev(3);
  ev[1].x = 321; ev[1].y = 123; ev[1].ts.sec = 7; ev[1].ts.nsec = 999999999u; ev[1].polarity = 1;
  const emvs_event* pod = emvs_host::as_pod(ev);
  CHECK((const void*)pod == (const void*)ev.data());
  CHECK(pod[1].x == 321 && pod[1].y == 123 && pod[1].sec == 7 && pod[1].nsec == 999999999u && pod[1].polarity == 1);
  CHECK(offsetof(dvs_msgs::Event, ts) == offsetof(emvs_event, sec) && offsetof(dvs_msgs::Event, polarity) == offsetof(emvs_event, polarity));

  // pose: quaternion in (w, x, y, z) order + position
  const double s = std::sqrt(0.5);
  const kindr::minimal::QuatTransformation T(s, 0.0, s, 0.0, 1.5, -2.5, 3.5);
  const geometry_utils::Transformation P = emvs_host::to_pod(T);
  CHECK(P.pod().q[0] == s && P.pod().q[1] == 0.0 && P.pod().q[2] == s && P.pod().q[3] == 0.0);
  CHECK(P.pod().t[0] == 1.5 && P.pod().t[1] == -2.5 && P.pod().t[2] == 3.5);
  // the POD composes like a rigid transformation: T * T^-1 = identity (through the C-ABI's host-side geometry, no GPU)
  const geometry_utils::Transformation I = P * P.inverse();
  CHECK(std::fabs(I.pod().q[0]) > 1.0 - 1e-12 && std::fabs(I.pod().t[0]) < 1e-12 && std::fabs(I.pod().t[1]) < 1e-12 &&
        std::fabs(I.pod().t[2]) < 1e-12);

  // camera: projection-matrix intrinsics + the caller's precomputed_rectified_points_ (2 x N column-major == interleaved x, y)
  const image_geometry::PinholeCameraModel cam(4, 3, 200.5, 201.5, 2.25, 1.75);
  std::vector<float> lut(2 * 12);
  for (size_t i = 0; i < lut.size(); ++i) lut[i] = 0.5f * (float)i;
  const geometry_utils::CameraInfo c = emvs_host::to_pod(cam, lut.data(), 12);
  CHECK(c.width == 4 && c.height == 3 && c.fx == 200.5f && c.fy == 201.5f && c.cx == 2.25f && c.cy == 1.75f);
  CHECK(c.rectified_points.size() == 24 && c.rectified_points[23] == 11.5f);
  CHECK(emvs_host::to_pod(cam, nullptr, 0).rectified_points.empty());

  // the adapter's evaluateDSI: referenced so that it is compiled; called only on a GPU box
  bool (*fn)(EMVS::MapperEMVS&, const std::vector<dvs_msgs::Event>&, const LinearTrajectory&,
             const kindr::minimal::QuatTransformation&) = &emvs_host::evaluateDSI;
  CHECK(fn != nullptr);
  if (argc > 1 && std::strcmp(argv[1], "gpu") == 0) {
    geometry_utils::CameraInfo ci;
    ci.width = 240; ci.height = 180; ci.fx = ci.fy = 200.f; ci.cx = 120.f; ci.cy = 90.f;
    EMVS::MapperEMVS mapper(ci, EMVS::ShapeDSI(0, 0, 32, 1.0f, 5.0f, 0.0f));
    LinearTrajectory::PoseMap poses;
    const double q[4] = {1, 0, 0, 0}, t0[3] = {0, 0, 0}, t1[3] = {0.1, 0, 0};
    poses[geometry_utils::Time(0, 0)] = geometry_utils::Transformation(q, t0);
    poses[geometry_utils::Time(1, 0)] = geometry_utils::Transformation(q, t1);
    const LinearTrajectory traj(poses);
    std::vector<dvs_msgs::Event> few(100), many(4096);
    for (size_t i = 0; i < many.size(); ++i) {
      many[i].x = (uint16_t)(i % 240); many[i].y = (uint16_t)(i % 180);
      many[i].ts.sec = 0; many[i].ts.nsec = (uint32_t)(100000000u + 100000u * i);
    }
    const kindr::minimal::QuatTransformation Trv(1, 0, 0, 0, 0.05, 0, 0);
    CHECK(fn(mapper, few, traj, Trv) == false);      // < 1024 events -> false, like the reference
    CHECK(fn(mapper, many, traj, Trv) == true);
    unsigned long long votes = 0;
    for (unsigned long long v : mapper.voteCounts()) votes += v;
    CHECK(votes > 0);
  }
  std::printf("ros adapters ok\n");
  return 0;
}
