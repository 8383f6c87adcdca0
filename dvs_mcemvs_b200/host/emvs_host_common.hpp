// emvs_host_common.hpp — shared plumbing of the host-side C++ mirror: the process-wide default
// engine context, status -> exception mapping, and the POD stand-ins for the ROS / OpenCV /
// minkindr types of the reference's signatures (none of those libraries is required here; when a
// caller has them, the adapters at the bottom of mapper_emvs_stereo.hpp convert).
#pragma once

#include <emvs_b200.h>

#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace emvs_host {

// The reference's voxel accessors use std::vector::at() and may throw std::out_of_range; its
// setup uses glog CHECKs.  The mirror throws std::runtime_error carrying emvs_last_error().
inline void check(int status, const char* where)
{
  if (status != EMVS_OK) throw std::runtime_error(std::string(where) + ": " + emvs_last_error());
}

// One engine context per process (the reference is single-device, single-threaded).  The CUDA
// device is taken from $EMVS_DEVICE (default 0); under torchrun-style launchers set it to
// LOCAL_RANK before the first Grid3D / MapperEMVS is constructed.
inline emvs_context* default_context()
{
  static emvs_context* ctx = [] {
    const char* env = std::getenv("EMVS_DEVICE");
    emvs_context* c = nullptr;
    check(emvs_context_create(env ? std::atoi(env) : 0, &c), "emvs_context_create");
    return c;
  }();
  return ctx;
}

// cv::Mat stand-in for the three 2-D outputs of the hot path (row-major, rows x cols).
template <typename T>
struct Image {
  int rows = 0, cols = 0;
  std::vector<T> data;
  void create(int r, int c) { rows = r; cols = c; data.assign((size_t)r * c, T()); }
  T& at(int y, int x) { return data[(size_t)y * cols + x]; }
  const T& at(int y, int x) const { return data[(size_t)y * cols + x]; }
};

}  // namespace emvs_host
