// mapper_emvs_stereo.hpp — host-side mirror of EMVS::ShapeDSI / EMVS::MapperEMVS / LinearTrajectory
// (mapper_emvs_stereo/include/mapper_emvs_stereo/{mapper_emvs_stereo,trajectory,geometry_utils}.hpp)
// over the C-ABI of the B200 engine.  Same names, argument order and error behaviour for the
// mapping hot path:
//   MapperEMVS(cam, dsi_shape)                     mapper_emvs_stereo.cpp:29-64
//   bool evaluateDSI(events, trajectory, T_rv_w)   mapper_emvs_stereo.cpp:67-148  (false when < 1024 events)
//   getDepthMapFromDSI(...)                        mapper_emvs_stereo.cpp:344-375 + 302-313 (hot part, method = -1)
//   public members  dsi_ (Grid3D), name            mapper_emvs_stereo.hpp:116-117
// ROS / minkindr / OpenCV types are replaced by PODs of identical content:
//   dvs_msgs::Event                      -> emvs_event  {x, y, ts.sec, ts.nsec, polarity}
//   ros::Time                            -> geometry_utils::Time
//   kindr::minimal::QuatTransformation   -> geometry_utils::Transformation (unit quaternion w,x,y,z + position)
//   image_geometry::PinholeCameraModel   -> geometry_utils::CameraInfo (projection-matrix fx,fy,cx,cy,
//                                           resolution and the rectification LUT the reference derives
//                                           with OpenCV in precomputeRectifiedPoints, :256-299)
// Post-processing of the depth map (adaptive threshold, median, inpainting; :378-436) and
// getPointcloud stay with the caller's OpenCV / PCL code: they are outside the hot path.
#pragma once

#include "../cartesian3dgrid/cartesian3dgrid.h"

#include <algorithm>
#include <map>
#include <string>
#include <vector>

namespace geometry_utils {

struct Time {
  uint32_t sec = 0, nsec = 0;
  Time() {}
  Time(uint32_t s, uint32_t n) : sec(s), nsec(n) {}
  explicit Time(double t) : sec((uint32_t)t), nsec((uint32_t)((t - (uint32_t)t) * 1e9 + 0.5)) {}
  bool operator<(const Time& o) const { return sec != o.sec ? sec < o.sec : nsec < o.nsec; }
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
};

// kindr::minimal::QuatTransformation subset used on the mapping path
class Transformation {
 public:
  Transformation() { p_.q[0] = 1; p_.q[1] = p_.q[2] = p_.q[3] = 0; p_.t[0] = p_.t[1] = p_.t[2] = 0; }
  Transformation(const double q_wxyz[4], const double t[3])
  {
    for (int i = 0; i < 4; ++i) p_.q[i] = q_wxyz[i];
    for (int i = 0; i < 3; ++i) p_.t[i] = t[i];
  }
  explicit Transformation(const emvs_pose& p) : p_(p) {}
  Transformation operator*(const Transformation& rhs) const
  {
    Transformation out;
    emvs_host::check(emvs_pose_compose(&p_, &rhs.p_, &out.p_), "Transformation::operator*");
    return out;
  }
  Transformation inverse() const
  {
    Transformation out;
    emvs_host::check(emvs_pose_inverse(&p_, &out.p_), "Transformation::inverse");
    return out;
  }
  const double* getPosition() const { return p_.t; }
  const double* getRotationQuaternionWXYZ() const { return p_.q; }
  const emvs_pose& pod() const { return p_; }

 private:
  emvs_pose p_;
};

struct CameraInfo {
  uint32_t width = 0, height = 0;
  float fx = 0, fy = 0, cx = 0, cy = 0;   // of the PROJECTION matrix (mapper_emvs_stereo.cpp:46-48)
  std::vector<float> rectified_points;    // interleaved (x,y), index y*width + x; empty -> identity (no distortion)
};

}  // namespace geometry_utils

// trajectory.hpp:11-127
class LinearTrajectory {
 public:
  typedef std::map<geometry_utils::Time, geometry_utils::Transformation> PoseMap;
  LinearTrajectory() {}
  explicit LinearTrajectory(const PoseMap& poses)
  {
    if (poses.size() < 2u) throw std::runtime_error("At least two poses need to be provided");  // trajectory.hpp:89
    for (const auto& kv : poses) {
      emvs_stamped_pose sp;
      sp.sec = kv.first.sec; sp.nsec = kv.first.nsec; sp.T = kv.second.pod();
      poses_.push_back(sp);
    }
  }
  // false when t is outside the control poses: no extrapolation (trajectory.hpp:99-112)
  bool getPoseAt(const geometry_utils::Time& t, geometry_utils::Transformation& T) const
  {
    emvs_pose out;
    int found = 0;
    emvs_host::check(emvs_trajectory_pose_at(poses_.data(), poses_.size(), t.sec, t.nsec, &out, &found), "getPoseAt");
    if (found) T = geometry_utils::Transformation(out);
    return found != 0;
  }
  size_t getNumControlPoses() const { return poses_.size(); }
  const std::vector<emvs_stamped_pose>& pods() const { return poses_; }

 private:
  std::vector<emvs_stamped_pose> poses_;
};

namespace EMVS {

struct ShapeDSI {
  ShapeDSI() {}
  ShapeDSI(size_t dimX, size_t dimY, size_t dimZ, float min_depth, float max_depth, float fov)
      : dimX_(dimX), dimY_(dimY), dimZ_(dimZ), min_depth_(min_depth), max_depth_(max_depth), fov_(fov) {}
  size_t dimX_ = 0, dimY_ = 0, dimZ_ = 0;
  float min_depth_ = 0, max_depth_ = 0;
  float fov_ = 0;
};

// mapper_emvs_stereo.hpp:68-82 (the members getDepthMapFromDSI reads; the rest is I/O bookkeeping of main.cpp)
struct OptionsDepthMap {
  int adaptive_threshold_kernel_size_ = 5;
  double adaptive_threshold_c_ = 5.;
  double max_confidence = 0.;
  bool full_sequence = false;
  bool save_conf_stats = false;
  bool save_mono = false;
  bool save_dsi = false;
  double rv_pos = 0.;
  int median_filter_size_ = 5;
};

typedef LinearTrajectory TrajectoryType;

class MapperEMVS {
 public:
  MapperEMVS() {}
  MapperEMVS(const MapperEMVS&) = delete;
  MapperEMVS& operator=(const MapperEMVS&) { return *this; }  // a no-op in the reference too (hpp:99)

  // USE_INVERSE_DEPTH is a compile-time switch in the reference (CMakeLists.txt:41-44)
#ifdef USE_INVERSE_DEPTH
  static constexpr int kInverseDepth = 1;
#else
  static constexpr int kInverseDepth = 0;
#endif

  MapperEMVS(const geometry_utils::CameraInfo& cam, const ShapeDSI& dsi_shape) : dsi_shape_(dsi_shape)
  {
    emvs_camera c{cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy};
    emvs_shape s{(uint32_t)dsi_shape.dimX_, (uint32_t)dsi_shape.dimY_, (uint32_t)dsi_shape.dimZ_, dsi_shape.min_depth_,
                 dsi_shape.max_depth_, dsi_shape.fov_, kInverseDepth};
    emvs_host::check(emvs_mapper_create(emvs_host::default_context(), &c, &s, &m_), "MapperEMVS");  // CHECKs of :210-211
    std::vector<float> lut = cam.rectified_points;
    if (lut.empty()) {
      lut.resize((size_t)cam.width * cam.height * 2);
      for (uint32_t y = 0; y < cam.height; ++y)
        for (uint32_t x = 0; x < cam.width; ++x) {
          lut[2 * ((size_t)y * cam.width + x)] = (float)x;
          lut[2 * ((size_t)y * cam.width + x) + 1] = (float)y;
        }
    }
    emvs_host::check(emvs_mapper_set_lut(m_, lut.data(), (size_t)cam.width * cam.height), "MapperEMVS LUT");
    emvs_shape resolved;
    emvs_host::check(emvs_mapper_shape(m_, &resolved, nullptr), "MapperEMVS shape");
    dsi_shape_.dimX_ = resolved.dimX;
    dsi_shape_.dimY_ = resolved.dimY;
    raw_depths_vec_.resize(resolved.dimZ);
    emvs_host::check(emvs_mapper_depths(m_, raw_depths_vec_.data()), "MapperEMVS depths");
    emvs_grid* g = nullptr;
    emvs_host::check(emvs_mapper_grid(m_, &g), "MapperEMVS grid");
    dsi_.attach(g);
  }
  ~MapperEMVS()
  {
    dsi_.deallocate();
    if (m_) emvs_mapper_destroy(m_);
  }

  bool evaluateDSI(const std::vector<emvs_event>& events, const TrajectoryType& trajectory,
                   const geometry_utils::Transformation& T_rv_w)
  {
    return evaluateDSI(events.data(), events.size(), trajectory, T_rv_w);
  }
  // Same on a borrowed range (process_2 / process_5 hand over sub-intervals of one list without copying them).
  bool evaluateDSI(const emvs_event* events, size_t n_events, const TrajectoryType& trajectory,
                   const geometry_utils::Transformation& T_rv_w)
  {
    dsi_.touch();
    const int rc = emvs_mapper_evaluate_dsi(m_, events, n_events, trajectory.pods().data(), trajectory.pods().size(),
                                            &T_rv_w.pod());
    if (rc == EMVS_ERR_TOO_FEW) {
      std::cerr << "Number of events ( " << n_events << ") < packet size (" << EMVS_PACKET_SIZE << ")" << std::endl;
      return false;  // mapper_emvs_stereo.cpp:71-75
    }
    emvs_host::check(rc, "evaluateDSI");
    return true;
  }

  // Structure-of-arrays event list (x, y, t stored separately, e.g. read from a DSEC / TUM-VIE HDF5 file): same DSI,
  // a quarter of the PCIe traffic.  No counterpart in the reference, whose only event container is the
  // std::vector<dvs_msgs::Event> above.
  bool evaluateDSI(const emvs_events_soa& events, const TrajectoryType& trajectory, const geometry_utils::Transformation& T_rv_w)
  {
    dsi_.touch();
    const int rc = emvs_mapper_evaluate_dsi_soa(m_, &events, trajectory.pods().data(), trajectory.pods().size(), &T_rv_w.pod(),
                                                EMVS_BUILD_RESET);
    if (rc == EMVS_ERR_TOO_FEW) {
      std::cerr << "Number of events ( " << events.n << ") < packet size (" << EMVS_PACKET_SIZE << ")" << std::endl;
      return false;
    }
    emvs_host::check(rc, "evaluateDSI");
    return true;
  }

  // Streaming callers (main.cpp's full_seq loop): announce the arguments of the NEXT evaluateDSI call on this mapper.
  // The event upload starts and the host packet stage runs now, under whatever the device is computing; the later
  // evaluateDSI with the same arguments only launches kernels.  `events` and `trajectory` must stay unchanged until then.
  // (No counterpart in the reference: it is what hides PCIe behind the previous window's votes.)
  void prefetchDSI(const emvs_event* events, size_t n_events, const TrajectoryType& trajectory,
                   const geometry_utils::Transformation& T_rv_w)
  {
    emvs_host::check(emvs_mapper_prefetch_dsi(m_, events, n_events, trajectory.pods().data(), trajectory.pods().size(),
                                              &T_rv_w.pod()),
                     "prefetchDSI");
  }
  void prefetchDSI(const std::vector<emvs_event>& events, const TrajectoryType& trajectory,
                   const geometry_utils::Transformation& T_rv_w)
  {
    prefetchDSI(events.data(), events.size(), trajectory, T_rv_w);
  }

  // Hot part of getDepthMapFromDSI (method = -1): collapseMaxZSlice + convertDepthIndicesToValues.
  // depth_map = raw_depths_vec_[argmax] for EVERY pixel; the caller's mask / median / inpainting
  // (mapper_emvs_stereo.cpp:378-436) runs on these three maps.
  void getDepthMapFromDSI(emvs_host::Image<float>& depth_map, emvs_host::Image<float>& confidence_map,
                          emvs_host::Image<uint8_t>& depth_cell_indices)
  {
    if (raw_depths_vec_.size() > 256) throw std::runtime_error("getDepthMapFromDSI: dimZ > 256 (main.cpp:156)");
    const int rows = (int)dsi_shape_.dimY_, cols = (int)dsi_shape_.dimX_;
    depth_map.create(rows, cols);
    confidence_map.create(rows, cols);
    depth_cell_indices.create(rows, cols);
    emvs_host::check(emvs_grid_collapse_max(dsi_.handle(), raw_depths_vec_.data(), confidence_map.data.data(),
                                            depth_cell_indices.data.data(), depth_map.data.data()),
                     "getDepthMapFromDSI");
  }

  // getDepthMapFromDSI(depth_map, confidence_map, mask, options_depth_map) for method = -1
  // (mapper_emvs_stereo.cpp:339-436) without the Telea inpainting: the depth map is the depth of the
  // median-filtered indices, the mask is the adaptive-threshold mask with its border removed, and
  // confidence_map(0,0) is overwritten with max_confidence exactly like the reference leaves it.
  // The dense (inpainted) map stays with the caller's cv::inpaint on depth_cell_indices_filtered.
  void getDepthMapFromDSI(emvs_host::Image<float>& depth_map, emvs_host::Image<float>& confidence_map,
                          emvs_host::Image<uint8_t>& mask, const OptionsDepthMap& options_depth_map,
                          emvs_host::Image<uint8_t>* depth_cell_indices_filtered = nullptr)
  {
    const int rows = (int)dsi_shape_.dimY_, cols = (int)dsi_shape_.dimX_;
    depth_map.create(rows, cols);
    confidence_map.create(rows, cols);
    mask.create(rows, cols);
    if (depth_cell_indices_filtered) depth_cell_indices_filtered->create(rows, cols);
    emvs_depthmap_options o{options_depth_map.adaptive_threshold_kernel_size_, options_depth_map.adaptive_threshold_c_,
                            options_depth_map.max_confidence, options_depth_map.median_filter_size_};
    emvs_grid* g = dsi_.handle();
    emvs_host::check(emvs_depth_map_from_dsi(&g, 1, EMVS_FUSE_MAX, raw_depths_vec_.data(), &o, depth_map.data.data(),
                                             confidence_map.data.data(), mask.data.data(),
                                             depth_cell_indices_filtered ? depth_cell_indices_filtered->data.data() : nullptr),
                     "getDepthMapFromDSI");
  }

  // build-defined integer observable: accepted votes per plane of the last evaluateDSI
  std::vector<uint64_t> voteCounts() const
  {
    std::vector<uint64_t> c(raw_depths_vec_.size());
    emvs_host::check(emvs_mapper_counts(m_, c.data()), "voteCounts");
    return c;
  }

  const std::vector<float>& depths() const { return raw_depths_vec_; }
  emvs_mapper* handle() const { return m_; }

  Grid3D dsi_;
  std::string name;

 private:
  emvs_mapper* m_ = nullptr;
  ShapeDSI dsi_shape_;
  std::vector<float> raw_depths_vec_;
};

}  // namespace EMVS
