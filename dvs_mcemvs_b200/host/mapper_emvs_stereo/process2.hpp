// process2.hpp — compute steps of process_2 (Alg. 2: fusion across cameras and time,
// mapper_emvs_stereo/src/process2.cpp:28-290) and of its shuffled variant process_5
// (process5.cpp:27-253; declared in process5.hpp) on the B200 engine.
//
// Same argument list as the reference.  What the reference writes to disk with saveDepthMaps / cv::imwrite (the PNG
// and TXT encoders are outside the mapping path) is handed back in memory instead: pass a Process2Outputs to receive
// every depth map under the suffix the reference would have used as its file name.  --save_dsi output (.npy) is
// written like the reference (process2.cpp:281-287).
//
// Semantics kept literally:
//   * sub-interval k of camera c is events[k*n_c, (k+1)*n_c) with n_c = unsigned(size_c) / num_subintervals — the
//     remainder of the list is never used, and left / right sub-intervals are NOT time-aligned (process2.cpp:46-47,
//     104-107, 134-137);
//   * temporal_fusion 2 = harmonic (sum of 1/(0.01+x), then n/sum), 4 = arithmetic (sum, then /n); every other id
//     accumulates nothing, so mapper_fused stays all-zero (the empty switch cases of process2.cpp:211-242);
//   * mapper_fused.dsi_ is reset, mapper_fused_camera_time.dsi_ is NOT (process2.cpp:92, 267: addTwoGrids onto
//     whatever the caller left in it);
//   * the time-then-camera switch maps stereo_fusion 3 -> arithmeticMean and 4 -> geometricMean, swapped relative
//     to every other switch (process2.cpp:274-279).  literal_camera_time_ids = true (default) reproduces that;
//     false uses the documented ids (3 = GM, 4 = AM).
#pragma once

#include "mapper_emvs_stereo.hpp"

#include <chrono>
#include <map>

struct Process2Maps {
  emvs_host::Image<float> depth_map, confidence_map;
  emvs_host::Image<uint8_t> semidense_mask;
};

struct Process2Outputs {
  // keyed by the reference's saveDepthMaps suffix: "0_000", "1_000", "fused_000", ..., "left_temporal_4",
  // "right_temporal_4", "stereo_temporal_4", "stereo_temporal_camera_time4"
  std::map<std::string, Process2Maps> maps;
  std::vector<double> mean_square[2];   // per sub-interval, left / right (the reference LOGs them, process2.cpp:120,150)
  double mean_square_fused = 0;         // process2.cpp:254
  double fusion_ms = 0;                 // "Time taken to fuse across space and time" (device-synchronised)
  geometry_utils::Transformation T_rv_w;
};

namespace emvs_host {
namespace detail {

inline std::string suffix3(const char* prefix, unsigned k)
{
  char buf[64];
  std::snprintf(buf, sizeof buf, "%s%03u", prefix, k);
  return buf;
}

inline void depth_maps_of(EMVS::MapperEMVS& m, const EMVS::OptionsDepthMap& o, Process2Outputs* out, const std::string& key)
{
  if (!out) return;
  Process2Maps& r = out->maps[key];
  m.getDepthMapFromDSI(r.depth_map, r.confidence_map, r.semidense_mask, o);
}

// false on "Improper fusion method selected"
inline bool stereo_fuse(Grid3D& a, const Grid3D& b, int id)
{
  switch (id) {
    case 1: a.minTwoGrids(b); return true;
    case 2: a.harmonicMeanTwoGrids(b); return true;
    case 3: a.geometricMeanTwoGrids(b); return true;
    case 4: a.arithmeticMeanTwoGrids(b); return true;
    case 5: a.rmsTwoGrids(b); return true;
    case 6: a.maxTwoGrids(b); return true;
    default: return false;
  }
}

// Shared body of process_2 (shift = 0, all outputs) and process_5 (shift = num_subintervals / 2, final map only).
inline bool process_2_impl(const geometry_utils::CameraInfo& cam0, const geometry_utils::CameraInfo& cam1,
                           const LinearTrajectory& trajectory0, const LinearTrajectory& trajectory1,
                           const std::vector<emvs_event>& events0, const std::vector<emvs_event>& events1,
                           const EMVS::OptionsDepthMap& opts_depth_map, const EMVS::ShapeDSI& dsi_shape,
                           const int num_subintervals, EMVS::MapperEMVS& mapper_fused,
                           EMVS::MapperEMVS* mapper_fused_camera_time, const std::string& out_path, double ts,
                           int stereo_fusion, int temporal_fusion, bool shuffle, bool literal_camera_time_ids,
                           Process2Outputs* out)
{
  using clock = std::chrono::high_resolution_clock;
  if (num_subintervals < 1) throw std::runtime_error("process_2: num_subintervals must be >= 1");
  const unsigned int n_sub = (unsigned int)num_subintervals;
  const unsigned int per[2] = {static_cast<unsigned int>(events0.size()) / n_sub,
                               static_cast<unsigned int>(events1.size()) / n_sub};

  EMVS::MapperEMVS mapper0(cam0, dsi_shape), mapper1(cam1, dsi_shape);
  EMVS::MapperEMVS mapper_fused_subinterval(cam0, dsi_shape), mapper_fused_left(cam0, dsi_shape),
      mapper_fused_right(cam0, dsi_shape);

  // reference view: pose of the left camera at ts (a miss leaves the identity, as in the reference)
  geometry_utils::Transformation T_w_rv;
  trajectory0.getPoseAt(geometry_utils::Time(ts), T_w_rv);
  const geometry_utils::Transformation T_rv_w = T_w_rv.inverse();
  if (out) out->T_rv_w = T_rv_w;

  auto sync = [] { emvs_host::check(emvs_context_sync(emvs_host::default_context()), "sync"); };
  auto ms = [](clock::time_point a, clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  double fusion_ms = 0;
  const bool per_interval_maps = !opts_depth_map.full_sequence && !shuffle;   // process_5 saves only the final map

  mapper_fused.dsi_.resetGrid();
  const unsigned int shift = shuffle ? n_sub / 2 : 0;
  unsigned int idx_first_ev[2] = {0, shift * per[1]};
  std::vector<emvs_event> wrapped;
  for (unsigned int k = 0; k < n_sub; ++k) {
    {  // left camera
      const emvs_event* sub = events0.data() + idx_first_ev[0];
      idx_first_ev[0] += per[0];
      if (!mapper0.evaluateDSI(sub, per[0], trajectory0, T_rv_w)) mapper0.dsi_.resetGrid();   // process2.cpp:100
      if (out) out->mean_square[0].push_back(mapper0.dsi_.computeMeanSquare());
      if (per_interval_maps) depth_maps_of(mapper0, opts_depth_map, out, suffix3("0_", k));
    }
    {  // right camera; process_5 starts it `shift` sub-intervals later and wraps around (process5.cpp:134-150)
      const emvs_event* sub = events1.data() + idx_first_ev[1];
      if (shuffle && (size_t)idx_first_ev[1] + per[1] >= events1.size()) {
        const size_t over = (size_t)idx_first_ev[1] + per[1] - events1.size();
        wrapped.assign(events1.begin() + idx_first_ev[1], events1.end());
        wrapped.insert(wrapped.end(), events1.begin(), events1.begin() + over);
        sub = wrapped.data();
        idx_first_ev[1] = (unsigned int)over;
        if (!mapper1.evaluateDSI(sub, wrapped.size(), trajectory1, T_rv_w)) mapper1.dsi_.resetGrid();
      } else {
        idx_first_ev[1] += per[1];
        if (!mapper1.evaluateDSI(sub, per[1], trajectory1, T_rv_w)) mapper1.dsi_.resetGrid();
      }
      if (out) out->mean_square[1].push_back(mapper1.dsi_.computeMeanSquare());
      if (per_interval_maps) depth_maps_of(mapper1, opts_depth_map, out, suffix3("1_", shuffle ? (k + shift) % n_sub : k));
    }

    // across cameras, one sub-interval (process2.cpp:158-194)
    mapper_fused_subinterval.dsi_.copyFrom(mapper0.dsi_);
    sync();
    auto t0 = clock::now();
    if (!stereo_fuse(mapper_fused_subinterval.dsi_, mapper1.dsi_, stereo_fusion)) {
      std::cerr << "Improper fusion method selected" << std::endl;
      return false;
    }
    sync();
    fusion_ms += ms(t0, clock::now());
    if (per_interval_maps) depth_maps_of(mapper_fused_subinterval, opts_depth_map, out, suffix3("fused_", k));

    // across time (process2.cpp:206-247)
    t0 = clock::now();
    if (temporal_fusion == 2) {
      mapper_fused_left.dsi_.addInverseOfTwoGrids(mapper0.dsi_);
      mapper_fused_right.dsi_.addInverseOfTwoGrids(mapper1.dsi_);
      mapper_fused.dsi_.addInverseOfTwoGrids(mapper_fused_subinterval.dsi_);
      if (k == n_sub - 1) {
        mapper_fused_left.dsi_.computeHMfromSumOfInv(num_subintervals);
        mapper_fused_right.dsi_.computeHMfromSumOfInv(num_subintervals);
        mapper_fused.dsi_.computeHMfromSumOfInv(num_subintervals);
      }
    } else if (temporal_fusion == 4) {
      mapper_fused_left.dsi_.addTwoGrids(mapper0.dsi_);
      mapper_fused_right.dsi_.addTwoGrids(mapper1.dsi_);
      mapper_fused.dsi_.addTwoGrids(mapper_fused_subinterval.dsi_);
      if (k == n_sub - 1) {
        mapper_fused_left.dsi_.computeAMfromSum(num_subintervals);
        mapper_fused_right.dsi_.computeAMfromSum(num_subintervals);
        mapper_fused.dsi_.computeAMfromSum(num_subintervals);
      }
    }
    sync();
    fusion_ms += ms(t0, clock::now()) / 3;   // the reference reports a third of the three accumulations
  }
  if (out) {
    out->fusion_ms = fusion_ms;
    out->mean_square_fused = mapper_fused.dsi_.computeMeanSquare();
  }

  const std::string tf = std::to_string(temporal_fusion);
  if (per_interval_maps) {
    depth_maps_of(mapper_fused_left, opts_depth_map, out, "left_temporal_" + tf);
    depth_maps_of(mapper_fused_right, opts_depth_map, out, "right_temporal_" + tf);
  }
  depth_maps_of(mapper_fused, opts_depth_map, out, "stereo_temporal_" + tf);
  if (!mapper_fused_camera_time) return true;   // process_5 ends here

  // time first, then cameras (process2.cpp:266-289)
  Grid3D& ct = mapper_fused_camera_time->dsi_;
  ct.addTwoGrids(mapper_fused_left.dsi_);
  int id = stereo_fusion;
  if (literal_camera_time_ids && (id == 3 || id == 4)) id = 7 - id;
  if (!stereo_fuse(ct, mapper_fused_right.dsi_, id)) {
    std::cerr << "Improper stereo fusion method selected" << std::endl;
    return false;
  }
  if (opts_depth_map.save_dsi) {
    mapper_fused_left.dsi_.writeGridNpy((out_path + "dsi_fused_0_temporalfusion.npy").c_str());
    mapper_fused_right.dsi_.writeGridNpy((out_path + "dsi_fused_1_temporalfusion.npy").c_str());
    mapper_fused.dsi_.writeGridNpy((out_path + "dsi_stereo_temporalfusion.npy").c_str());
    ct.writeGridNpy((out_path + "dsi_stereo_temporalfusion_camera_time.npy").c_str());
  }
  depth_maps_of(*mapper_fused_camera_time, opts_depth_map, out, "stereo_temporal_camera_time" + tf);
  return true;
}

}  // namespace detail
}  // namespace emvs_host

// Alg 2 (process2.cpp:28-290).  Returns false where the reference returns early on an improper fusion id.
inline bool process_2(const geometry_utils::CameraInfo& cam0, const geometry_utils::CameraInfo& cam1,
                      const LinearTrajectory& trajectory0, const LinearTrajectory& trajectory1,
                      const std::vector<emvs_event>& events0, const std::vector<emvs_event>& events1,
                      const EMVS::OptionsDepthMap& opts_depth_map, const EMVS::ShapeDSI& dsi_shape,
                      const int num_subintervals, EMVS::MapperEMVS& mapper_fused,
                      EMVS::MapperEMVS& mapper_fused_camera_time, const std::string& out_path, double ts,
                      int stereo_fusion, int temporal_fusion, Process2Outputs* out = nullptr,
                      bool literal_camera_time_ids = true)
{
  return emvs_host::detail::process_2_impl(cam0, cam1, trajectory0, trajectory1, events0, events1, opts_depth_map, dsi_shape,
                                           num_subintervals, mapper_fused, &mapper_fused_camera_time, out_path, ts,
                                           stereo_fusion, temporal_fusion, /*shuffle=*/false, literal_camera_time_ids, out);
}
