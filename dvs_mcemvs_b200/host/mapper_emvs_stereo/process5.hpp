// process5.hpp — process_5 (Alg. 2 + shuffling, mapper_emvs_stereo/src/process5.cpp:27-253): the right camera's
// sub-intervals start num_subintervals/2 later and wrap around the end of its event list, so mini-DSIs of different
// times are fused across cameras.  Shares its body with process_2 (process2.hpp); like the reference it produces only
// the final "stereo_temporal_<id>" depth map and has no time-then-camera output.
#pragma once

#include "process2.hpp"

inline bool process_5(const geometry_utils::CameraInfo& cam0, const geometry_utils::CameraInfo& cam1,
                      const LinearTrajectory& trajectory0, const LinearTrajectory& trajectory1,
                      const std::vector<emvs_event>& events0, const std::vector<emvs_event>& events1,
                      const EMVS::OptionsDepthMap& opts_depth_map, const EMVS::ShapeDSI& dsi_shape,
                      const int num_subintervals, EMVS::MapperEMVS& mapper_fused, const std::string& out_path, double ts,
                      int stereo_fusion, int temporal_fusion, Process2Outputs* out = nullptr)
{
  return emvs_host::detail::process_2_impl(cam0, cam1, trajectory0, trajectory1, events0, events1, opts_depth_map, dsi_shape,
                                           num_subintervals, mapper_fused, nullptr, out_path, ts, stereo_fusion,
                                           temporal_fusion, /*shuffle=*/true, false, out);
}
