// process1.hpp — compute steps of process_1 (Alg. 1: fusion across cameras),
// mapper_emvs_stereo/src/process1.cpp:28-224, on the B200 engine.  File output (saveDepthMaps,
// writeGridNpy) and the OpenCV post-processing stay with the caller.
#pragma once

#include "mapper_emvs_stereo.hpp"

#include <chrono>

struct Process1Timing {
  double build_ms[3] = {0, 0, 0};  // per camera, the scope of process1.cpp:72-81
  double fuse_ms = 0;              // process1.cpp:132-166 (here: fused with the argmax sweep)
};

// stereo_fusion ids as documented in docs/running.md: 1 min, 2 HM, 3 GM, 4 AM, 5 RMS, 6 max.
// Returns false on "Improper fusion method selected" (process1.cpp:155-157).
// With materialise_fused the fused DSI is written to mapper_fused.dsi_ exactly like the reference
// (needed for --save_dsi); otherwise fusion and argmax run as one sweep and only the maps exist.
inline bool process_1(EMVS::MapperEMVS& mapper_fused, EMVS::MapperEMVS& mapper0, EMVS::MapperEMVS& mapper1,
                      EMVS::MapperEMVS* mapper2, const std::vector<emvs_event>& events0,
                      const std::vector<emvs_event>& events1, const std::vector<emvs_event>* events2,
                      const LinearTrajectory& trajectory0, const LinearTrajectory& trajectory1,
                      const LinearTrajectory* trajectory2, const geometry_utils::Transformation& T_rv_w,
                      int stereo_fusion, emvs_host::Image<float>& depth_map, emvs_host::Image<float>& confidence_map,
                      emvs_host::Image<uint8_t>& depth_cell_indices, bool materialise_fused = true,
                      Process1Timing* timing = nullptr)
{
  using clock = std::chrono::high_resolution_clock;
  auto ms = [](clock::time_point a, clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  if (stereo_fusion < 1 || stereo_fusion > 6) {
    std::cerr << "Improper fusion method selected" << std::endl;
    return false;
  }
  EMVS::MapperEMVS* mappers[3] = {&mapper0, &mapper1, mapper2};
  const std::vector<emvs_event>* events[3] = {&events0, &events1, events2};
  const LinearTrajectory* trajs[3] = {&trajectory0, &trajectory1, trajectory2};
  int n = mapper2 ? 3 : 2;
  for (int i = 0; i < n; ++i) {
    const auto t0 = clock::now();
    mappers[i]->evaluateDSI(*events[i], *trajs[i], T_rv_w);
    if (timing) {  // evaluateDSI returns once its inputs are uploaded: wait for the votes before reading the clock
      emvs_host::check(emvs_context_sync(emvs_host::default_context()), "sync");
      timing->build_ms[i] = ms(t0, clock::now());
    }
  }
  // the reference ignores the third camera for GM / AM / RMS (process1.cpp:178-183)
  if (n == 3 && (stereo_fusion == 3 || stereo_fusion == 4 || stereo_fusion == 5)) n = 2;
  emvs_grid* grids[3];
  for (int i = 0; i < n; ++i) grids[i] = mappers[i]->dsi_.handle();
  const auto t0 = clock::now();
  int dimX, dimY, dimZ;
  mapper0.dsi_.getDimensions(&dimX, &dimY, &dimZ);
  depth_map.create(dimY, dimX);
  confidence_map.create(dimY, dimX);
  depth_cell_indices.create(dimY, dimX);
  if (materialise_fused) mapper_fused.dsi_.touch();
  emvs_host::check(emvs_fuse_collapse(grids, n, stereo_fusion, mapper_fused.depths().data(),
                                      materialise_fused ? mapper_fused.dsi_.handle() : nullptr, confidence_map.data.data(),
                                      depth_cell_indices.data.data(), depth_map.data.data()),
                   "process_1 fusion");
  if (timing) timing->fuse_ms = ms(t0, clock::now());
  return true;
}
