// example_process2.cpp — a caller of process_2 / process_5 (Alg. 2 and its shuffled variant) written against the
// reference's argument lists, running on the B200 engine.  Known answers checked:
//   * arithmetic mean commutes: with stereo_fusion = AM and temporal_fusion = AM, "cameras then time"
//     (mapper_fused) equals "time then cameras" (mapper_fused_camera_time) up to float rounding;
//   * the reference's id swap in the time-then-camera switch (process2.cpp:274-279): literal ids with
//     stereo_fusion = 3 give the same time-then-camera volume as documented ids with stereo_fusion = 4;
//   * process_5 with one sub-interval has shift 0 and therefore equals process_2 with one sub-interval;
//   * an improper fusion id returns false; temporal ids other than 2 / 4 leave mapper_fused all-zero.
//   make -C dvs_mcemvs_b200/host && dvs_mcemvs_b200/host/example_process2
#include "mapper_emvs_stereo/process5.hpp"

#include <cmath>
#include <cstdio>
#include <random>

static double max_rel_diff(Grid3D& a, Grid3D& b)
{
  int X, Y, Z;
  a.getDimensions(&X, &Y, &Z);
  double worst = 0;
  for (int z = 0; z < Z; ++z) {
    const float* pa = a.getPointerToSlice(z);
    const float* pb = b.getPointerToSlice(z);
    for (int i = 0; i < X * Y; ++i) {
      const double d = std::fabs((double)pa[i] - (double)pb[i]) / (1e-3 + std::fabs((double)pb[i]));
      if (d > worst) worst = d;
    }
  }
  return worst;
}

int main()
{
  const uint32_t W = 240, H = 180;
  const float f = 200.f, cx = 120.f, cy = 90.f, baseline = 0.2f, Zstar = 2.5f;
  geometry_utils::CameraInfo cam;
  cam.width = W; cam.height = H; cam.fx = cam.fy = f; cam.cx = cx; cam.cy = cy;
  EMVS::ShapeDSI shape(0, 0, 32, 1.0f, 5.0f, 0.0f);
  try {
    LinearTrajectory::PoseMap poses0, poses1;
    const double q[4] = {1, 0, 0, 0};
    for (int i = 0; i <= 10; ++i) {
      const double t = 1000.0 + 0.02 * i - 0.01, x = 1.0 * (0.02 * i - 0.01);
      const double p0[3] = {x, 0, 0}, p1[3] = {x + baseline, 0, 0};
      poses0[geometry_utils::Time(t)] = geometry_utils::Transformation(q, p0);
      poses1[geometry_utils::Time(t)] = geometry_utils::Transformation(q, p1);
    }
    LinearTrajectory traj0(poses0), traj1(poses1);

    std::mt19937 rng(2);
    std::uniform_real_distribution<float> ux(-1.2f, 1.6f), uy(-0.9f, 0.9f);
    std::vector<float> px(2000), py(2000);
    for (size_t i = 0; i < px.size(); ++i) { px[i] = ux(rng); py[i] = uy(rng); }
    std::uniform_int_distribution<size_t> pick(0, px.size() - 1);
    std::vector<emvs_event> ev[2];
    const size_t n_ev[2] = {40003, 36007};   // not multiples of the sub-interval count: the remainders are dropped
    for (int c = 0; c < 2; ++c)
      for (size_t i = 0; i < n_ev[c]; ++i) {
        const double t = 1000.0 + 0.18 * (double)i / n_ev[c];
        const double camx = 1.0 * (t - 1000.0) + (c ? baseline : 0.0);
        for (;;) {
          const size_t k = pick(rng);
          const float u = f * (px[k] - (float)camx) / Zstar + cx, v = f * py[k] / Zstar + cy;
          const long xi = std::lround(u), yi = std::lround(v);
          if (xi < 0 || yi < 0 || xi >= (long)W || yi >= (long)H) continue;
          emvs_event e{};
          e.x = (uint16_t)xi; e.y = (uint16_t)yi;
          e.sec = (uint32_t)t; e.nsec = (uint32_t)((t - (uint32_t)t) * 1e9);
          ev[c].push_back(e);
          break;
        }
      }

    EMVS::OptionsDepthMap opts;
    const double ts = 1000.09;
    bool ok = true;

    // --- AM / AM: cameras-then-time == time-then-cameras ---------------------------------------------------------
    EMVS::MapperEMVS fused_a(cam, shape), ct_a(cam, shape);
    Process2Outputs out_a;
    ok = ok && process_2(cam, cam, traj0, traj1, ev[0], ev[1], opts, shape, 4, fused_a, ct_a, "/tmp/", ts, 4, 4, &out_a,
                         /*literal_camera_time_ids=*/false);
    const double d_commute = max_rel_diff(fused_a.dsi_, ct_a.dsi_);
    std::printf("AM/AM: mean square fused %.6g, |cameras-then-time - time-then-cameras| max rel %.3g, fusion %.2f ms\n",
                out_a.mean_square_fused, d_commute, out_a.fusion_ms);
    ok = ok && out_a.mean_square_fused > 0 && d_commute < 1e-4;
    // every map the reference would have saved is there (4 x {0_, 1_, fused_} + left/right/stereo temporal + camera_time)
    ok = ok && out_a.maps.size() == 16 && out_a.maps.count("0_003") && out_a.maps.count("fused_000") &&
         out_a.maps.count("left_temporal_4") && out_a.maps.count("stereo_temporal_4") &&
         out_a.maps.count("stereo_temporal_camera_time4");
    ok = ok && out_a.mean_square[0].size() == 4 && out_a.mean_square[1].size() == 4 && out_a.mean_square[1][3] > 0;
    {  // the final stereo map sits on the Z* plane where it is confident
      const Process2Maps& m = out_a.maps["stereo_temporal_4"];
      size_t n_mask = 0, good = 0;
      for (int y = 0; y < (int)H; ++y)
        for (int x = 0; x < (int)W; ++x)
          if (m.semidense_mask.at(y, x)) {
            ++n_mask;
            if (std::fabs(m.depth_map.at(y, x) - Zstar) <= 3 * (5.0f - 1.0f) / 32 + 1e-4f) ++good;
          }
      std::printf("stereo_temporal_4: %zu mask pixels, %zu within 3 cells of Z* = %.2f m\n", n_mask, good, Zstar);
      ok = ok && n_mask > 300 && good >= n_mask / 2;
    }

    // --- the literal id swap: stereo 3 with literal ids fuses time-then-camera by AM ------------------------------
    EMVS::MapperEMVS fused_b(cam, shape), ct_b(cam, shape);
    ok = ok && process_2(cam, cam, traj0, traj1, ev[0], ev[1], opts, shape, 4, fused_b, ct_b, "/tmp/", ts, 3, 4);
    const double d_swap = max_rel_diff(ct_b.dsi_, ct_a.dsi_);
    std::printf("literal ids, stereo 3: time-then-cameras vs AM max rel %.3g\n", d_swap);
    ok = ok && d_swap < 1e-4 && max_rel_diff(fused_b.dsi_, fused_a.dsi_) > 1e-2;   // cameras-then-time really used GM

    // --- process_5 with one sub-interval == process_2 with one sub-interval --------------------------------------
    EMVS::MapperEMVS fused_c(cam, shape), ct_c(cam, shape), fused_d(cam, shape);
    opts.full_sequence = true;
    Process2Outputs out_c, out_d;
    ok = ok && process_2(cam, cam, traj0, traj1, ev[0], ev[1], opts, shape, 1, fused_c, ct_c, "/tmp/", ts, 2, 2, &out_c);
    ok = ok && process_5(cam, cam, traj0, traj1, ev[0], ev[1], opts, shape, 1, fused_d, "/tmp/", ts, 2, 2, &out_d);
    const double d_15 = max_rel_diff(fused_d.dsi_, fused_c.dsi_);
    std::printf("process_5 vs process_2, one sub-interval, HM/HM: max rel %.3g; maps %zu / %zu\n", d_15, out_c.maps.size(),
                out_d.maps.size());
    ok = ok && d_15 < 1e-4 && out_c.maps.size() == 2 && out_d.maps.size() == 1 && out_d.maps.count("stereo_temporal_2");

    // --- process_5 with shuffling runs and differs from the unshuffled fusion ------------------------------------
    EMVS::MapperEMVS fused_e(cam, shape), fused_f(cam, shape), ct_f(cam, shape);
    ok = ok && process_5(cam, cam, traj0, traj1, ev[0], ev[1], opts, shape, 4, fused_e, "/tmp/", ts, 2, 4);
    ok = ok && process_2(cam, cam, traj0, traj1, ev[0], ev[1], opts, shape, 4, fused_f, ct_f, "/tmp/", ts, 2, 4);
    const double ms_e = fused_e.dsi_.computeMeanSquare(), ms_f = fused_f.dsi_.computeMeanSquare();
    std::printf("HM stereo / AM time: mean square shuffled %.6g, unshuffled %.6g\n", ms_e, ms_f);
    ok = ok && ms_e > 0 && ms_f > 0 && std::fabs(ms_e - ms_f) > 1e-9 * ms_f;

    // --- error behaviour ------------------------------------------------------------------------------------------
    EMVS::MapperEMVS fused_g(cam, shape), ct_g(cam, shape);
    ok = ok && !process_2(cam, cam, traj0, traj1, ev[0], ev[1], opts, shape, 2, fused_g, ct_g, "/tmp/", ts, 9, 4);
    ok = ok && process_2(cam, cam, traj0, traj1, ev[0], ev[1], opts, shape, 2, fused_g, ct_g, "/tmp/", ts, 2, 1);
    ok = ok && fused_g.dsi_.computeMeanSquare() == 0.0;   // temporal id 1 accumulates nothing

    std::printf(ok ? "example_process2 ok\n" : "example_process2 FAILED\n");
    return ok ? 0 : 1;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 3;
  }
}
