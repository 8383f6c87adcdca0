// example_process1.cpp — a caller written against the reference's class API (MapperEMVS, Grid3D,
// LinearTrajectory, process_1) running on the B200 engine: ESIM-like stereo rig, synthetic events
// of a fronto-parallel plane at 2.5 m seen from a translating camera (SURVEY.md §4 known-answer
// test: the argmax of every voted pixel must be the plane that contains 2.5 m).
//   make -C dvs_mcemvs_b200/host && dvs_mcemvs_b200/host/example_process1
#include "mapper_emvs_stereo/process1.hpp"

#include <cmath>
#include <cstdio>
#include <random>

int main()
{
  const uint32_t W = 240, H = 180;
  const float f = 200.f, cx = 120.f, cy = 90.f, baseline = 0.2f, Zstar = 2.5f;
  geometry_utils::CameraInfo cam;
  cam.width = W; cam.height = H; cam.fx = cam.fy = f; cam.cx = cx; cam.cy = cy;
  EMVS::ShapeDSI shape(0, 0, 64, 1.0f, 5.0f, 0.0f);
  try {
    EMVS::MapperEMVS mapper_fused(cam, shape), mapper0(cam, shape), mapper1(cam, shape);
    mapper_fused.name = "fused"; mapper0.name = "0"; mapper1.name = "1";

    // trajectory: camera 0 translates 0.2 m along x in 0.2 s; camera 1 sits `baseline` to its right
    LinearTrajectory::PoseMap poses0, poses1;
    const double q[4] = {1, 0, 0, 0};
    for (int i = 0; i <= 10; ++i) {
      const double t = 1000.0 + 0.02 * i - 0.01, x = 1.0 * (0.02 * i - 0.01);
      const double p0[3] = {x, 0, 0}, p1[3] = {x + baseline, 0, 0};
      poses0[geometry_utils::Time(t)] = geometry_utils::Transformation(q, p0);
      poses1[geometry_utils::Time(t)] = geometry_utils::Transformation(q, p1);
    }
    LinearTrajectory traj0(poses0), traj1(poses1);
    geometry_utils::Transformation T_w_rv;
    traj0.getPoseAt(geometry_utils::Time(1000.09), T_w_rv);
    const geometry_utils::Transformation T_rv_w = T_w_rv.inverse();

    // scene: 2000 fixed points on the plane Z = Zstar (world frame == camera-0 frame at x = 0);
    // every event is one of them seen from the camera position at the event's time
    std::mt19937 rng(1);
    std::uniform_real_distribution<float> ux(-1.2f, 1.6f), uy(-0.9f, 0.9f);
    std::vector<float> px(2000), py(2000);
    for (size_t i = 0; i < px.size(); ++i) { px[i] = ux(rng); py[i] = uy(rng); }
    std::uniform_int_distribution<size_t> pick(0, px.size() - 1);
    std::vector<emvs_event> ev[2];
    const size_t n_ev = 50000;
    for (int c = 0; c < 2; ++c)
      for (size_t i = 0; i < n_ev; ++i) {
        const double t = 1000.0 + 0.18 * (double)i / n_ev;
        const double camx = 1.0 * (t - 1000.0) + (c ? baseline : 0.0);
        for (;;) {
          const size_t k = pick(rng);
          const float X = px[k], Y = py[k];
          const float u = f * (X - (float)camx) / Zstar + cx, v = f * Y / Zstar + cy;
          const long xi = std::lround(u), yi = std::lround(v);
          if (xi < 0 || yi < 0 || xi >= (long)W || yi >= (long)H) continue;
          emvs_event e{};
          e.x = (uint16_t)xi; e.y = (uint16_t)yi;
          e.sec = (uint32_t)t; e.nsec = (uint32_t)((t - (uint32_t)t) * 1e9);
          ev[c].push_back(e);
          break;
        }
      }

    emvs_host::Image<float> depth, conf;
    emvs_host::Image<uint8_t> idx;
    Process1Timing tm;
    if (!process_1(mapper_fused, mapper0, mapper1, nullptr, ev[0], ev[1], nullptr, traj0, traj1, nullptr, T_rv_w,
                   /*stereo_fusion=*/2, depth, conf, idx, true, &tm))
      return 2;
    std::printf("Mean square = %g (fused), %g (cam0)\n", mapper_fused.dsi_.computeMeanSquare(), mapper0.dsi_.computeMeanSquare());
    std::printf("evaluateDSI: %.2f ms + %.2f ms (%.1f Mev/s), fusion+argmax %.2f ms\n", tm.build_ms[0], tm.build_ms[1],
                2.0 * n_ev / (1e3 * (tm.build_ms[0] + tm.build_ms[1])), tm.fuse_ms);
    // known answer: confident pixels sit on the plane containing Z*
    const std::vector<float>& z = mapper0.depths();
    size_t good = 0, total = 0;
    for (int y = 10; y < (int)H - 10; ++y)
      for (int x = 10; x < (int)W - 10; ++x)
        if (conf.at(y, x) > 15.f) {
          ++total;
          if (std::fabs(z[idx.at(y, x)] - Zstar) <= 2 * (5.0f - 1.0f) / 64 + 1e-4f) ++good;
        }
    std::printf("confident pixels: %zu, on the Z* = %.2f m plane (+-2 cells: events are rounded to integer pixels): %zu\n", total, Zstar, good);
    bool ok = total > 500 && good >= total * 90 / 100;

    // --- getDepthMapFromDSI with the reference's options (adaptive threshold + masked median + border) ---
    EMVS::OptionsDepthMap opts;          // defaults of main.cpp: kernel 5, C 5, median 5, max_confidence 0
    emvs_host::Image<float> depth_sd, conf_sd;
    emvs_host::Image<uint8_t> mask, idx_f;
    mapper_fused.getDepthMapFromDSI(depth_sd, conf_sd, mask, opts, &idx_f);
    size_t n_mask = 0, n_mask_on_plane = 0;
    for (int y = 0; y < (int)H; ++y)
      for (int x = 0; x < (int)W; ++x)
        if (mask.at(y, x)) {
          ++n_mask;
          if (std::fabs(depth_sd.at(y, x) - Zstar) <= 3 * (5.0f - 1.0f) / 64 + 1e-4f) ++n_mask_on_plane;
        }
    std::printf("semi-dense mask: %zu pixels, %zu within 3 cells of Z*\n", n_mask, n_mask_on_plane);
    ok = ok && n_mask > 500 && n_mask_on_plane >= n_mask * 50 / 100 && mask.at(0, 0) == 0 && mask.at(2, 2) == 0;

    // --- Grid3D voxel ops and the .npy writer through the mirror ---
    Grid3D a(4, 3, 2), b(4, 3, 2);
    std::vector<float> va(24), vb(24);
    for (int i = 0; i < 24; ++i) { va[i] = (float)i; vb[i] = (float)(24 - i); }
    a.upload(va.data());
    b.upload(vb.data());
    a.harmonicMeanTwoGrids(b);                       // 2ab / (a + b + 0.1)
    const float want = 2.f * (5.f * 19.f) / (5.f + 19.f + 0.1f);
    ok = ok && a.getGridValueAt(1, 1, 0) == want;    // p = 1 + 4*(1 + 3*0) = 5
    a.maxTwoGrids(b);
    ok = ok && a.getGridValueAt(0) == 24.f && std::fabs(a.computeMeanSquare() - b.computeMeanSquare()) < 400.0;
    const char* npy = "/tmp/emvs_example_dsi.npy";
    ok = ok && mapper0.dsi_.writeGridNpy(npy) == 0;
    if (FILE* f = std::fopen(npy, "rb")) {
      std::fseek(f, 0, SEEK_END);
      const long sz = std::ftell(f);
      std::fclose(f);
      std::printf("wrote %s: %ld bytes\n", npy, sz);
      ok = ok && sz == 128 + (long)W * H * 64 * 4;
    } else ok = false;
    // --- structure-of-arrays event list (x, y, t stored separately): the same DSI as the dvs_msgs::Event list ---
    {
      const std::vector<uint64_t> counts_aos = mapper0.voteCounts();
      const double ms_aos = mapper0.dsi_.computeMeanSquare();
      std::vector<uint16_t> sx(ev[0].size()), sy(ev[0].size());
      std::vector<int64_t> st(ev[0].size());
      for (size_t i = 0; i < ev[0].size(); ++i) {
        sx[i] = ev[0][i].x; sy[i] = ev[0][i].y;
        st[i] = (int64_t)ev[0][i].sec * 1000000000ll + ev[0][i].nsec;
      }
      const emvs_events_soa soa{sx.data(), sy.data(), st.data(), sx.size()};
      ok = ok && mapper0.evaluateDSI(soa, traj0, T_rv_w);
      ok = ok && mapper0.voteCounts() == counts_aos;                                    // integer observable: bit-exact
      ok = ok && std::fabs(mapper0.dsi_.computeMeanSquare() - ms_aos) <= 1e-6 * ms_aos;  // float sums: atomic order only
      std::printf("SoA evaluateDSI: counts %s, mean square %g vs %g\n", mapper0.voteCounts() == counts_aos ? "equal" : "DIFFER",
                  mapper0.dsi_.computeMeanSquare(), ms_aos);
    }
    // --- assignment to a mapper's public dsi_ overwrites the volume the mapper keeps voting into (the reference's
    //     plain member semantics, mapper_emvs_stereo.hpp:116), it does not detach the view ---
    {
      Grid3D copy0(mapper0.dsi_);                      // deep copy, like the reference's by-value arguments
      mapper1.dsi_ = copy0;                            // same dimensions: copies INTO mapper1's volume
      ok = ok && std::fabs(mapper1.dsi_.computeMeanSquare() - mapper0.dsi_.computeMeanSquare()) <= 1e-12 * mapper0.dsi_.computeMeanSquare();
      ok = ok && mapper1.evaluateDSI(ev[1], traj1, T_rv_w);          // ... and the mapper still builds into THAT volume
      emvs_host::Image<float> d1, c1;
      emvs_host::Image<uint8_t> i1;
      mapper1.getDepthMapFromDSI(d1, c1, i1);
      double s = 0;
      for (float v : c1.data) s += v;
      ok = ok && s > 0 && std::fabs(mapper1.dsi_.computeMeanSquare() - mapper0.dsi_.computeMeanSquare()) > 1e-9;
      bool threw = false;
      try { Grid3D small(4, 3, 2); mapper1.dsi_ = small; } catch (const std::runtime_error&) { threw = true; }
      ok = ok && threw;                                // different dimensions: an error, never a silently detached view
    }
    std::printf(ok ? "example_process1 ok\n" : "example_process1 FAILED\n");
    return ok ? 0 : 1;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 3;
  }
}
