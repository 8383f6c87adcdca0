// ref_glue.cpp — C entry points over the REFERENCE's own Grid3D (cartesian3dgrid.h / .cpp compiled
// in place from /root/reference, never copied), used to pin the restated oracle and to generate
// tests/golden/.  TEST INFRASTRUCTURE ONLY.  Only the cv::Mat container is a stand-in
// (oracle/shim/emvs_cv_shim.h); every float operation executed here is the reference's.
#include <cartesian3dgrid/cartesian3dgrid.h>
#include <mapper_emvs_stereo/depth_vector.hpp>
#include <mapper_emvs_stereo/median_filtering.hpp>

#include <chrono>
#include <cstdint>
#include <cstring>

namespace {
void load(Grid3D& g, const float* src, size_t n) { std::memcpy(g.getPointerToSlice(0), src, n * sizeof(float)); }
void store(Grid3D& g, float* dst, size_t n) { std::memcpy(dst, g.getPointerToSlice(0), n * sizeof(float)); }
}  // namespace

extern "C" {

// Grid3D::accumulateGridValueAt(x_f, y_f, grid) for n points on slice k of a (dimX,dimY,dimZ) grid
// whose contents are vol (in/out).  cartesian3dgrid.h:253-273
void ref_vote(uint32_t dimX, uint32_t dimY, uint32_t dimZ, float* vol, uint32_t k, const float* x, const float* y, uint64_t n)
{
  Grid3D g(dimX, dimY, dimZ);
  const size_t cells = (size_t)dimX * dimY * dimZ;
  load(g, vol, cells);
  float* slice = g.getPointerToSlice((int)k);
  for (uint64_t i = 0; i < n; ++i) g.accumulateGridValueAt(x[i], y[i], slice);
  store(g, vol, cells);
}

// Voxel-wise ops, ids as in include/emvs_b200.h EMVS_OP_*.  cartesian3dgrid.h:64-192
void ref_grid_op(int op, uint32_t dimX, uint32_t dimY, uint32_t dimZ, float* a, const float* b, int n, float eps)
{
  Grid3D ga(dimX, dimY, dimZ), gb(dimX, dimY, dimZ);
  const size_t cells = (size_t)dimX * dimY * dimZ;
  load(ga, a, cells);
  if (b) load(gb, b, cells);
  switch (op) {
    case 0: ga.addTwoGrids(gb); break;
    case 1: ga.minTwoGrids(gb); break;
    case 2: ga.harmonicMeanTwoGrids(gb, eps); break;
    case 3: ga.geometricMeanTwoGrids(gb); break;
    case 4: ga.arithmeticMeanTwoGrids(gb); break;
    case 5: ga.rmsTwoGrids(gb); break;
    case 6: ga.maxTwoGrids(gb); break;
    case 7: ga.harmonicMeanTwoGrids(gb, n, eps); break;
    case 8: ga.addInverseOfTwoGrids(gb, eps); break;
    case 9: ga.computeHMfromSumOfInv(n); break;
    case 10: ga.computeAMfromSum(n); break;
    default: break;
  }
  store(ga, a, cells);
}

// Grid3D::collapseMaxZSlice.  cartesian3dgrid.cpp:115-137 (index is uchar: dimZ <= 256)
void ref_collapse_max(uint32_t dimX, uint32_t dimY, uint32_t dimZ, const float* vol, float* conf, uint8_t* idx)
{
  Grid3D g(dimX, dimY, dimZ);
  load(g, vol, (size_t)dimX * dimY * dimZ);
  cv::Mat mv, mi;
  g.collapseMaxZSlice(&mv, &mi);
  std::memcpy(conf, mv.data, (size_t)dimX * dimY * sizeof(float));
  std::memcpy(idx, mi.data, (size_t)dimX * dimY);
}

// raw_depths_vec_ as MapperEMVS::setupDSI builds it (mapper_emvs_stereo.cpp:213-214) from the
// reference's own depth_vector.hpp: LinearDepthVector (:76-114) or InverseDepthVector (:119-163).
void ref_depth_vector(int inverse, float zmin, float zmax, uint64_t nz, float* out)
{
  std::vector<float> v;
  if (inverse) v = EMVS::InverseDepthVector(zmin, zmax, (size_t)nz).getDepthVector();
  else v = EMVS::LinearDepthVector(zmin, zmax, (size_t)nz).getDepthVector();
  std::memcpy(out, v.data(), v.size() * sizeof(float));
}

// huangMedianFilter(img, out, mask, patch_size) — mapper_emvs_stereo/src/median_filtering.cpp:33-158
// (the masked median that cleans the depth-cell indices, mapper_emvs_stereo.cpp:418-423).
void ref_huang_median(const uint8_t* img, const uint8_t* mask, int rows, int cols, int patch_size, uint8_t* out)
{
  cv::Mat mi(rows, cols, CV_8U, (void*)img), mm(rows, cols, CV_8U, (void*)mask), mo;
  huangMedianFilter(mi, mo, mm, patch_size);
  std::memcpy(out, mo.data, (size_t)rows * cols);
}

// Times the reference's OWN fusion + argmax code on two volumes, in the order process_1 issues it
// (process1.cpp:126-166: fused.resetGrid(); fused.addTwoGrids(dsi0); fused.<op>TwoGrids(dsi1)) followed by
// collapseMaxZSlice (mapper_emvs_stereo.cpp:367-369).  The by-value argument copies and .at() bounds checks
// are the reference's.  method: stereo_fusion id 1..6.  Loading the inputs is outside the timed scopes.
int ref_time_fuse_collapse(uint32_t dimX, uint32_t dimY, uint32_t dimZ, const float* a, const float* b, int method,
                           double* fuse_ms, double* argmax_ms, float* conf, uint8_t* idx)
{
  using clock = std::chrono::high_resolution_clock;
  Grid3D g0(dimX, dimY, dimZ), g1(dimX, dimY, dimZ), fused(dimX, dimY, dimZ);
  const size_t cells = (size_t)dimX * dimY * dimZ;
  load(g0, a, cells);
  load(g1, b, cells);
  auto t0 = clock::now();
  fused.resetGrid();
  fused.addTwoGrids(g0);
  switch (method) {
    case 1: fused.minTwoGrids(g1); break;
    case 2: fused.harmonicMeanTwoGrids(g1); break;
    case 3: fused.geometricMeanTwoGrids(g1); break;
    case 4: fused.arithmeticMeanTwoGrids(g1); break;
    case 5: fused.rmsTwoGrids(g1); break;
    case 6: fused.maxTwoGrids(g1); break;
    default: return 1;
  }
  auto t1 = clock::now();
  cv::Mat mv, mi;
  fused.collapseMaxZSlice(&mv, &mi);
  auto t2 = clock::now();
  *fuse_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  *argmax_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
  if (conf) std::memcpy(conf, mv.data, (size_t)dimX * dimY * sizeof(float));
  if (idx) std::memcpy(idx, mi.data, (size_t)dimX * dimY);
  return 0;
}

// Grid3D::computeMeanSquare.  cartesian3dgrid.cpp:164-174
double ref_mean_square(uint32_t dimX, uint32_t dimY, uint32_t dimZ, const float* vol)
{
  Grid3D g(dimX, dimY, dimZ);
  load(g, vol, (size_t)dimX * dimY * dimZ);
  return g.computeMeanSquare();
}

}  // extern "C"
