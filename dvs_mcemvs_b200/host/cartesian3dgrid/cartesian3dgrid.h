// cartesian3dgrid.h — host-side mirror of the reference's Grid3D
// (cartesian3dgrid/include/cartesian3dgrid/cartesian3dgrid.h:22-247) whose volume lives in B200
// HBM.  Same class name, method names, argument meaning and defaults; every method forwards to
// the C-ABI (include/emvs_b200.h).  Differences a caller can observe:
//   * the fusion methods take `const Grid3D&` (the reference copies the argument by value,
//     cartesian3dgrid.h:64,111,...: a 315 MB copy per call at 640x480x256);
//   * getPointerToSlice() returns a pointer into a HOST MIRROR that is downloaded on demand and is
//     read-only: writes through it do not reach the device volume (use upload());
//   * collapseMaxZSlice fills emvs_host::Image<> (cv::Mat when EMVS_HOST_WITH_OPENCV is defined);
//     for dimZ > 256 use collapseMaxZSlice16 (the reference's uchar index cannot represent it);
//   * the focus-based collapses (cartesian3dgrid.cpp:192-483) are not provided: no caller
//     selects them (mapper_emvs_stereo.cpp:367-369 runs collapseMaxZSlice for method = -1).
#pragma once

#include "../emvs_host_common.hpp"

#include <cstdio>
#include <cstring>
#include <iostream>

#ifdef EMVS_HOST_WITH_OPENCV
#include <opencv2/core/core.hpp>
#endif

class Grid3D {
 public:
  Grid3D() {}
  Grid3D(const unsigned int dimX, const unsigned int dimY, const unsigned int dimZ) { allocate(dimX, dimY, dimZ); }
  ~Grid3D() { deallocate(); }
  // deep copy like the reference's implicit copy constructor (used by its by-value arguments)
  Grid3D(const Grid3D& o) { *this = o; }
  Grid3D& operator=(const Grid3D& o)
  {
    if (this == &o) return *this;
    if (g_ && !owned_) {
      // a mapper-attached dsi_ (MapperEMVS::dsi_): the reference's `mapper.dsi_ = other` overwrites the member the
      // mapper keeps voting into, so copy INTO the attached volume instead of detaching from it
      if (!o.g_ || o.size_[0] != size_[0] || o.size_[1] != size_[1] || o.size_[2] != size_[2])
        throw std::runtime_error("Grid3D: assignment to a mapper's dsi_ needs a grid of the same dimensions");
      emvs_host::check(emvs_grid_copy(g_, o.g_), "Grid3D copy");
      mirror_.clear();
      return *this;
    }
    deallocate();
    if (o.g_) {
      allocate(o.size_[0], o.size_[1], o.size_[2]);
      emvs_host::check(emvs_grid_copy(g_, o.g_), "Grid3D copy");
    }
    return *this;
  }

  void allocate(const unsigned int dimX, const unsigned int dimY, const unsigned int dimZ)
  {
    deallocate();
    size_[0] = dimX; size_[1] = dimY; size_[2] = dimZ;
    numCells_ = dimX * dimY * dimZ;
    emvs_host::check(emvs_grid_create(emvs_host::default_context(), dimX, dimY, dimZ, &g_), "Grid3D::allocate");
    owned_ = true;
  }
  void deallocate()
  {
    if (g_ && owned_) emvs_grid_destroy(g_);
    g_ = nullptr;
    size_[0] = size_[1] = size_[2] = 0;
    numCells_ = 0;
    mirror_.clear();
  }
  // Non-owning view of a volume owned by a mapper (MapperEMVS::dsi_).
  void attach(emvs_grid* g)
  {
    deallocate();
    g_ = g;
    owned_ = false;
    emvs_host::check(emvs_grid_dims(g, &size_[0], &size_[1], &size_[2]), "Grid3D::attach");
    numCells_ = size_[0] * size_[1] * size_[2];
  }

  void printInfo() const
  {
    std::cout << "Grid3D Dimensions: (" << size_[0] << "," << size_[1] << "," << size_[2] << ")" << std::endl;
    std::cout << "Grid3D Data_array size: " << numCells_ << std::endl;
  }

  // volume[x + dimX*(y + dimY*z)] = value at (x,y,z)
  float getGridValueAt(const unsigned int ix, const unsigned int iy, const unsigned int iz) const
  {
    return getGridValueAt(ix + size_[0] * (iy + size_[1] * iz));
  }
  float getGridValueAt(const unsigned int p) const
  {
    sync_mirror();
    return mirror_.at(p);
  }

  void resetGrid() { touch(); emvs_host::check(emvs_grid_reset(g_), "resetGrid"); }
  // resetGrid() + addTwoGrids(src) in one device copy (the reference's idiom for "initialise with", process1.cpp:126-127)
  void copyFrom(const Grid3D& src) { touch(); emvs_host::check(emvs_grid_copy(g_, src.g_), "Grid3D::copyFrom"); }

  // voxel-wise operations, cartesian3dgrid.h:64-192
  void addTwoGrids(const Grid3D& grid2) { op(grid2, EMVS_OP_ADD, 0, 0.f); }
  void addInverseOfTwoGrids(const Grid3D& grid2, const float eps = 1e-2) { op(grid2, EMVS_OP_ADD_INV, 0, eps); }
  void computeHMfromSumOfInv(int n) { unary(EMVS_OP_HM_FROM_SUMINV, n); }
  void computeAMfromSum(int n) { unary(EMVS_OP_AM_FROM_SUM, n); }
  void minTwoGrids(const Grid3D& grid2) { op(grid2, EMVS_OP_MIN, 0, 0.f); }
  void harmonicMeanTwoGrids(const Grid3D& grid2, const float eps = 1e-1) { op(grid2, EMVS_OP_HM, 2, eps); }
  void harmonicMeanTwoGrids(const Grid3D& grid2, int n, const float eps = 1e-1) { op(grid2, EMVS_OP_HM_N, n, eps); }
  void rmsTwoGrids(const Grid3D& grid2) { op(grid2, EMVS_OP_RMS, 0, 0.f); }
  void geometricMeanTwoGrids(const Grid3D& grid2) { op(grid2, EMVS_OP_GM, 0, 0.f); }
  void arithmeticMeanTwoGrids(const Grid3D& grid2) { op(grid2, EMVS_OP_AM, 0, 0.f); }
  void maxTwoGrids(const Grid3D& grid2) { op(grid2, EMVS_OP_MAX, 0, 0.f); }

  // cartesian3dgrid.cpp:115-137.  max_pos_idx is 8-bit like the reference's CV_8U (dimZ <= 256).
  void collapseMaxZSlice(emvs_host::Image<float>* max_val, emvs_host::Image<uint8_t>* max_pos_idx) const
  {
    if (size_[2] > 256) throw std::runtime_error("collapseMaxZSlice: dimZ > 256 needs collapseMaxZSlice16");
    max_val->create(size_[1], size_[0]);
    max_pos_idx->create(size_[1], size_[0]);
    emvs_host::check(emvs_grid_collapse_max(g_, nullptr, max_val->data.data(), max_pos_idx->data.data(), nullptr),
                     "collapseMaxZSlice");
  }
  void collapseMaxZSlice16(emvs_host::Image<float>* max_val, emvs_host::Image<uint16_t>* max_pos_idx) const
  {
    max_val->create(size_[1], size_[0]);
    max_pos_idx->create(size_[1], size_[0]);
    if (size_[2] > 256) {
      emvs_host::check(emvs_grid_collapse_max(g_, nullptr, max_val->data.data(), max_pos_idx->data.data(), nullptr),
                       "collapseMaxZSlice16");
    } else {
      emvs_host::Image<uint8_t> tmp;
      collapseMaxZSlice(max_val, &tmp);
      for (size_t i = 0; i < tmp.data.size(); ++i) max_pos_idx->data[i] = tmp.data[i];
    }
  }
#ifdef EMVS_HOST_WITH_OPENCV
  void collapseMaxZSlice(cv::Mat* max_val, cv::Mat* max_pos_idx) const
  {
    *max_val = cv::Mat(size_[1], size_[0], CV_32FC1);
    *max_pos_idx = cv::Mat(size_[1], size_[0], CV_8U);
    emvs_host::check(emvs_grid_collapse_max(g_, nullptr, (float*)max_val->data, max_pos_idx->data, nullptr),
                     "collapseMaxZSlice");
  }
#endif

  double computeMeanSquare() const
  {
    double r = 0.;
    emvs_host::check(emvs_grid_mean_square(g_, &r), "computeMeanSquare");
    return r;
  }

  // cartesian3dgrid_IO.cpp:30-36: NumPy .npy, float32, shape {dimZ, dimY, dimX}, C order.
  int writeGridNpy(const char szFilename[]) const
  {
    sync_mirror();
    FILE* f = std::fopen(szFilename, "wb");
    if (!f) return -1;
    char dict[160];
    int n = std::snprintf(dict, sizeof dict, "{'descr': '<f4', 'fortran_order': False, 'shape': (%u, %u, %u), }", size_[2],
                          size_[1], size_[0]);
    const int total = ((10 + n + 1 + 63) / 64) * 64;  // magic(6)+ver(2)+len(2)+dict+'\n', padded to 64
    std::string hdr(dict, n);
    hdr.append((size_t)(total - 10 - n - 1), ' ');
    hdr.push_back('\n');
    const unsigned char magic[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (unsigned char)(hdr.size() & 0xff),
                                     (unsigned char)(hdr.size() >> 8)};
    std::fwrite(magic, 1, 10, f);
    std::fwrite(hdr.data(), 1, hdr.size(), f);
    std::fwrite(mirror_.data(), sizeof(float), mirror_.size(), f);
    std::fclose(f);
    return 0;
  }

  void getDimensions(int* dimX, int* dimY, int* dimZ) const
  {
    *dimX = size_[0]; *dimY = size_[1]; *dimZ = size_[2];
  }

  // Read-only host mirror of slice `layer` (see the header comment).
  float* getPointerToSlice(int layer)
  {
    sync_mirror();
    return &mirror_.data()[(size_t)layer * size_[0] * size_[1]];
  }

  void upload(const float* host) { touch(); emvs_host::check(emvs_grid_upload(g_, host), "Grid3D::upload"); }
  void touch() const { mirror_valid_ = false; }  // device contents changed behind the mirror
  emvs_grid* handle() const { return g_; }

 private:
  void op(const Grid3D& b, int id, int n, float eps) { touch(); emvs_host::check(emvs_grid_op(g_, b.g_, id, n, eps), "Grid3D op"); }
  void unary(int id, int n) { touch(); emvs_host::check(emvs_grid_op(g_, nullptr, id, n, 0.f), "Grid3D op"); }
  void sync_mirror() const
  {
    if (mirror_valid_ && mirror_.size() == numCells_) return;
    mirror_.resize(numCells_);
    emvs_host::check(emvs_grid_download(g_, mirror_.data()), "Grid3D download");
    mirror_valid_ = true;
  }

  emvs_grid* g_ = nullptr;
  bool owned_ = false;
  unsigned int numCells_ = 0;
  unsigned int size_[3] = {0, 0, 0};
  mutable std::vector<float> mirror_;
  mutable bool mirror_valid_ = false;
};
