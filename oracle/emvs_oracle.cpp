// emvs_oracle.cpp — CPU restatement ("oracle") of the MC-EMVS mapping hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under dvs_mcemvs_b200/ may include, link or call
// this file.  Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline
// leg and `bench.py --impl reference`.
//
// The reference (tub-rip/dvs_mcemvs @ 49fbc7e) ships no tests, golden vectors or fixtures for
// this path, and as a whole it cannot be compiled here (needs ROS/catkin, Eigen, minkindr,
// OpenCV C++, glog, gflags ...; none installed, no network).  This file restates the reference's
// loops from its sources; every function cites the lines it follows.
// PINNED bit-exactly against the reference's own code compiled in place (oracle/Makefile target
// `ref`: cartesian3dgrid.{h,cpp}, depth_vector.hpp, median_filtering.cpp) and against cv2:
// vote(), fuse_op(), collapse_max(), mean_square(), depth_vector(), depth_map_post().
// PARITY UNPINNED: packetize(), pose_at() and friends, warp_events(), fill_voxel_grid()'s Eq.15
// expression order (Eigen / minkindr arithmetic of mapper_emvs_stereo.cpp, which cannot be built).  Paths are relative to the reference root; abbreviations:
//   MAP = mapper_emvs_stereo/src/mapper_emvs_stereo.cpp
//   G3H = cartesian3dgrid/include/cartesian3dgrid/cartesian3dgrid.h
//   G3C = cartesian3dgrid/src/cartesian3dgrid.cpp
//   DV  = mapper_emvs_stereo/include/mapper_emvs_stereo/depth_vector.hpp
//   GU  = mapper_emvs_stereo/include/mapper_emvs_stereo/geometry_utils.hpp
//   TRJ = mapper_emvs_stereo/include/mapper_emvs_stereo/trajectory.hpp
//   P1  = mapper_emvs_stereo/src/process1.cpp,  P2 = .../process2.cpp
//
// Third-party arithmetic that is NOT under the reference tree and is restated from its
// published algorithm (dependencies.yaml pins all of these to `master`, i.e. un-pinned):
//   Eigen (eigen_catkin)   fixed-size float 3x3 product / cofactor inverse, 4x4*4x1
//   minkindr (ethz-asl)    QuatTransformation compose / inverse / log / exp
//   dvs_msgs (rpg_dvs_ros) Event {uint16 x, uint16 y, ros::Time ts, bool polarity}
//
// Normative scalar order: IEEE binary32, round-to-nearest-even, no FMA
// contraction (build with -ffp-contract=off), true division.
//
// Build: see oracle/Makefile  (g++ -std=c++14 -O3 -fopenmp -ffp-contract=off, no -march,
// mirroring mapper_emvs_stereo/CMakeLists.txt:2,14).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace oracle {

// ---------------------------------------------------------------------------------------
// POD types shared with the tests through ctypes.
// ---------------------------------------------------------------------------------------
struct Event {        // dvs_msgs/Event as laid out by roscpp message generation (16 B)
  uint16_t x, y;
  uint32_t sec, nsec; // ros::Time
  uint8_t polarity;
  uint8_t pad[3];
};
static_assert(sizeof(Event) == 16, "event layout");

struct Pose {         // kindr::minimal::QuatTransformation: unit quaternion (w,x,y,z) + position
  double q[4];
  double t[3];
};

struct Packet {       // per-packet output of the packet stage (MAP:88-126)
  float H[9];         // H_z0_px, row-major
  float C[3];         // camera centre in the reference view
  uint64_t first_event;
};
static_assert(sizeof(Packet) == 56, "packet layout");

static const size_t kPacketSize = 1024;  // mapper_emvs_stereo.hpp:153

// ---------------------------------------------------------------------------------------
// Depth sampling — DV:88-103 (linear), DV:131-148 (inverse).  Returns raw_depths_vec_
// exactly as MAP:213-214 builds it: out[i] = cellIndexToDepth(i).
// ---------------------------------------------------------------------------------------
static void depth_vector(int inverse, float min_depth, float max_depth, size_t nz, float* out)
{
  if (min_depth > max_depth) std::swap(min_depth, max_depth);  // DV:34-35
  if (!inverse) {
    const float mult = (float)((nz) / (max_depth - min_depth));          // DV:91
    for (size_t i = 0; i < nz; ++i) out[i] = min_depth + (float)i / mult; // DV:96, :102
  } else {
    const float inv_min = 1.f / min_depth, inv_max = 1.f / max_depth;     // DV:134-135
    const float mult = (float)((nz) / (inv_min - inv_max));               // DV:136
    for (size_t i = 0; i < nz; ++i) {
      const float rho = inv_max + (float)i / mult;                        // DV:141
      out[i] = 1.f / rho;                                                 // DV:147
    }
  }
}

// ---------------------------------------------------------------------------------------
// Eigen-style fixed-size float 3x3 helpers (row-major storage m[3*r+c]).
// Eigen >= 3.3 evaluates a coefficient of a small fixed-size product as
// (lhs.row(i).cwiseProduct(rhs.col(j))).sum(), and sum() of THREE terms unrolls (redux_novec_unroller,
// HalfLength = 3/2 = 1) to a0 + (a1 + a2); four terms give (a0 + a1) + (a2 + a3).  The 3x3 inverse is
// cofactor / determinant with det = (cofactor column 0 .* matrix column 0).sum() — again a0 + (a1 + a2) —
// and a single reciprocal (compute_inverse_size3_helper).  Eigen is not part of the reference tree, so this
// order is restated from Eigen's sources, not checkable here (parity unpinned).
// ---------------------------------------------------------------------------------------
static void mat3_mul(const float* a, const float* b, float* out)
{
  float r[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      r[3 * i + j] = a[3 * i + 0] * b[0 + j] + (a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j]);
  std::memcpy(out, r, sizeof r);
}

static inline float cof3(const float* m, int i, int j)
{
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m[3 * i1 + j1] * m[3 * i2 + j2] - m[3 * i1 + j2] * m[3 * i2 + j1];
}

static void mat3_inv(const float* m, float* out)
{
  const float c0 = cof3(m, 0, 0), c1 = cof3(m, 1, 0), c2 = cof3(m, 2, 0);
  const float det = c0 * m[0] + (c1 * m[3] + c2 * m[6]);
  const float invdet = 1.f / det;
  float r[9];
  r[0] = c0 * invdet; r[1] = c1 * invdet; r[2] = c2 * invdet;
  r[3] = cof3(m, 0, 1) * invdet; r[4] = cof3(m, 1, 1) * invdet; r[5] = cof3(m, 2, 1) * invdet;
  r[6] = cof3(m, 0, 2) * invdet; r[7] = cof3(m, 1, 2) * invdet; r[8] = cof3(m, 2, 2) * invdet;
  std::memcpy(out, r, sizeof r);
}

// GU:43-47: K_ then Kinv_ = K_.inverse()
static void pinhole_kinv(float fx, float fy, float cx, float cy, float* kinv)
{
  const float K[9] = {fx, 0.f, cx, 0.f, fy, cy, 0.f, 0.f, 1.f};
  mat3_inv(K, kinv);
}

// MAP:219-229: focal length of the virtual camera.
static float virtual_focal(float dvs_fx, size_t dimX, float fov_deg)
{
  if (fov_deg < 10.f) return dvs_fx;
  const float dsi_fov_rad = fov_deg * 3.1415926535897932384626433832795 / 180.0;  // CV_PI
  return 0.5 * (float)dimX / std::tan(0.5 * dsi_fov_rad);
}

// ---------------------------------------------------------------------------------------
// minkindr restatement (double).  q = (w,x,y,z) unit quaternion.
// ---------------------------------------------------------------------------------------
static void q_mul(const double* a, const double* b, double* o)
{
  const double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  const double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  const double y = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
  const double z = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}

// Eigen QuaternionBase::_transformVector: v + w*uv + q.vec x uv, uv = 2 (q.vec x v)
static void q_rot(const double* q, const double* v, double* o)
{
  double uv[3] = {q[2] * v[2] - q[3] * v[1], q[3] * v[0] - q[1] * v[2], q[1] * v[1] - q[2] * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  const double c[3] = {q[2] * uv[2] - q[3] * uv[1], q[3] * uv[0] - q[1] * uv[2], q[1] * uv[1] - q[2] * uv[0]};
  const double r0 = v[0] + q[0] * uv[0] + c[0];
  const double r1 = v[1] + q[0] * uv[1] + c[1];
  const double r2 = v[2] + q[0] * uv[2] + c[2];
  o[0] = r0; o[1] = r1; o[2] = r2;
}

static Pose pose_mul(const Pose& a, const Pose& b)  // T_a * T_b
{
  Pose o;
  q_mul(a.q, b.q, o.q);
  double r[3];
  q_rot(a.q, b.t, r);
  for (int i = 0; i < 3; ++i) o.t[i] = a.t[i] + r[i];
  return o;
}

static Pose pose_inv(const Pose& a)
{
  Pose o;
  o.q[0] = a.q[0]; o.q[1] = -a.q[1]; o.q[2] = -a.q[2]; o.q[3] = -a.q[3];
  double r[3];
  q_rot(o.q, a.t, r);
  for (int i = 0; i < 3; ++i) o.t[i] = -r[i];
  return o;
}

// Eigen QuaternionBase::toRotationMatrix (row-major out)
static void q_to_mat(const double* q, double* R)
{
  const double tx = 2 * q[1], ty = 2 * q[2], tz = 2 * q[3];
  const double twx = tx * q[0], twy = ty * q[0], twz = tz * q[0];
  const double txx = tx * q[1], txy = ty * q[1], txz = tz * q[1];
  const double tyy = ty * q[2], tyz = tz * q[2], tzz = tz * q[3];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

static double asin_x_over_x(double x)
{
  if (std::fabs(x) < 1.220703125e-4 /* eps^(1/4) */) return 1.0 + x * x * (1.0 / 6.0);
  return std::asin(x) / x;
}

// minkindr RotationQuaternion::log -> rotation vector (angle * axis)
static void q_log(const double* q, double* o)
{
  const double na = std::sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double eta = q[0];
  double scale;
  if (std::fabs(eta) < na) {
    scale = (eta >= 0) ? std::acos(eta) / na : -std::acos(-eta) / na;
  } else {
    scale = (eta > 0) ? asin_x_over_x(na) : -asin_x_over_x(na);
  }
  for (int i = 0; i < 3; ++i) o[i] = q[1 + i] * (2.0 * scale);
}

// minkindr RotationQuaternion::exp(rotation vector)
static void q_exp(const double* dx, double* q)
{
  const double theta = std::sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]);
  double na;
  if (theta < 1.220703125e-4) na = 0.5 - theta * theta * (1.0 / 48.0);
  else na = std::sin(theta * 0.5) / theta;
  q[0] = std::cos(theta * 0.5);
  q[1] = dx[0] * na; q[2] = dx[1] * na; q[3] = dx[2] * na;
}

// ---------------------------------------------------------------------------------------
// LinearTrajectory::getPoseAt — TRJ:92-127.  Control poses sorted by strictly increasing
// time (std::map<ros::Time,...>).  Returns 0 on "cannot extrapolate".
// ---------------------------------------------------------------------------------------
static inline bool time_less(uint32_t as, uint32_t an, uint32_t bs, uint32_t bn)
{
  return as < bs || (as == bs && an < bn);
}

static inline double duration_sec(uint32_t as, uint32_t an, uint32_t bs, uint32_t bn)  // (a - b).toSec()
{
  int64_t s = (int64_t)as - (int64_t)bs, n = (int64_t)an - (int64_t)bn;
  while (n < 0) { n += 1000000000LL; --s; }
  while (n >= 1000000000LL) { n -= 1000000000LL; ++s; }
  return (double)s + 1e-9 * (double)n;
}

static int pose_at(const uint32_t* tsec, const uint32_t* tnsec, const Pose* poses, size_t n,
                   uint32_t sec, uint32_t nsec, Pose* out)
{
  // upper_bound: first control pose with time > t  (TRJ:98)
  size_t lo = 0, hi = n;
  while (lo < hi) {
    const size_t mid = (lo + hi) / 2;
    if (time_less(sec, nsec, tsec[mid], tnsec[mid])) hi = mid; else lo = mid + 1;
  }
  if (lo == 0 || lo == n) return 0;  // TRJ:99-112
  const size_t i0 = lo - 1, i1 = lo;
  const Pose T_rel = pose_mul(pose_inv(poses[i0]), poses[i1]);                   // TRJ:123
  const double dt = duration_sec(sec, nsec, tsec[i0], tnsec[i0]) /
                    duration_sec(tsec[i1], tnsec[i1], tsec[i0], tnsec[i0]);      // TRJ:124
  double rv[3];
  q_log(T_rel.q, rv);                                                            // log() = [t; so3]
  Pose inc;
  const double drv[3] = {dt * rv[0], dt * rv[1], dt * rv[2]};
  q_exp(drv, inc.q);
  for (int i = 0; i < 3; ++i) inc.t[i] = dt * T_rel.t[i];
  *out = pose_mul(poses[i0], inc);                                               // TRJ:125
  return 1;
}

// ---------------------------------------------------------------------------------------
// Packet stage of evaluateDSI — MAP:86-126.  Returns number of packets written.
//   K       = {fx,fy,cx,cy} of the event camera's projection matrix (MAP:46-48)
//   virt    = {fx,fy,cx,cy} of the virtual camera (MAP:237-239)
//   z0      = raw_depths_vec_[0] (MAP:111)
// ---------------------------------------------------------------------------------------
static size_t packetize(const Event* ev, size_t n_ev,
                        const uint32_t* tsec, const uint32_t* tnsec, const Pose* poses, size_t n_poses,
                        const Pose& T_rv_w, const float* K, const float* virt, float z0,
                        Packet* out, size_t max_out)
{
  if (n_ev < kPacketSize) return 0;                                   // MAP:71-75
  const float Km[9] = {K[0], 0.f, K[2], 0.f, K[1], K[3], 0.f, 0.f, 1.f};
  float Kinv_virtual[9];
  pinhole_kinv(virt[0], virt[1], virt[2], virt[3], Kinv_virtual);
  size_t cur = 0, n_out = 0;
  while (cur + kPacketSize < n_ev) {                                  // MAP:88 (strict <)
    const Event& mid = ev[cur + kPacketSize / 2];                     // MAP:91
    Pose T_w_ev;
    if (!pose_at(tsec, tnsec, poses, n_poses, mid.sec, mid.nsec, &T_w_ev)) {
      ++cur;                                                          // MAP:95-99
      continue;
    }
    const Pose T_ev_rv = pose_inv(pose_mul(T_rv_w, T_w_ev));          // MAP:101-103
    double Rd[9];
    q_to_mat(T_ev_rv.q, Rd);
    float R[9], t[3];
    for (int i = 0; i < 9; ++i) R[i] = (float)Rd[i];                  // MAP:104
    for (int i = 0; i < 3; ++i) t[i] = (float)T_ev_rv.t[i];           // MAP:105
    if (n_out >= max_out) return n_out;
    Packet& p = out[n_out++];
    p.first_event = cur;
    for (int i = 0; i < 3; ++i)                                       // MAP:108  -R^T t
      p.C[i] = (-R[0 + i]) * t[0] + ((-R[3 + i]) * t[1] + (-R[6 + i]) * t[2]);
    float Hinv[9];
    for (int i = 0; i < 9; ++i) Hinv[i] = R[i] * z0;                  // MAP:114-115
    Hinv[2] += t[0]; Hinv[5] += t[1]; Hinv[8] += t[2];                // MAP:116
    float tmp[9], Hinv_px[9];
    mat3_mul(Km, Hinv, tmp);                                          // MAP:119
    mat3_mul(tmp, Kinv_virtual, Hinv_px);
    mat3_inv(Hinv_px, p.H);                                           // MAP:120
    cur += kPacketSize;                                               // MAP:129-131
  }
  return n_out;
}

// ---------------------------------------------------------------------------------------
// Event stage — MAP:129-142.  xy0 receives (X0, Y0) for the 1024 events of each packet.
// lut = interleaved (x,y) rectified points indexed y*W+x (MAP:134, :296).
// ---------------------------------------------------------------------------------------
static void warp_events(const Event* ev, const Packet* pk, size_t n_pk, const float* lut, int W, float* xy0)
{
  for (size_t j = 0; j < n_pk; ++j) {
    const float* H = pk[j].H;
    for (size_t i = 0; i < kPacketSize; ++i) {
      const Event& e = ev[pk[j].first_event + i];
      const float xr = lut[2 * ((size_t)e.y * W + e.x)], yr = lut[2 * ((size_t)e.y * W + e.x) + 1];
      const float p0 = (H[0] * xr + H[1] * yr) + H[2];
      const float p1 = (H[3] * xr + H[4] * yr) + H[5];
      const float p2 = (H[6] * xr + H[7] * yr) + H[8];
      xy0[2 * (j * kPacketSize + i)] = p0 / p2;                       // MAP:139
      xy0[2 * (j * kPacketSize + i) + 1] = p1 / p2;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Bilinear vote — G3H:253-273.  Returns 1 when the vote is accepted (build-defined integer
// observable).  `const int x = x_f` on x86-64 yields INT_MIN for values that do
// not fit, which then fails the unsigned bound test; stated here explicitly.
// ---------------------------------------------------------------------------------------
static inline int vote(float x_f, float y_f, float* grid, unsigned dimX, unsigned dimY)
{
  if (x_f >= 0.f && y_f >= 0.f) {
    if (!(x_f < 2147483648.f) || !(y_f < 2147483648.f)) return 0;
    const int x = (int)x_f, y = (int)y_f;
    if ((unsigned)(x + 1) < dimX && (unsigned)(y + 1) < dimY) {
      float* g = grid + x + (size_t)y * dimX;
      const float fx = x_f - x, fy = y_f - y, fx1 = 1.f - fx, fy1 = 1.f - fy;
      g[0] += fx1 * fy1;
      g[1] += fx * fy1;
      g[dimX] += fx1 * fy;
      g[dimX + 1] += fx * fy;
      return 1;
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// fillVoxelGrid — MAP:151-205.  Same loop nest: OpenMP over planes (enabled at >= 20000
// warped events), packets, batches of 128, vote.  inb[k] (may be null) = accepted votes
// on plane k.
// ---------------------------------------------------------------------------------------
static void fill_voxel_grid(const float* xy0, const Packet* pk, size_t n_pk, const float* depths, size_t nz,
                            const float* virt, unsigned dimX, unsigned dimY, float* dsi, uint64_t* inb)
{
  static const int N = 128;                                           // MAP:160
  const float z0 = depths[0];                                         // MAP:163
  const float vfx = virt[0], vfy = virt[1], vcx = virt[2], vcy = virt[3];
  const size_t n_warped = n_pk * kPacketSize;
#pragma omp parallel for if (n_warped >= 20000)                       // MAP:168
  for (size_t k = 0; k < nz; ++k) {
    const float* pe = xy0;
    float* pgrid = dsi + k * (size_t)dimX * dimY;                     // G3H:237-240
    uint64_t acc = 0;
    for (size_t j = 0; j < n_pk; ++j) {
      const float* C = pk[j].C;
      const float zi = depths[k],                                     // MAP:178-182
                  a = z0 * (zi - C[2]),
                  bx = (z0 - zi) * (C[0] * vfx + C[2] * vcx),
                  by = (z0 - zi) * (C[1] * vfy + C[2] * vcy),
                  d = zi * (z0 - C[2]);
      for (size_t batch = 0; batch < kPacketSize / N; ++batch, pe += 2 * N) {
        float X[N], Y[N];
        for (int i = 0; i < N; ++i) { X[i] = pe[2 * i]; Y[i] = pe[2 * i + 1]; }
        for (int i = 0; i < N; ++i) { X[i] = (X[i] * a + bx) / d; Y[i] = (Y[i] * a + by) / d; }  // MAP:194-195
        for (int i = 0; i < N; ++i) acc += vote(X[i], Y[i], pgrid, dimX, dimY);                  // MAP:200
      }
    }
    if (inb) inb[k] = acc;
  }
}

// ---------------------------------------------------------------------------------------
// Voxel-wise fusion — G3H:64-192.  The reference passes grid2 BY VALUE (a full copy per
// call) and uses bounds-checked .at(); `copy_arg` reproduces the copy so the timed CPU
// baseline pays what the reference pays.
// ops: 0 add, 1 min, 2 HM, 3 GM, 4 AM, 5 RMS, 6 max, 7 HM-n (n, eps), 8 addInverse (eps),
//      9 HM-from-sum-of-inverse (n), 10 AM-from-sum (n)
// ---------------------------------------------------------------------------------------
static void fuse_op(int op, float* a, const float* b_in, size_t n_cells, int n, float eps, int copy_arg)
{
  std::vector<float> copy;
  const float* b = b_in;
  if (copy_arg && b_in) { copy.assign(b_in, b_in + n_cells); b = copy.data(); }
  switch (op) {
    case 0: for (size_t p = 0; p < n_cells; ++p) a[p] += b[p]; break;                             // G3H:64-70
    case 1: for (size_t p = 0; p < n_cells; ++p) a[p] = std::min(a[p], b[p]); break;              // G3H:111-117
    case 2: for (size_t p = 0; p < n_cells; ++p) {                                                // G3H:119-127
        const float prod = a[p] * b[p], sum = a[p] + b[p];
        a[p] = 2 * prod / (sum + eps);
      } break;
    case 3: for (size_t p = 0; p < n_cells; ++p) a[p] = std::sqrt(a[p] * b[p]); break;            // G3H:150-156
    case 4: for (size_t p = 0; p < n_cells; ++p) a[p] = 0.5 * (a[p] + b[p]); break;               // G3H:158-164
    case 5: for (size_t p = 0; p < n_cells; ++p) {                                                // G3H:141-148
        const float ms = 0.5 * (std::pow((double)a[p], 2) + std::pow((double)b[p], 2));
        a[p] = std::sqrt(ms);
      } break;
    case 6: for (size_t p = 0; p < n_cells; ++p) a[p] = std::max(a[p], b[p]); break;              // G3H:186-192
    case 7: for (size_t p = 0; p < n_cells; ++p) {                                                // G3H:130-139
        const float aa = a[p] / (float)(n - 1);
        const float prod = aa * b[p], sum = aa + b[p];
        a[p] = n * prod / (sum + eps);
      } break;
    case 8: for (size_t p = 0; p < n_cells; ++p) a[p] = a[p] + 1.0f / (eps + b[p]); break;        // G3H:72-78
    case 9: for (size_t p = 0; p < n_cells; ++p) a[p] = (float)n / a[p]; break;                   // G3H:80-86
    case 10: for (size_t p = 0; p < n_cells; ++p) a[p] = a[p] / (float)n; break;                  // G3H:87-93
    default: break;
  }
}

// ---------------------------------------------------------------------------------------
// EXTENSION, not in the reference: n-ary fusion for more cameras than the
// reference handles.  Defined so that n == 2 is exactly the pairwise reference op, and HM with
// n == 3 is exactly process1.cpp:176.  GM_n = (prod)^(1/n) (sqrt for 2, sqrt(sqrt) for 4, pow
// in double otherwise), AM_n = sum/n, RMS_n = sqrt(float(sum of double squares / n)).
// ---------------------------------------------------------------------------------------
static void fuse_nary(int method, const float* const* v, int n, size_t n_cells, float* out)
{
  for (size_t p = 0; p < n_cells; ++p) {
    float a = v[0][p];
    if (n > 1) switch (method) {
      case 1: for (int i = 1; i < n; ++i) a = std::min(a, v[i][p]); break;
      case 6: for (int i = 1; i < n; ++i) a = std::max(a, v[i][p]); break;
      case 2: {
        { const float prod = a * v[1][p], sum = a + v[1][p]; a = 2 * prod / (sum + 0.1f); }
        for (int i = 2; i < n; ++i) {
          const float aa = a / (float)i;  // n_ref - 1 with n_ref = i + 1
          const float prod = aa * v[i][p], sum = aa + v[i][p];
          a = (i + 1) * prod / (sum + 0.1f);
        }
      } break;
      case 3: {
        float prod = a * v[1][p];
        for (int i = 2; i < n; ++i) prod = prod * v[i][p];
        if (n == 2) a = std::sqrt(prod);
        else if (n == 4) a = std::sqrt(std::sqrt(prod));
        else a = (float)std::pow((double)prod, 1.0 / (double)n);
      } break;
      case 4: {
        float s = a + v[1][p];
        for (int i = 2; i < n; ++i) s = s + v[i][p];
        a = (n == 2) ? 0.5f * s : s / (float)n;
      } break;
      case 5: {
        double s = (double)a * (double)a + (double)v[1][p] * (double)v[1][p];
        for (int i = 2; i < n; ++i) s = s + (double)v[i][p] * (double)v[i][p];
        a = std::sqrt((float)(s / (double)n));
      } break;
      default: break;
    }
    out[p] = a;
  }
}

// ---------------------------------------------------------------------------------------
// collapseMaxZSlice — G3C:115-137: per pixel copy the Z column, std::max_element (first
// maximum, operator<).  idx is written as u16 (reference: uchar, dimZ <= 256).
// depth (may be null) = depths[idx] for every pixel (MAP:302-313).
// ---------------------------------------------------------------------------------------
static void collapse_max(const float* dsi, unsigned dimX, unsigned dimY, unsigned dimZ,
                         const float* depths, float* conf, uint16_t* idx, float* depth)
{
  std::vector<float> col(dimZ);
  const size_t plane = (size_t)dimX * dimY;
  for (unsigned v = 0; v < dimY; ++v)
    for (unsigned u = 0; u < dimX; ++u) {
      for (unsigned k = 0; k < dimZ; ++k) col[k] = dsi[u + (size_t)dimX * v + plane * k];
      const auto mx = std::max_element(col.begin(), col.end());
      const size_t o = (size_t)v * dimX + u;
      conf[o] = *mx;
      idx[o] = (uint16_t)std::distance(col.begin(), mx);
      if (depth) depth[o] = depths[idx[o]];
    }
}

// computeMeanSquare — G3C:164-174
static double mean_square(const float* dsi, size_t n_cells)
{
  double r = 0.;
  for (size_t i = 0; i < n_cells; ++i) { const double t = (double)dsi[i]; r += t * t; }
  return r / (double)n_cells;
}

// ---------------------------------------------------------------------------------------
// Depth-map post-processing — MAP:393-436 without the Telea inpainting.
// OpenCV arithmetic restated (and pinned against cv2 4.13 + the reference's own
// huangMedianFilter by tests/golden/make_golden_depthmap.py):
//   conf(0,0) = max_confidence                                                   MAP:396
//   cv::normalize(conf, conf8, 0, 255, NORM_MINMAX) on CV_32F:                    MAP:397
//       scale = float(255 / (smax - smin))  (0 when smax - smin <= DBL_EPSILON),
//       shift = 0.f - float(smin * scale),   dst = src * scale + shift   (float)
//   conf8(0,0) = 0; convertTo(CV_8U) = saturate(round-half-even)                  MAP:399-400
//   cv::adaptiveThreshold(conf8, mask, 1, GAUSSIAN_C, THRESH_BINARY, ks, -c):     MAP:405-411
//       mean = round-half-even(separable [1 2 1]/4, [1 4 6 4 1]/16 or [2 7 14 18 14 7 2]/64
//       blur, replicated border) — exact integer arithmetic because the sigma=0 kernels of
//       size <= 7 are dyadic; mask = (conf8 - mean > -ceil(-c)) ? 1 : 0
//   huangMedianFilter(idx, idx_filtered, mask, median_size)                       MAP:419-423
//       = for EVERY pixel the lower median of the masked-in indices in the window, 0 if none
//   removeMaskBoundary(mask, max(ks/2, 1))                                        MAP:426-427
//   depth = depths[idx_filtered]                                                  MAP:435
// Returns 0, or -1 for an unsupported kernel size.
// ---------------------------------------------------------------------------------------
static int depth_map_post(float* conf, const uint8_t* idx, int rows, int cols, int ks, double c, double max_confidence,
                          int median_size, const float* depths, uint8_t* conf8_out, uint8_t* mask, uint8_t* idx_filtered,
                          float* depth)
{
  static const int tab3[] = {1, 2, 1}, tab5[] = {1, 4, 6, 4, 1}, tab7[] = {2, 7, 14, 18, 14, 7, 2};
  const int* tab = ks == 3 ? tab3 : ks == 5 ? tab5 : ks == 7 ? tab7 : nullptr;
  if (!tab || median_size < 1 || median_size % 2 == 0) return -1;
  const size_t n = (size_t)rows * cols;
  conf[0] = (float)max_confidence;
  float smin = conf[0], smax = conf[0];
  for (size_t i = 1; i < n; ++i) { smin = std::min(smin, conf[i]); smax = std::max(smax, conf[i]); }
  const double dscale = 255.0 * (((double)smax - (double)smin) > 2.220446049250313e-16 ? 1.0 / ((double)smax - (double)smin) : 0.0);
  const float scale = (float)dscale;
  const float shift = 0.f - (float)((double)smin * (double)scale);
  std::vector<uint8_t> c8(n);
  for (size_t i = 0; i < n; ++i) {
    float v = conf[i] * scale + shift;
    if (i == 0) v = 0.f;
    const float r = std::nearbyint(v);   // round-half-even (default rounding mode) == cvRound
    c8[i] = (uint8_t)(r < 0.f ? 0 : r > 255.f ? 255 : (int)r);   // NaN compares false twice -> (int)NaN; conf is finite here
  }
  if (conf8_out) std::memcpy(conf8_out, c8.data(), n);
  int den = 0;
  for (int i = 0; i < ks; ++i) den += tab[i];
  const long long D = (long long)den * den;
  const int p = ks / 2;
  const int idelta = (int)std::ceil(-c);
  auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; };
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      long long acc = 0;
      for (int i = -p; i <= p; ++i)
        for (int j = -p; j <= p; ++j)
          acc += (long long)tab[i + p] * tab[j + p] * c8[(size_t)clampi(y + i, 0, rows - 1) * cols + clampi(x + j, 0, cols - 1)];
      long long q = acc / D;
      const long long r = acc % D;
      if (2 * r > D || (2 * r == D && (q & 1))) ++q;
      mask[(size_t)y * cols + x] = ((int)c8[(size_t)y * cols + x] - (int)q > -idelta) ? 1 : 0;
    }
  const int mp = median_size / 2;
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      int h[256] = {0}, num = 0;
      for (int i = -mp; i <= mp; ++i)
        for (int j = -mp; j <= mp; ++j) {
          const int yy = y + i, xx = x + j;
          if (yy >= 0 && xx >= 0 && yy < rows && xx < cols && mask[(size_t)yy * cols + xx] > 0) { ++h[idx[(size_t)yy * cols + xx]]; ++num; }
        }
      const int middle = (num + 1) / 2;
      int m = 0, v = 0;
      for (; v < 256; ++v) { m += h[v]; if (m >= middle) break; }
      idx_filtered[(size_t)y * cols + x] = (uint8_t)v;
    }
  const int border = std::max(ks / 2, 1);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      if (x <= border || x >= cols - border || y <= border || y >= rows - border) mask[(size_t)y * cols + x] = 0;
      if (depth) depth[(size_t)y * cols + x] = depths[idx_filtered[(size_t)y * cols + x]];
    }
  return 0;
}

}  // namespace oracle

// ---------------------------------------------------------------------------------------
// C entry points for the tests (ctypes).
// ---------------------------------------------------------------------------------------
extern "C" {

void oracle_depth_vector(int inverse, float zmin, float zmax, uint64_t nz, float* out)
{ oracle::depth_vector(inverse, zmin, zmax, (size_t)nz, out); }

float oracle_virtual_focal(float dvs_fx, uint64_t dimX, float fov_deg)
{ return oracle::virtual_focal(dvs_fx, (size_t)dimX, fov_deg); }

void oracle_pinhole_kinv(float fx, float fy, float cx, float cy, float* kinv9)
{ oracle::pinhole_kinv(fx, fy, cx, cy, kinv9); }

void oracle_mat3_inv(const float* m, float* out) { oracle::mat3_inv(m, out); }
void oracle_mat3_mul(const float* a, const float* b, float* out) { oracle::mat3_mul(a, b, out); }

void oracle_pose_mul(const oracle::Pose* a, const oracle::Pose* b, oracle::Pose* o) { *o = oracle::pose_mul(*a, *b); }
void oracle_pose_inv(const oracle::Pose* a, oracle::Pose* o) { *o = oracle::pose_inv(*a); }

int oracle_pose_at(const uint32_t* tsec, const uint32_t* tnsec, const oracle::Pose* poses, uint64_t n,
                   uint32_t sec, uint32_t nsec, oracle::Pose* out)
{ return oracle::pose_at(tsec, tnsec, poses, (size_t)n, sec, nsec, out); }

uint64_t oracle_packetize(const oracle::Event* ev, uint64_t n_ev, const uint32_t* tsec, const uint32_t* tnsec,
                          const oracle::Pose* poses, uint64_t n_poses, const oracle::Pose* T_rv_w,
                          const float* K4, const float* virt4, float z0, oracle::Packet* out, uint64_t max_out)
{ return oracle::packetize(ev, n_ev, tsec, tnsec, poses, n_poses, *T_rv_w, K4, virt4, z0, out, max_out); }

void oracle_warp_events(const oracle::Event* ev, const oracle::Packet* pk, uint64_t n_pk, const float* lut, int W, float* xy0)
{ oracle::warp_events(ev, pk, n_pk, lut, W, xy0); }

void oracle_fill_voxel_grid(const float* xy0, const oracle::Packet* pk, uint64_t n_pk, const float* depths, uint64_t nz,
                            const float* virt4, uint32_t dimX, uint32_t dimY, float* dsi, uint64_t* inb)
{ oracle::fill_voxel_grid(xy0, pk, n_pk, depths, nz, virt4, dimX, dimY, dsi, inb); }

// evaluateDSI event stage + reset + fill (MAP:129-146) from already-computed packets.
void oracle_build_dsi(const oracle::Event* ev, const oracle::Packet* pk, uint64_t n_pk, const float* lut, int W,
                      const float* depths, uint64_t nz, const float* virt4, uint32_t dimX, uint32_t dimY,
                      float* dsi, uint64_t* inb)
{
  std::vector<float> xy0(2 * n_pk * oracle::kPacketSize);
  oracle::warp_events(ev, pk, n_pk, lut, W, xy0.data());
  std::fill(dsi, dsi + (size_t)dimX * dimY * nz, 0.f);               // G3C:67-70
  oracle::fill_voxel_grid(xy0.data(), pk, n_pk, depths, nz, virt4, dimX, dimY, dsi, inb);
}

int oracle_vote(float x, float y, float* grid, uint32_t dimX, uint32_t dimY) { return oracle::vote(x, y, grid, dimX, dimY); }

void oracle_fuse_op(int op, float* a, const float* b, uint64_t n_cells, int n, float eps, int copy_arg)
{ oracle::fuse_op(op, a, b, n_cells, n, eps, copy_arg); }

void oracle_fuse_nary(int method, const float* const* vols, int n, uint64_t n_cells, float* out)
{ oracle::fuse_nary(method, vols, n, n_cells, out); }

void oracle_collapse_max(const float* dsi, uint32_t dimX, uint32_t dimY, uint32_t dimZ, const float* depths,
                         float* conf, uint16_t* idx, float* depth)
{ oracle::collapse_max(dsi, dimX, dimY, dimZ, depths, conf, idx, depth); }

double oracle_mean_square(const float* dsi, uint64_t n) { return oracle::mean_square(dsi, n); }

int oracle_depth_map_post(float* conf, const uint8_t* idx, int rows, int cols, int ks, double c, double max_confidence,
                          int median_size, const float* depths, uint8_t* conf8, uint8_t* mask, uint8_t* idx_filtered, float* depth)
{ return oracle::depth_map_post(conf, idx, rows, cols, ks, c, max_confidence, median_size, depths, conf8, mask, idx_filtered, depth); }

// torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline must still use every host thread.
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
