// host_geometry.cpp — host half of the mapping path: depth sampling, virtual camera, SE(3)
// trajectory interpolation and the per-packet homography stage of MapperEMVS::evaluateDSI.
// One pose + one 3x3 inverse per 1024 events, so it stays on the host.
//
// Reference behaviour (paths relative to the reference root):
//   depth tables        mapper_emvs_stereo/include/mapper_emvs_stereo/depth_vector.hpp:88-103,131-148
//   virtual camera      mapper_emvs_stereo/src/mapper_emvs_stereo.cpp:208-241
//   pinhole K, K^-1     mapper_emvs_stereo/include/mapper_emvs_stereo/geometry_utils.hpp:28-48
//   pose interpolation  mapper_emvs_stereo/include/mapper_emvs_stereo/trajectory.hpp:92-127
//   packet stage        mapper_emvs_stereo/src/mapper_emvs_stereo.cpp:67-126
// minkindr (SE(3)) and Eigen (3x3 float algebra) are external to the reference tree; their
// arithmetic is restated here (see DESIGN.md "Third-party arithmetic").
//
// Compiled with -ffp-contract=off: the float chain must not be contracted into FMAs, the
// device kernels consume these bits and the oracle recomputes them.

#include "emvs_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace emvs {
namespace {

// ---- small fixed-size float algebra with Eigen's (>= 3.3) evaluation order --------------------
// three-term sums are a0 + (a1 + a2): (row .* col).sum() unrolls through redux_novec_unroller with HalfLength = 3/2 = 1
struct Mat3f {
  float m[3][3];
};

inline Mat3f operator*(const Mat3f& a, const Mat3f& b)
{
  Mat3f r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      // Eigen >= 3.3: (row .* col).sum() of three terms unrolls to a0 + (a1 + a2)
      const float s12 = a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
      r.m[i][j] = a.m[i][0] * b.m[0][j] + s12;
    }
  return r;
}

inline float minor2(const Mat3f& a, int r, int c)
{
  const int r1 = (r + 1) % 3, r2 = (r + 2) % 3, c1 = (c + 1) % 3, c2 = (c + 2) % 3;
  return a.m[r1][c1] * a.m[r2][c2] - a.m[r1][c2] * a.m[r2][c1];
}

// Cofactor inverse: adj(A)/det(A), determinant expanded along column 0, one reciprocal.
inline Mat3f inverse(const Mat3f& a)
{
  float cof[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) cof[r][c] = minor2(a, r, c);
  const float det12 = cof[1][0] * a.m[1][0] + cof[2][0] * a.m[2][0];
  const float det = cof[0][0] * a.m[0][0] + det12;
  const float inv_det = 1.f / det;
  Mat3f out;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) out.m[r][c] = cof[c][r] * inv_det;
  return out;
}

inline Mat3f pinhole_K(float fx, float fy, float cx, float cy)
{
  Mat3f K = {{{fx, 0.f, cx}, {0.f, fy, cy}, {0.f, 0.f, 1.f}}};
  return K;
}

// ---- SE(3) in double, quaternion (w, x, y, z) -------------------------------------------------
struct Vec3 {
  double x, y, z;
};
inline Vec3 cross(const Vec3& a, const Vec3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

struct Quat {
  double w, x, y, z;
  Vec3 vec() const { return {x, y, z}; }
  Quat conj() const { return {w, -x, -y, -z}; }
  Quat operator*(const Quat& o) const
  {
    return {w * o.w - x * o.x - y * o.y - z * o.z,
            w * o.x + x * o.w + y * o.z - z * o.y,
            w * o.y + y * o.w + z * o.x - x * o.z,
            w * o.z + z * o.w + x * o.y - y * o.x};
  }
  // v' = v + w*uv + q_v x uv with uv = 2 (q_v x v)
  Vec3 rotate(const Vec3& v) const
  {
    Vec3 uv = cross(vec(), v);
    uv.x += uv.x; uv.y += uv.y; uv.z += uv.z;
    const Vec3 c = cross(vec(), uv);
    return {v.x + w * uv.x + c.x, v.y + w * uv.y + c.y, v.z + w * uv.z + c.z};
  }
  void to_matrix(double R[3][3]) const
  {
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0][0] = 1 - (tyy + tzz); R[0][1] = txy - twz;       R[0][2] = txz + twy;
    R[1][0] = txy + twz;       R[1][1] = 1 - (txx + tzz); R[1][2] = tyz - twx;
    R[2][0] = txz - twy;       R[2][1] = tyz + twx;       R[2][2] = 1 - (txx + tyy);
  }
};

const double kEps4thRoot = 1.220703125e-4;  // DBL_EPSILON^(1/4)

inline double asin_over_x(double x) { return std::fabs(x) < kEps4thRoot ? 1.0 + x * x * (1.0 / 6.0) : std::asin(x) / x; }

// rotation vector of a unit quaternion, log(-q) == log(q)
inline Vec3 quat_log(const Quat& q)
{
  const double na = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  double scale;
  if (std::fabs(q.w) < na) scale = q.w >= 0 ? std::acos(q.w) / na : -std::acos(-q.w) / na;
  else scale = q.w > 0 ? asin_over_x(na) : -asin_over_x(na);
  const double s2 = 2.0 * scale;
  return {q.x * s2, q.y * s2, q.z * s2};
}

inline Quat quat_exp(const Vec3& r)
{
  const double theta = std::sqrt(r.x * r.x + r.y * r.y + r.z * r.z);
  const double na = theta < kEps4thRoot ? 0.5 - theta * theta * (1.0 / 48.0) : std::sin(theta * 0.5) / theta;
  return {std::cos(theta * 0.5), r.x * na, r.y * na, r.z * na};
}

struct SE3 {
  Quat q;
  Vec3 t;
  SE3 operator*(const SE3& o) const
  {
    const Vec3 r = q.rotate(o.t);
    return {q * o.q, {t.x + r.x, t.y + r.y, t.z + r.z}};
  }
  SE3 inverse() const
  {
    const Quat qi = q.conj();
    const Vec3 r = qi.rotate(t);
    return {qi, {-r.x, -r.y, -r.z}};
  }
};

inline SE3 load(const emvs_pose& p) { return {{p.q[0], p.q[1], p.q[2], p.q[3]}, {p.t[0], p.t[1], p.t[2]}}; }
inline void store(const SE3& s, emvs_pose* p)
{
  p->q[0] = s.q.w; p->q[1] = s.q.x; p->q[2] = s.q.y; p->q[3] = s.q.z;
  p->t[0] = s.t.x; p->t[1] = s.t.y; p->t[2] = s.t.z;
}

// ros::Time ordering and ros::Duration::toSec()
inline bool before(uint32_t as, uint32_t an, uint32_t bs, uint32_t bn) { return as != bs ? as < bs : an < bn; }
inline double seconds_between(uint32_t as, uint32_t an, uint32_t bs, uint32_t bn)  // a - b
{
  int64_t s = (int64_t)as - (int64_t)bs;
  int64_t n = (int64_t)an - (int64_t)bn;
  if (n < 0) { n += 1000000000LL; s -= 1; }
  if (n >= 1000000000LL) { n -= 1000000000LL; s += 1; }
  return (double)s + 1e-9 * (double)n;
}

bool interpolate(const emvs_stamped_pose* traj, size_t n, uint32_t sec, uint32_t nsec, SE3* out)
{
  // first control pose strictly after t
  size_t lo = 0, hi = n;
  while (lo < hi) {
    const size_t mid = lo + (hi - lo) / 2;
    if (before(sec, nsec, traj[mid].sec, traj[mid].nsec)) hi = mid; else lo = mid + 1;
  }
  if (lo == 0 || lo == n) return false;  // no extrapolation, past or future
  const emvs_stamped_pose& p0 = traj[lo - 1];
  const emvs_stamped_pose& p1 = traj[lo];
  const SE3 T0 = load(p0.T), T1 = load(p1.T);
  const SE3 rel = T0.inverse() * T1;
  const double s = seconds_between(sec, nsec, p0.sec, p0.nsec) / seconds_between(p1.sec, p1.nsec, p0.sec, p0.nsec);
  const Vec3 w = quat_log(rel.q);
  const SE3 step = {quat_exp({s * w.x, s * w.y, s * w.z}), {s * rel.t.x, s * rel.t.y, s * rel.t.z}};
  *out = T0 * step;
  return true;
}

}  // namespace

// ---- exported host helpers --------------------------------------------------------------------

void host_depth_vector(const emvs_shape& sh, float* out)
{
  float zmin = sh.min_depth, zmax = sh.max_depth;
  if (zmin > zmax) { const float t = zmin; zmin = zmax; zmax = t; }
  const size_t nz = sh.dimZ;
  if (!sh.inverse_depth) {
    const float mult = (float)(nz / (zmax - zmin));
    for (size_t i = 0; i < nz; ++i) out[i] = zmin + (float)i / mult;
  } else {
    const float rho_near = 1.f / zmin, rho_far = 1.f / zmax;
    const float mult = (float)(nz / (rho_near - rho_far));
    for (size_t i = 0; i < nz; ++i) out[i] = 1.f / (rho_far + (float)i / mult);
  }
}

void host_virtual_camera(const emvs_camera& cam, const emvs_shape& sh, float out[4])
{
  const uint32_t dimX = sh.dimX ? sh.dimX : cam.width;
  float f;
  if (sh.fov_deg < 10.f) {
    f = cam.fx;
  } else {
    const float fov_rad = sh.fov_deg * 3.1415926535897932384626433832795 / 180.0;
    f = 0.5 * (float)dimX / std::tan(0.5 * fov_rad);
  }
  out[0] = f; out[1] = f; out[2] = cam.cx; out[3] = cam.cy;  // principal point of the real camera
}

bool host_pose_at(const emvs_stamped_pose* traj, size_t n, uint32_t sec, uint32_t nsec, emvs_pose* out)
{
  SE3 T;
  if (!interpolate(traj, n, sec, nsec, &T)) return false;
  store(T, out);
  return true;
}

void host_pose_compose(const emvs_pose& a, const emvs_pose& b, emvs_pose* out) { store(load(a) * load(b), out); }
void host_pose_inverse(const emvs_pose& a, emvs_pose* out) { store(load(a).inverse(), out); }

// ---- rectification LUT (precomputeRectifiedPoints, mapper_emvs_stereo.cpp:244-299) -------------------
// The reference fills the LUT with image_geometry::PinholeCameraModel::rectifyPoint (plumb_bob: a float
// pixel through cv::undistortPoints(K, D, R, P), 5 fixed-point iterations, result stored as float; all-zero
// D returns the raw pixel) or cv::fisheye::undistortPoints (equidistant model, Newton iterations on theta,
// COUNT 10 + EPS 1e-8).  OpenCV is not part of the reference tree; both are restated from OpenCV 4.x
// (calib3d/undistort.dispatch.cpp, fisheye.cpp) and pinned against this image's cv2 in
// tests/test_rectify_lut.py.
static void mat3_mul_d(const double* a, const double* b, double* o)
{
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += a[3 * i + k] * b[3 * k + j];
      o[3 * i + j] = s;
    }
}

int host_rectify_lut(int model, const double K[9], const double* D, int n_d, const double R[9], const double P[12],
                     uint32_t W, uint32_t H, float* out)
{
  if (model != EMVS_DISTORTION_NONE && model != EMVS_DISTORTION_PLUMB_BOB && model != EMVS_DISTORTION_FISHEYE) return 1;
  double k[14] = {0};
  for (int i = 0; i < n_d && i < 14; ++i) k[i] = D[i];
  bool all_zero = true;
  for (int i = 0; i < 14; ++i) all_zero = all_zero && k[i] == 0.0;
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  const double PP[9] = {P[0], P[1], P[2], P[4], P[5], P[6], P[8], P[9], P[10]};
  double RR[9];
  mat3_mul_d(PP, R, RR);
  const double ifx = 1. / fx, ify = 1. / fy;
  for (uint32_t py = 0; py < H; ++py)
    for (uint32_t px = 0; px < W; ++px) {
      float* o = out + 2 * ((size_t)py * W + px);
      if (model == EMVS_DISTORTION_NONE || (model == EMVS_DISTORTION_PLUMB_BOB && all_zero)) {
        o[0] = (float)px; o[1] = (float)py;   // image_geometry: distortion_state NONE -> the raw pixel
        continue;
      }
      const double u = (double)(float)px, v = (double)(float)py;
      if (model == EMVS_DISTORTION_PLUMB_BOB) {
        double x = (u - cx) * ifx, y = (v - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; ++j) {
          const double r2 = x * x + y * y;
          const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
          if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
          const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
          const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
          x = (x0 - deltaX) * icdist;
          y = (y0 - deltaY) * icdist;
        }
        const double xx = RR[0] * x + RR[1] * y + RR[2], yy = RR[3] * x + RR[4] * y + RR[5];
        const double ww = 1. / (RR[6] * x + RR[7] * y + RR[8]);
        o[0] = (float)(xx * ww); o[1] = (float)(yy * ww);
      } else {
        const double pwx = (u - cx) / fx, pwy = (v - cy) / fy;
        double theta_d = std::sqrt(pwx * pwx + pwy * pwy);
        const double half_pi = 3.1415926535897932384626433832795 / 2.;
        theta_d = std::min(std::max(-half_pi, theta_d), half_pi);
        bool converged = false;
        double theta = theta_d, scale = 0.0;
        const double eps = 1e-8;
        if (std::fabs(theta_d) > eps) {
          for (int j = 0; j < 10; ++j) {
            const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
            const double k0 = k[0] * t2, k1 = k[1] * t4, k2 = k[2] * t6, k3 = k[3] * t8;
            const double fix = (theta * (1 + k0 + k1 + k2 + k3) - theta_d) / (1 + 3 * k0 + 5 * k1 + 7 * k2 + 9 * k3);
            theta = theta - fix;
            if (std::fabs(fix) < eps) { converged = true; break; }
          }
          scale = std::tan(theta) / theta_d;
        } else {
          converged = true;
        }
        const bool flipped = (theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0);
        if (converged && !flipped) {
          const double ux = pwx * scale, uy = pwy * scale;
          const double r0 = RR[0] * ux + RR[1] * uy + RR[2], r1 = RR[3] * ux + RR[4] * uy + RR[5];
          const double r2 = RR[6] * ux + RR[7] * uy + RR[8];
          o[0] = (float)(r0 / r2); o[1] = (float)(r1 / r2);
        } else {
          o[0] = -1000000.f; o[1] = -1000000.f;
        }
      }
    }
  return 0;
}

size_t host_packetize(const EventTimes& ev, size_t n_ev, const emvs_stamped_pose* traj, size_t n_poses,
                      const emvs_pose& T_rv_w_pod, const emvs_camera& cam, const float virt[4], float z0,
                      emvs_packet* out, size_t max_out, bool* truncated)
{
  size_t cur = 0;
  return host_packetize_range(ev, n_ev, traj, n_poses, T_rv_w_pod, cam, virt, z0, &cur, n_ev, out, max_out, truncated);
}

namespace {

struct PacketStage {   // per-call constants of the packet loop
  Mat3f K, Kinv_virtual;
  SE3 T_rv_w;
  float z0;
};

// One iteration of mapper_emvs_stereo.cpp:91-120 for the packet that starts at event `cur`: false on a pose miss.
inline bool make_packet(const PacketStage& S, const EventTimes& ev, size_t cur, const emvs_stamped_pose* traj, size_t n_poses,
                        emvs_packet* pk)
{
  uint32_t mid_sec, mid_nsec;
  ev.at(cur + EMVS_PACKET_SIZE / 2, &mid_sec, &mid_nsec);
  SE3 T_w_ev;
  if (!interpolate(traj, n_poses, mid_sec, mid_nsec, &T_w_ev)) return false;
  const SE3 T_ev_rv = (S.T_rv_w * T_w_ev).inverse();
  double Rd[3][3];
  T_ev_rv.q.to_matrix(Rd);
  Mat3f R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R.m[i][j] = (float)Rd[i][j];
  const float t[3] = {(float)T_ev_rv.t.x, (float)T_ev_rv.t.y, (float)T_ev_rv.t.z};

  pk->first_event = cur;
  for (int i = 0; i < 3; ++i) {  // C = -R^T t
    const float s12 = (-R.m[1][i]) * t[1] + (-R.m[2][i]) * t[2];
    pk->C[i] = (-R.m[0][i]) * t[0] + s12;
  }
  Mat3f Hinv = R;  // (H_z0)^-1 = z0 R + t e3^T
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Hinv.m[i][j] *= S.z0;
  for (int i = 0; i < 3; ++i) Hinv.m[i][2] += t[i];
  const Mat3f H = inverse((S.K * Hinv) * S.Kinv_virtual);
  std::memcpy(pk->H, H.m, sizeof pk->H);
  return true;
}

unsigned packet_threads()
{
  static const unsigned n = [] {
    if (const char* e = std::getenv("EMVS_HOST_THREADS")) return (unsigned)std::max(1, std::atoi(e));
    const unsigned hw = std::thread::hardware_concurrency();
    return std::max(1u, std::min(4u, hw ? hw : 1u));   // measured: 1.48 -> 0.64 ms at 4 threads, no gain beyond (thread start-up)
  }();
  return n;
}

}  // namespace

// Resumable form: continues the packet loop of mapper_emvs_stereo.cpp:86-126 from *cur and stops before the first
// packet that would reach past `event_limit` (or past the list, strict '<' as :88).  Calling it with growing
// limits produces exactly the packets of one call over the whole list.
//
// The loop is sequential only through its cursor: a pose miss advances it by ONE event, a success by 1024.  Misses
// happen at the ends of a trajectory; in between every packet starts 1024 events after the previous one.  So after a
// streak of successes the remaining packets are computed speculatively at cur + 1024*j by a few threads, and the
// longest all-successful prefix is kept — exactly what the sequential loop would have produced; at the first miss
// the loop continues sequentially from there.  (One pose interpolation + two 3x3 inverses per packet: 1.4 ms per
// 5 M events on one core, which is what a prefetched or device-resident caller waits for.)
size_t host_packetize_range(const EventTimes& ev, size_t n_ev, const emvs_stamped_pose* traj, size_t n_poses,
                            const emvs_pose& T_rv_w_pod, const emvs_camera& cam, const float virt[4], float z0,
                            size_t* cur_inout, size_t event_limit, emvs_packet* out, size_t max_out, bool* truncated)
{
  PacketStage S;
  S.K = pinhole_K(cam.fx, cam.fy, cam.cx, cam.cy);
  S.Kinv_virtual = inverse(pinhole_K(virt[0], virt[1], virt[2], virt[3]));
  S.T_rv_w = load(T_rv_w_pod);
  S.z0 = z0;
  constexpr size_t kStreak = 4;
  // below this many remaining packets the thread start-up costs more than it saves (tests lower it to exercise the
  // speculative path on short lists)
  static const size_t kMinBatch = [] {
    const char* e = std::getenv("EMVS_PACKET_MIN_BATCH");
    return (size_t)std::max(1, e ? std::atoi(e) : 512);
  }();
  size_t produced = 0, streak = 0;
  size_t cur = *cur_inout;
  auto fits = [&](size_t c) { return c + EMVS_PACKET_SIZE < n_ev && c + EMVS_PACKET_SIZE <= event_limit; };
  while (fits(cur) && produced < max_out) {
    const unsigned T = packet_threads();
    if (streak >= kStreak && T > 1) {
      // packets that fit if no further miss occurs: cur + 1024*j for j < n_fit
      const size_t last = std::min(n_ev - 1, event_limit);   // need cur + 1024*(j+1) <= last (both bounds folded: '<' n_ev)
      size_t n_fit = (last - cur) / EMVS_PACKET_SIZE;
      n_fit = std::min(n_fit, max_out - produced);
      if (n_fit >= kMinBatch) {
        // This code is reached from extern "C" entry points that never throw: a failed allocation or thread start
        // (std::bad_alloc, std::system_error at the thread limit) must not cross the C ABI.  Workers that did start
        // are always joined; whatever could not be started is simply not done speculatively — the sequential loop
        // below recomputes it, so the result is the same.
        constexpr unsigned kMaxT = 64;
        const unsigned Tn = std::min(T, kMaxT);
        size_t first_miss[kMaxT];
        for (unsigned w = 0; w < Tn; ++w) first_miss[w] = n_fit * w / Tn;   // "nothing done" until the worker reports
        auto work = [&](unsigned w) {
          const size_t lo = n_fit * w / Tn, hi = n_fit * (w + 1) / Tn;
          size_t j = lo;
          for (; j < hi; ++j)
            if (!make_packet(S, ev, cur + j * EMVS_PACKET_SIZE, traj, n_poses, out + produced + j)) break;
          first_miss[w] = j;   // == hi when every packet of the range succeeded
        };
        std::thread pool[kMaxT];
        unsigned started = 1;   // worker 0 is this thread
        try {
          for (unsigned w = 1; w < Tn; ++w) {
            pool[w] = std::thread(work, w);
            started = w + 1;
          }
        } catch (...) {
          // fewer workers than planned: the ranges of the missing ones keep first_miss == their lower bound
        }
        work(0);
        for (unsigned w = 1; w < started; ++w) pool[w].join();
        // longest all-successful prefix: ranges are contiguous, a range counts fully only if it reached its end
        size_t prefix = 0;
        for (unsigned w = 0; w < Tn; ++w) {
          const size_t hi = n_fit * (w + 1) / Tn;
          prefix = first_miss[w];
          if (first_miss[w] < hi) break;
        }
        produced += prefix;
        cur += prefix * EMVS_PACKET_SIZE;
        streak = 0;       // either nothing fits any more, or the packet at `cur` missed / was not computed: go on one by one
        continue;
      }
    }
    if (make_packet(S, ev, cur, traj, n_poses, out + produced)) {
      ++produced;
      ++streak;
      cur += EMVS_PACKET_SIZE;
    } else {
      ++cur;  // drop one event and retry
      streak = 0;
    }
  }
  *cur_inout = cur;
  if (truncated) *truncated = produced == max_out && fits(cur);
  return produced;
}

}  // namespace emvs
