// TEST SHIM (not minkindr): the accessors of kindr::minimal::QuatTransformation the adapters use — getRotation()
// .toImplementation() (an Eigen::Quaterniond: w(), x(), y(), z()) and getPosition() (an Eigen::Vector3d: operator[]).
#pragma once
namespace kindr {
namespace minimal {
struct ShimQuaternion {
  double w_, x_, y_, z_;
  double w() const { return w_; }
  double x() const { return x_; }
  double y() const { return y_; }
  double z() const { return z_; }
};
struct ShimRotation {
  ShimQuaternion q;
  const ShimQuaternion& toImplementation() const { return q; }
};
struct ShimVector3 {
  double v[3];
  double operator[](int i) const { return v[i]; }
};
class QuatTransformation {
 public:
  QuatTransformation(double w, double x, double y, double z, double tx, double ty, double tz) : r_{{w, x, y, z}}, p_{{tx, ty, tz}} {}
  const ShimRotation& getRotation() const { return r_; }
  const ShimVector3& getPosition() const { return p_; }

 private:
  ShimRotation r_;
  ShimVector3 p_;
};
}  // namespace minimal
}  // namespace kindr
