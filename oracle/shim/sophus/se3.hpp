// Minimal stand-in for Sophus (b474f05) SE3Group<Scalar>: inverse, composition, accessors, as
// published (SO3 product followed by the first-order renormalisation of the unit quaternion).
// Test infrastructure only (oracle/_ref).
#pragma once
#include "Eigen/Core"

namespace Sophus {

template <typename T>
class SE3Group {
 public:
  typedef Eigen::Quaternion<T> Quat;
  typedef Eigen::Matrix<T, 3, 1> Vec;
  SE3Group() : q_(), t_() {}
  SE3Group(const Quat& q, const Vec& t) : q_(q), t_(t) {}
  const Quat& unit_quaternion() const { return q_; }
  const Vec& translation() const { return t_; }
  Vec& translation() { return t_; }
  SE3Group inverse() const {
    Quat qi = q_.conjugate();
    return SE3Group(qi, qi * (t_ * T(-1)));
  }
  SE3Group operator*(const SE3Group& o) const {
    SE3Group r(*this);
    r.t_ += q_ * o.t_;
    r.q_ *= o.q_;
    T n2 = r.q_.squaredNorm();
    if (n2 != T(1)) r.q_.coeffs() *= T(2) / (T(1) + n2);
    return r;
  }
  Vec operator*(const Vec& p) const { return q_ * p + t_; }

 private:
  Quat q_;
  Vec t_;
};
typedef SE3Group<float> SE3f;
typedef SE3Group<double> SE3d;

}  // namespace Sophus
