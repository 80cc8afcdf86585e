// Stand-in for libvis/sophus.h: SE3f with inverse() and matrix3x4() (Sophus::SE3<float>).
#pragma once
namespace vis {
struct Matrix3x4f {
  float m[3][4];
  float operator()(int r, int c) const { return m[r][c]; }
};
class SE3f {
 public:
  SE3f() { for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) m_.m[r][c] = (r == c) ? 1.f : 0.f; }
  SE3f inverse() const {
    SE3f out;
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) out.m_.m[r][c] = m_.m[c][r];
      out.m_.m[r][3] = -(m_.m[0][r] * m_.m[0][3] + m_.m[1][r] * m_.m[1][3] + m_.m[2][r] * m_.m[2][3]);
    }
    return out;
  }
  Matrix3x4f matrix3x4() const { return m_; }
 private:
  Matrix3x4f m_;
};
}  // namespace vis
