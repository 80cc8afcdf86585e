// Stand-in for libvis/camera.h: PinholeCamera4f as far as the adapter uses it
// (libvis/src/libvis/camera.h:1606-1611: parameters() = {fx, fy, cx, cy}).
#pragma once
#include "libvis/libvis.h"
namespace vis {
class PinholeCamera4f {
 public:
  PinholeCamera4f(int width, int height, const float* parameters) : width_(width), height_(height) {
    for (int i = 0; i < 4; ++i) parameters_[i] = parameters[i];
  }
  int width() const { return width_; }
  int height() const { return height_; }
  const float* parameters() const { return parameters_; }
 private:
  int width_, height_;
  float parameters_[4];
};
}  // namespace vis
