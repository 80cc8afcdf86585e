// TEST INFRASTRUCTURE (oracle build aid): the two point types and the cloud container of libvis/point_cloud.h that
// surfel_meshing.{h,cc} name (positions + colours in std::vector), without the OpenGL / file-IO parts.
#pragma once
#include <vector>
#include "libvis/eigen.h"
namespace vis {
struct Point3f {
  Point3f() {}
  explicit Point3f(const Vec3f& p) : position_(p) {}
  Vec3f& position() { return position_; }
  const Vec3f& position() const { return position_; }
  Vec3f position_;
};
struct Point3fC3u8 {
  Point3fC3u8() {}
  Point3fC3u8(const Vec3f& p, const Vec3u8& c) : position_(p), color_(c) {}
  Vec3f& position() { return position_; }
  const Vec3f& position() const { return position_; }
  Vec3u8& color() { return color_; }
  const Vec3u8& color() const { return color_; }
  Vec3f position_;
  Vec3u8 color_;
};
template <class PointT>
class PointCloud {
 public:
  PointCloud() {}
  explicit PointCloud(usize size) : data_(size) {}
  void Resize(usize size) { data_.resize(size); }
  usize size() const { return data_.size(); }
  PointT& operator[](usize i) { return data_[i]; }
  const PointT& operator[](usize i) const { return data_[i]; }
  PointT& at(usize i) { return data_[i]; }
  const PointT& at(usize i) const { return data_[i]; }
  PointT* data_mutable() { return data_.data(); }
  const PointT* data() const { return data_.data(); }
 private:
  std::vector<PointT> data_;
};
typedef PointCloud<Point3f> Point3fCloud;
typedef PointCloud<Point3fC3u8> Point3fC3u8Cloud;
}  // namespace vis
