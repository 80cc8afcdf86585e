// TEST INFRASTRUCTURE (oracle build aid): the triangle-mesh container of libvis/mesh.h as far as
// SurfelMeshing::ConvertToMesh3fCu8 fills it (a vertex cloud + index triples).
#pragma once
#include <algorithm>
#include <memory>
#include <vector>
#include "libvis/eigen.h"
#include "libvis/point_cloud.h"
namespace vis {
template <class T>
struct Triangle {
  Triangle() {}
  Triangle(T a, T b, T c) : indices_{a, b, c} {}
  T& index(int i) { return indices_[i]; }
  const T& index(int i) const { return indices_[i]; }
  T indices_[3];
};
template <class PointT>
class Mesh3 {
 public:
  std::shared_ptr<PointCloud<PointT>>* vertices_mutable() { return &vertices_; }
  const std::shared_ptr<PointCloud<PointT>>& vertices() const { return vertices_; }
  std::vector<Triangle<u32>>* triangles_mutable() { return &triangles_; }
  const std::vector<Triangle<u32>>& triangles() const { return triangles_; }
 private:
  std::shared_ptr<PointCloud<PointT>> vertices_;
  std::vector<Triangle<u32>> triangles_;
};
typedef Mesh3<Point3fC3u8> Mesh3fCu8;
}  // namespace vis
