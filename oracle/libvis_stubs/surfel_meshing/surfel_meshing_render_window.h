// TEST INFRASTRUCTURE (oracle build aid): the debug render window of the reference (Qt + OpenGL) as a no-op; the CPU
// meshing code only calls it when a window was passed in (nullptr here).
#pragma once
#include <memory>
#include "libvis/eigen.h"
#include "libvis/mesh.h"
#include "libvis/point_cloud.h"
namespace vis {
class SurfelMeshingRenderWindow {
 public:
  template <class... Args> void CenterViewOn(Args&&...) {}
  template <class... Args> void UpdateVisualizationCloud(Args&&...) {}
  template <class... Args> void UpdateVisualizationMesh(Args&&...) {}
};
}  // namespace vis
