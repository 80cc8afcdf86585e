// Translation unit for tests/test_adapter_syntax.py: the calls APP/main.cc makes on
// vis::CUDASurfelReconstruction (:835-837, :1205-1223, :1261, :1323, :1576, :143, :1511), against
// include/surfel_b200_adapter.h and the stand-in libvis headers of this directory.
#include "surfel_b200_adapter.h"

namespace vis { class SurfelMeshingRenderWindow {}; }

int adapter_check_main() {
  using namespace vis;
  const float parameters[4] = {525.f, 525.f, 320.f, 240.f};
  PinholeCamera4f camera(640, 480, parameters);
  shared_ptr<SurfelMeshingRenderWindow> window;
  CUDASurfelReconstruction reconstruction(5000000, camera, nullptr, nullptr, nullptr, window);
  cudaStream_t stream = nullptr;
  CUDABuffer<u16> depth(480, 640);
  CUDABuffer<float2> normals(480, 640);
  CUDABuffer<float> radius(480, 640);
  CUDABuffer<Vec3u8> color(480, 640);
  SE3f global_T_frame;
  reconstruction.Integrate(stream, 7, 5000.f, &depth, normals, radius, color, global_T_frame, 0.05f, 5.f, 10.f, 30, true, 12,
                           1, 2.f, 40.f, 2147483647);
  reconstruction.Regularize(stream, 7, 10.f, 2.f, 30);
  CUDASurfelsCPU cpu_buffers(5000000);
  cpu_buffers.LockWriteBuffers();
  reconstruction.TransferAllToCPU(stream, 7, &cpu_buffers);
  cpu_buffers.UnlockWriteBuffers();
  reconstruction.UpdateVisualizationBuffers(stream, 7, 5, 1000, 2147483647, false, false, false, false);
  CUDABuffer<float> positions(1, 3 * 1000);
  CUDABuffer<u8> colors(1, 3 * 1000);
  reconstruction.ExportVertices(stream, &positions, &colors);
  float t[7];
  reconstruction.GetTimings(&t[0], &t[1], &t[2], &t[3], &t[4], &t[5], &t[6]);
  return static_cast<int>(reconstruction.surfel_count() + reconstruction.surfels_size());
}
