// surfel_b200_adapter.h — header-only drop-in for vis::CUDASurfelReconstruction.
//
// Same public surface as the reference class
// (applications/surfel_meshing/src/surfel_meshing/cuda_surfel_reconstruction.h:44-176); every
// method forwards to the C ABI of libsurfel_b200.so (surfel_b200.h). Include this header INSTEAD of
// "surfel_meshing/cuda_surfel_reconstruction.h" in main.cc and drop cuda_surfel_reconstruction.cc /
// cuda_surfel_reconstruction_kernels.{cc,cu} from the SurfelMeshing target.
// It needs the reference's own libvis headers (Eigen, Sophus) and therefore only compiles inside the
// reference's build tree; it is not built in this repository's image (no Eigen).
#pragma once

#include <cuda_runtime.h>

#include <libvis/camera.h>
#include <libvis/cuda/cuda_buffer.h>
#include <libvis/libvis.h>
#include <libvis/logging.h>
#include <libvis/sophus.h>

#include "surfel_b200.h"
#include "surfel_meshing/cuda_surfels_cpu.h"

namespace vis {

class SurfelMeshingRenderWindow;

class CUDASurfelReconstruction {
 public:
  // The three GL resources and the render window only serve UpdateVisualizationBuffers() and debug
  // displays (GUI); they are accepted and ignored.
  CUDASurfelReconstruction(usize max_surfel_count, const PinholeCamera4f& depth_camera,
                           cudaGraphicsResource_t /*vertex_buffer_resource*/,
                           cudaGraphicsResource_t /*neighbor_index_buffer_resource*/,
                           cudaGraphicsResource_t /*normal_vertex_buffer_resource*/,
                           const shared_ptr<SurfelMeshingRenderWindow>& /*render_window*/) {
    Check(sm_create(&handle_, max_surfel_count, depth_camera.width(), depth_camera.height(),
                    depth_camera.parameters()[0], depth_camera.parameters()[1], depth_camera.parameters()[2],
                    depth_camera.parameters()[3]));
  }
  ~CUDASurfelReconstruction() { sm_destroy(handle_); }
  CUDASurfelReconstruction(const CUDASurfelReconstruction&) = delete;
  CUDASurfelReconstruction& operator=(const CUDASurfelReconstruction&) = delete;

  void Integrate(cudaStream_t stream, u32 frame_index, float depth_scaling, CUDABuffer<u16>* depth_buffer,
                 const CUDABuffer<float2>& normals_buffer, const CUDABuffer<float>& radius_buffer,
                 const CUDABuffer<Vec3u8>& color_buffer, const SE3f& global_T_local, float sensor_noise_factor,
                 float max_surfel_confidence, float regularizer_weight, int regularization_frame_window_size,
                 bool do_blending, int measurement_blending_radius,
                 int regularization_iterations_per_integration_iteration,
                 float radius_factor_for_regularization_neighbors, float normal_compatibility_threshold_deg,
                 int surfel_integration_active_window_size) {
    sm_integrate_params p;
    p.depth_scaling = depth_scaling;
    p.sensor_noise_factor = sensor_noise_factor;
    p.max_surfel_confidence = max_surfel_confidence;
    p.regularizer_weight = regularizer_weight;
    p.regularization_frame_window_size = regularization_frame_window_size;
    p.do_blending = do_blending ? 1 : 0;
    p.measurement_blending_radius = measurement_blending_radius;
    p.regularization_iterations_per_integration_iteration = regularization_iterations_per_integration_iteration;
    p.radius_factor_for_regularization_neighbors = radius_factor_for_regularization_neighbors;
    p.normal_compatibility_threshold_deg = normal_compatibility_threshold_deg;
    p.surfel_integration_active_window_size = surfel_integration_active_window_size;
    float g[12], l[12];
    ToRowMajor(global_T_local, g);
    ToRowMajor(global_T_local.inverse(), l);  // as cuda_surfel_reconstruction.cc:144
    Check(sm_integrate(handle_, stream, frame_index, &p, depth_buffer->ToCUDA().address(),
                       depth_buffer->ToCUDA().pitch(),
                       reinterpret_cast<const float*>(normals_buffer.ToCUDA().address()),
                       normals_buffer.ToCUDA().pitch(), radius_buffer.ToCUDA().address(),
                       radius_buffer.ToCUDA().pitch(),
                       reinterpret_cast<const uint8_t*>(color_buffer.ToCUDA().address()),
                       color_buffer.ToCUDA().pitch(), g, l));
  }

  void Regularize(cudaStream_t stream, u32 frame_index, float regularizer_weight,
                  float radius_factor_for_regularization_neighbors, int regularization_frame_window_size) {
    Check(sm_regularize(handle_, stream, frame_index, regularizer_weight, radius_factor_for_regularization_neighbors,
                        regularization_frame_window_size));
  }

  // The "buffers" object must be locked with LockWriteBuffers() when this is called.
  void TransferAllToCPU(cudaStream_t stream, u32 frame_index, CUDASurfelsCPU* buffers) {
    CUDASurfelBuffersCPU* b = buffers->write_buffers();
    b->frame_index = frame_index;
    uint64_t count = 0;
    Check(sm_transfer_all_to_cpu(handle_, stream, frame_index, b->surfel_x_buffer, b->surfel_y_buffer,
                                 b->surfel_z_buffer, b->surfel_radius_squared_buffer, b->surfel_normal_x_buffer,
                                 b->surfel_normal_y_buffer, b->surfel_normal_z_buffer,
                                 b->surfel_last_update_stamp_buffer, &count));
    b->surfel_count = count;
  }

  // GUI only (CUDA-OpenGL interop): no-op in this drop-in.
  void UpdateVisualizationBuffers(cudaStream_t, u32, u32, u32, int, bool, bool, bool, bool) {}

  void ExportVertices(cudaStream_t stream, CUDABuffer<float>* position_buffer, CUDABuffer<u8>* color_buffer) {
    Check(sm_export_vertices(handle_, stream, position_buffer->ToCUDA().address(), color_buffer->ToCUDA().address()));
  }

  void GetTimings(float* data_association, float* surfel_merging, float* measurement_blending, float* integration,
                  float* neighbor_update, float* new_surfel_creation, float* regularization) {
    float t[7] = {0, 0, 0, 0, 0, 0, 0};
    sm_enable_timings(handle_, 1);  // takes effect from the next Integrate() on
    if (sm_get_timings(handle_, t) != SM_OK) { for (float& v : t) v = 0; }
    *data_association = t[0]; *surfel_merging = t[1]; *measurement_blending = t[2]; *integration = t[3];
    *neighbor_update = t[4]; *new_surfel_creation = t[5]; *regularization = t[6];
  }

  inline u32 surfel_count() const { u32 v = 0; Check(sm_surfel_count(handle_, &v)); return v; }
  inline u32 surfels_size() const { u32 v = 0; Check(sm_surfels_size(handle_, &v)); return v; }

 private:
  static void ToRowMajor(const SE3f& t, float* out) {
    const auto m = t.matrix3x4();
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) out[4 * r + c] = m(r, c);
  }
  // The reference aborts on any CUDA error (libvis/cuda/cuda_util.h:35-49).
  static void Check(int status) {
    if (status != SM_OK) LOG(FATAL) << "surfel_b200: " << sm_last_error();
  }
  sm_reconstruction* handle_ = nullptr;
};

}  // namespace vis
