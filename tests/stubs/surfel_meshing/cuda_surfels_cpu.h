// Stand-in for APP/cuda_surfels_cpu.h:40-124 (same members; the original only adds an Eigen include).
#pragma once
#include <mutex>
#include "libvis/libvis.h"
namespace vis {
struct CUDASurfelBuffersCPU {
  explicit CUDASurfelBuffersCPU(usize n)
      : surfel_x_buffer(new float[n]), surfel_y_buffer(new float[n]), surfel_z_buffer(new float[n]),
        surfel_radius_squared_buffer(new float[n]), surfel_normal_x_buffer(new float[n]),
        surfel_normal_y_buffer(new float[n]), surfel_normal_z_buffer(new float[n]),
        surfel_last_update_stamp_buffer(new u32[n]) {}
  u32 frame_index;
  usize surfel_count;
  float* surfel_x_buffer; float* surfel_y_buffer; float* surfel_z_buffer; float* surfel_radius_squared_buffer;
  float* surfel_normal_x_buffer; float* surfel_normal_y_buffer; float* surfel_normal_z_buffer;
  u32* surfel_last_update_stamp_buffer;
};
class CUDASurfelsCPU {
 public:
  explicit CUDASurfelsCPU(usize n) : write_buffers_(new CUDASurfelBuffersCPU(n)), read_buffers_(new CUDASurfelBuffersCPU(n)) {}
  void LockWriteBuffers() { lock_.lock(); }
  void UnlockWriteBuffers() { lock_.unlock(); }
  CUDASurfelBuffersCPU* write_buffers() { return write_buffers_; }
  const CUDASurfelBuffersCPU& read_buffers() const { return *read_buffers_; }
 private:
  std::mutex lock_;
  CUDASurfelBuffersCPU* write_buffers_;
  CUDASurfelBuffersCPU* read_buffers_;
};
}  // namespace vis
