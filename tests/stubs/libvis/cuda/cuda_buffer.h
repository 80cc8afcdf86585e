// Stand-in for libvis/cuda/cuda_buffer.h: CUDABuffer<T>::ToCUDA() -> CUDABuffer_<T>
// {address, height, width, pitch} (libvis/src/libvis/cuda/cuda_buffer.cuh:115-118).
#pragma once
#include <cstddef>
namespace vis {
template <typename T>
class CUDABuffer_ {
 public:
  CUDABuffer_(T* address, int height, int width, size_t pitch) : address_(address), height_(height), width_(width), pitch_(pitch) {}
  T* address() const { return address_; }
  int width() const { return width_; }
  int height() const { return height_; }
  size_t pitch() const { return pitch_; }
 private:
  T* address_; int height_; int width_; size_t pitch_;
};
template <typename T>
class CUDABuffer {
 public:
  CUDABuffer(int height, int width) : data_(nullptr, height, width, width * sizeof(T)) {}
  const CUDABuffer_<T>& ToCUDA() const { return data_; }
  CUDABuffer_<T>& ToCUDA() { return data_; }
 private:
  CUDABuffer_<T> data_;
};
}  // namespace vis
