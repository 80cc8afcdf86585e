// Minimal stand-in for libvis/src/libvis/libvis.h: just the typedefs the adapter touches.
// TEST INFRASTRUCTURE (tests/test_adapter_syntax.py): lets the header-only C++ adapter be
// syntax- and type-checked in an image without Eigen / Sophus / Qt.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
namespace vis {
typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef size_t usize;
using std::shared_ptr;
struct Vec3u8 { u8 v[3]; };
}  // namespace vis
