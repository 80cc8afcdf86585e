// Stand-in for libvis/logging.h (loguru streams): LOG(FATAL) << ... aborts.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
namespace vis_stub {
struct FatalStream {
  std::ostringstream s;
  template <typename T> FatalStream& operator<<(const T& v) { s << v; return *this; }
  ~FatalStream() { std::cerr << s.str() << std::endl; std::abort(); }
};
}  // namespace vis_stub
#define FATAL 0
#define LOG(level) ::vis_stub::FatalStream()
