// preprocess.cu — depth pre-processing kernels for sm_100a.
//
// Replaces the five launches of APP/main.cc:1015-1191
//   BilateralFilteringAndDepthCutoffCUDA   APP/cuda_depth_processing.cu:50-158
//   OutlierDepthMapFusionCUDA<K+1,u16>     :168-285 (all inliers) / :337-457 (>= required)
//   ErodeDepthMapCUDA / CopyWithoutBorder  :514-579 / :589-633
//   ComputeNormalsAndDropBadPixelsCUDA     :642-762
//   ComputePointRadiiAndRemoveIsolatedPixelsCUDA :765-883
// by two fused kernels:
//   k_bilateral_outlier : raw u16 tile (+halo 6) staged in shared memory as fp32 with
//                         128-bit global loads, 113-tap bilateral with both reciprocals
//                         hoisted (1 MUFU per tap instead of the reference's 3), then the
//                         multi-frame outlier test on the filtered value in registers.
//   k_erode_normals_radii : erosion -> normals -> radii through three shared-memory tiles
//                         (halo 4/2/1), optionally also clearing the association rasters
//                         of the following Integrate().
// plus one plain kernel per reference stage (used by the link-level shims and by the
// per-stage parity tests). All arithmetic follows the reference's sm_100a SASS op for op
// (see sm_math.cuh); the u16 outputs are bit-exact.

#include <cuda.h>  // CUtensorMap

#include "sm_kernels.cuh"

namespace smb {

namespace {

constexpr int kTileW = 32;   // output tile: 32 x 8 pixels, one pixel per thread
constexpr int kTileH = 8;
constexpr float kLog2e = 1.4426950216293334961f;  // 0x3FB8AA3B, the constant nvcc emits for exp()

// ---------------------------------------------------------------------------------------
// a1: bilateral filter + depth cutoff (cuda_depth_processing.cu:50-118)
// ---------------------------------------------------------------------------------------

struct BilateralArgs {
  float denom_xy;             // 2 * sigma_xy^2
  float sigma_value_factor;
  int radius;
  int radius_squared;
  u16 value_to_ignore;
  u16 max_depth;
  float valid_radius_squared;
  int width, height;
  const u16* in;
  size_t in_pitch;
  unsigned long long* timeline;  // device timeline slot of this launch or null (diagnostics)
  int skip;                      // placeholder launch of the frame graph: return at once
};

// Cooperative fill of a (TH + 2R) x (tile_w_pad) fp32 tile from a pitched u16 raster.
// The tile starts at x0 = tile_x - kPadX (kPadX = 8 >= R keeps x0 a multiple of 8 pixels so
// that every 8-pixel group is one aligned 128-bit load); out-of-image pixels read as `fill`.
template <int R, int PADX, int SW, int TH>
__device__ __forceinline__ void load_depth_tile_f32(float* tile, const u16* in, size_t pitch, int width, int height,
                                                    int tile_x, int tile_y, float fill) {
  constexpr int kRows = TH + 2 * R;
  constexpr int kVecPerRow = SW / 8;
  const int x0 = tile_x - PADX;
  const int y0 = tile_y - R;
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) | pitch) & 15) == 0;
  for (int v = threadIdx.x; v < kRows * kVecPerRow; v += blockDim.x) {
    const int row = v / kVecPerRow;
    const int col = (v - row * kVecPerRow) * 8;
    const int gy = y0 + row;
    const int gx = x0 + col;
    float vals[8];
    if (gy >= 0 && gy < height && gx >= 0 && gx + 8 <= width && aligned) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(row_ptr(in, pitch, gy) + gx));
      vals[0] = u2f(q.x & 0xFFFFu); vals[1] = u2f(q.x >> 16);
      vals[2] = u2f(q.y & 0xFFFFu); vals[3] = u2f(q.y >> 16);
      vals[4] = u2f(q.z & 0xFFFFu); vals[5] = u2f(q.z >> 16);
      vals[6] = u2f(q.w & 0xFFFFu); vals[7] = u2f(q.w >> 16);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int xx = gx + i;
        vals[i] = (gy >= 0 && gy < height && xx >= 0 && xx < width) ? u2f(row_ptr(in, pitch, gy)[xx]) : fill;
      }
    }
    float4* dst = reinterpret_cast<float4*>(tile + row * SW + col);
    dst[0] = make_float4(vals[0], vals[1], vals[2], vals[3]);
    dst[1] = make_float4(vals[4], vals[5], vals[6], vals[7]);
  }
}

// One tap (cuda_depth_processing.cu:98-112; SASS: FMUL d*-d, FMUL *rcp_v, FFMA(-gd2, rcp_xy, .),
// FMUL *log2e, MUFU.EX2, FFMA sum, FADD weight). c - s is exact in fp32 (both < 2^16).
//
// kIgnoredTapsVanish: value_to_ignore is 0 and sigma_value_factor is small enough that an ignored
// sample (s = 0) gets the weight exp(-c^2 / (2 (c sigma)^2) - ...) = exp(< -110), which
// MUFU.EX2 with flush-to-zero returns as exactly +0: `sum += 0 * 0` and `weight += 0` are then
// exact no-ops and the `s != ignore` test of the reference (:103) can be dropped (one
// instruction per tap, 10 %). The host checks the precondition (LaunchBilateral).
template <bool kIgnoredTapsVanish = false>
__device__ __forceinline__ void bilateral_tap(float s, float c, float neg_gd2, float rcp_xy, float rcp_v, float ignore,
                                              float& sum, float& weight) {
  const float d = fsub(c, s);
  const float q = fmul(d, -d);
  const float t = fmul(q, rcp_v);
  const float e = ffma(neg_gd2, rcp_xy, t);
  const float w = fex2_approx(fmul(e, kLog2e));
  if (kIgnoredTapsVanish || s != ignore) {
    sum = ffma(w, s, sum);
    weight = fadd(weight, w);
  }
}

// Bilateral value of the pixel whose tile coordinates are (lx, ly) (tile origin includes
// the halo). Returns the u16 result. R > 0: compile-time disc, fully unrolled.
template <int R, int SW, bool kIgnoredTapsVanish>
__device__ __forceinline__ u16 bilateral_pixel(const float* tile, int lx, int ly, float c, const BilateralArgs& a,
                                               float rcp_xy) {
  const float ignore = u2f(a.value_to_ignore);
  const float av = fmul(c, a.sigma_value_factor);   // adapted_sigma_value
  const float denom_v = fmul(av, fadd(av, av));     // 2 * s * s as FADD + FMUL
  const float rcp_v = frcp(denom_v);
  float sum = 0.f, weight = 0.f;
#pragma unroll
  for (int dy = -R; dy <= R; ++dy) {
#pragma unroll
    for (int dx = -R; dx <= R; ++dx) {
      if (dx * dx + dy * dy <= R * R) {
        bilateral_tap<kIgnoredTapsVanish>(tile[(ly + dy) * SW + lx + dx], c, static_cast<float>(-(dx * dx + dy * dy)), rcp_xy, rcp_v,
                      ignore, sum, weight);
      }
    }
  }
  if (weight != 0.f) {
    return static_cast<u16>(f2u_trunc(ffma(frcp(weight), sum, 0.5f)));
  }
  return a.value_to_ignore;
}

// Circle mask + cutoff (cuda_depth_processing.cu:64-79). The reference evaluates the
// squared centre distance in 32-bit unsigned arithmetic and converts it with I2FP.F32.U32.
__device__ __forceinline__ bool bilateral_pixel_masked(unsigned x, unsigned y, const BilateralArgs& a) {
  const unsigned hx = x - static_cast<unsigned>(a.width / 2);
  const unsigned hy = y - static_cast<unsigned>(a.height / 2);
  const float center_distance_squared = u2f(hx * hx + hy * hy);
  return center_distance_squared > a.valid_radius_squared;
}

// ---------------------------------------------------------------------------------------
// a2: multi-frame outlier fusion (cuda_depth_processing.cu:168-227 / :337-397)
// ---------------------------------------------------------------------------------------

constexpr int kMaxOthers = 8;

struct OutlierArgs {
  int other_count;      // K
  int required_count;   // < 0: all K must agree (early break), else >= required_count
  float max_tolerance_factor, min_tolerance_factor;
  float fx, fy, cx, cy;
  float fx_inv, fy_inv, cx_inv, cy_inv;
  int width, height;
  Mat3x4 other_TR_reference[kMaxOthers];
  const u16* other_depths[kMaxOthers];
  size_t other_pitches[kMaxOthers];
};

// Returns the depth value to keep (depth_value or 0). The reference walks the other frames
// one after the other and stops at the first failure; here the K projections are computed
// first and the K depth gathers are issued together (same decisions, one memory round trip).
__device__ __forceinline__ u16 outlier_pixel(const OutlierArgs& a, unsigned x, unsigned y, u16 depth_value) {
  if (depth_value == 0) return 0;
  const float d = u2f(depth_value);
  const float px = fmul(ffma(a.fx_inv, u2f(x), a.cx_inv), d);
  const float py = fmul(ffma(a.fy_inv, u2f(y), a.cy_inv), d);
  float oz[kMaxOthers];
  const u16* sample[kMaxOthers];
#pragma unroll
  for (int k = 0; k < kMaxOthers; ++k) {
    sample[k] = nullptr;
    if (k >= a.other_count) continue;
    const Mat3x4& m = a.other_TR_reference[k];
    oz[k] = transform_row(m.r2, px, py, d);
    if (oz[k] <= 0.f) continue;
    const float ox = transform_row(m.r0, px, py, d);
    const float oy = transform_row(m.r1, px, py, d);
    const float inv_z = frcp(oz[k]);
    const int ix = f2i_trunc(ffma(fmul(ox, inv_z), a.fx, a.cx));
    const int iy = f2i_trunc(ffma(fmul(oy, inv_z), a.fy, a.cy));
    if (ix < 0 || iy < 0 || ix >= a.width || iy >= a.height) continue;
    sample[k] = row_ptr(a.other_depths[k], a.other_pitches[k], iy) + ix;
  }
  u16 od[kMaxOthers];
#pragma unroll
  for (int k = 0; k < kMaxOthers; ++k) od[k] = sample[k] ? __ldg(sample[k]) : static_cast<u16>(0);
  int ok_count = 0;
#pragma unroll
  for (int k = 0; k < kMaxOthers; ++k) {
    if (k >= a.other_count || od[k] == 0) continue;
    const float odf = u2f(od[k]);
    if (fmul(a.max_tolerance_factor, oz[k]) < odf) continue;
    if (fmul(a.min_tolerance_factor, oz[k]) > odf) continue;
    ++ok_count;
  }
  const int required = a.required_count < 0 ? a.other_count : a.required_count;
  return ok_count >= required ? depth_value : static_cast<u16>(0);
}

// ---------------------------------------------------------------------------------------
// fused a1+a2
// ---------------------------------------------------------------------------------------

// Tile height of the fused bilateral kernel: 32 x 4 pixels, 128 threads, 16 blocks per SM. One
// pixel per thread over 307200 pixels is 1.35 % more than the 148 x 2048 thread slots of the GPU,
// so some blocks always run in a second wave; with small blocks that tail is one short block
// instead of doubling the kernel time (measured with 32 x 8 tiles: 18 us vs an issue bound of 9).
constexpr int kBilateralTileH = 4;

template <int R, bool kWithOutlier, bool kIgnoredTapsVanish>
__global__ void __launch_bounds__(32 * kBilateralTileH, 2048 / (32 * kBilateralTileH))
k_bilateral_outlier(BilateralArgs a, const __grid_constant__ OutlierArgs o, u16* out, size_t out_pitch) {
  pdl_prologue();
  if (a.skip) return;
  const TimelineScope timeline_scope(a.timeline);
  constexpr int PADX = 8;
  constexpr int SW = kTileW + 2 * PADX;  // 48 floats per tile row
  __shared__ __align__(16) float tile[(kBilateralTileH + 2 * R) * SW];

  const int tile_x = blockIdx.x * kTileW;
  const int tile_y = blockIdx.y * kBilateralTileH;
  load_depth_tile_f32<R, PADX, SW, kBilateralTileH>(tile, a.in, a.in_pitch, a.width, a.height, tile_x, tile_y,
                                                    u2f(a.value_to_ignore));
  __syncthreads();

  const int tx = threadIdx.x & 31;
  const int ly = threadIdx.x >> 5;  // 0 .. kBilateralTileH - 1
  const unsigned x = tile_x + tx;
  const unsigned y = tile_y + ly;
  if (x >= static_cast<unsigned>(a.width) || y >= static_cast<unsigned>(a.height)) return;
  u16 result = a.value_to_ignore;
  if (!bilateral_pixel_masked(x, y, a)) {
    const float c = tile[(ly + R) * SW + tx + PADX];
    const unsigned ci = f2u_trunc(c);
    if (ci != a.value_to_ignore && ci <= a.max_depth) {
      result = bilateral_pixel<R, SW, kIgnoredTapsVanish>(tile, tx + PADX, ly + R, c, a, frcp(a.denom_xy));
    }
  }
  if (kWithOutlier) result = outlier_pixel(o, x, y, result);
  row_ptr(out, out_pitch, y)[x] = result;
}

// Generic-radius fallback (any radius): one thread per pixel, taps read through L1/L2.
__global__ void __launch_bounds__(256)
k_bilateral_generic(BilateralArgs a, u16* out, size_t out_pitch) {
  pdl_prologue();
  const unsigned x = blockIdx.x * 32 + (threadIdx.x & 31);
  const unsigned y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= static_cast<unsigned>(a.width) || y >= static_cast<unsigned>(a.height)) return;
  u16 result = a.value_to_ignore;
  if (!bilateral_pixel_masked(x, y, a)) {
    const u16 center = row_ptr(a.in, a.in_pitch, y)[x];
    if (center != a.value_to_ignore && center <= a.max_depth) {
      const float c = u2f(center);
      const float ignore = u2f(a.value_to_ignore);
      const float rcp_xy = frcp(a.denom_xy);
      const float av = fmul(c, a.sigma_value_factor);
      const float rcp_v = frcp(fmul(av, fadd(av, av)));
      float sum = 0.f, weight = 0.f;
      const int min_y = max(0, static_cast<int>(y) - a.radius);
      const int max_y = min(a.height - 1, static_cast<int>(y) + a.radius);
      const int min_x = max(0, static_cast<int>(x) - a.radius);
      const int max_x = min(a.width - 1, static_cast<int>(x) + a.radius);
      for (int sy = min_y; sy <= max_y; ++sy) {
        const int dy = sy - static_cast<int>(y);
        const u16* row = row_ptr(a.in, a.in_pitch, sy);
        for (int sx = min_x; sx <= max_x; ++sx) {
          const int dx = sx - static_cast<int>(x);
          const int gd2 = dx * dx + dy * dy;
          if (gd2 > a.radius_squared) continue;
          bilateral_tap(u2f(row[sx]), c, i2f(-gd2), rcp_xy, rcp_v, ignore, sum, weight);
        }
      }
      if (weight != 0.f) result = static_cast<u16>(f2u_trunc(ffma(frcp(weight), sum, 0.5f)));
    }
  }
  row_ptr(out, out_pitch, y)[x] = result;
}

__global__ void __launch_bounds__(256)
k_outlier(const __grid_constant__ OutlierArgs o, const u16* in, size_t in_pitch, u16* out, size_t out_pitch) {
  pdl_prologue();
  const unsigned x = blockIdx.x * 32 + (threadIdx.x & 31);
  const unsigned y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= static_cast<unsigned>(o.width) || y >= static_cast<unsigned>(o.height)) return;
  row_ptr(out, out_pitch, y)[x] = outlier_pixel(o, x, y, row_ptr(in, in_pitch, y)[x]);
}

// ---------------------------------------------------------------------------------------
// a3: erosion / border copy (cuda_depth_processing.cu:514-538 / :589-607)
// ---------------------------------------------------------------------------------------

// `get(y, x)` must return 0 outside the image.
template <typename Get>
__device__ __forceinline__ u16 erode_pixel(int radius, int x, int y, int width, int height, Get get) {
  if (radius == 0) {
    constexpr int kBorderSize = 1;
    if (x < kBorderSize || y < kBorderSize || x >= width - kBorderSize || y >= height - kBorderSize) return 0;
    return get(y, x);
  }
  if (x < radius || y < radius || x >= width - radius || y >= height - radius) return 0;
  bool all_valid = true;
  for (int dy = y - radius; dy <= y + radius; ++dy) {
    for (int dx = x - radius; dx <= x + radius; ++dx) {
      if (get(dy, dx) == 0) all_valid = false;
    }
  }
  return all_valid ? get(y, x) : static_cast<u16>(0);
}

// ---------------------------------------------------------------------------------------
// a4: normals (cuda_depth_processing.cu:642-718)
// ---------------------------------------------------------------------------------------

struct NormalsArgs {
  float normal_dot_threshold;   // -cosf(M_PI / 180.f * angle)
  float inv_depth_scaling;
  float fx_inv, fy_inv, cx_inv, cy_inv;
};

// Precondition: centre and the four neighbours are non-zero. Returns the depth to keep.
__device__ __forceinline__ u16 normals_pixel(const NormalsArgs& a, int x, int y, u16 center, u16 left, u16 right,
                                             u16 top, u16 bottom, float2* normal_out) {
  const float ids = a.inv_depth_scaling;
  const float ld = fmul(u2f(left), ids);
  const float bd = fmul(u2f(bottom), ids);
  const float rd = fmul(u2f(right), ids);
  const float td = fmul(u2f(top), ids);
  const float fx_x = ffma(i2f(x), a.fx_inv, a.cx_inv);
  const float fx_xp1 = ffma(i2f(x + 1), a.fx_inv, a.cx_inv);
  const float fx_xm1 = ffma(i2f(x - 1), a.fx_inv, a.cx_inv);
  const float fy_y = ffma(i2f(y), a.fy_inv, a.cy_inv);
  const float fy_yp1 = ffma(i2f(y + 1), a.fy_inv, a.cy_inv);
  const float fy_ym1 = ffma(i2f(y - 1), a.fy_inv, a.cy_inv);
  const float left_x = fmul(ld, fx_xm1);
  const float left_y = fmul(ld, fy_y);
  const float bottom_x = fmul(fx_x, bd);
  const float bottom_y = fmul(bd, fy_yp1);
  // left_to_right = right - left, bottom_to_top = top - bottom; ptxas fuses the second
  // product of every difference into the subtraction.
  const float ax = ffma(rd, fx_xp1, -left_x);
  const float ay = ffma(fy_y, rd, -left_y);
  const float az = fsub(rd, ld);
  const float bx = ffma(td, fx_x, -bottom_x);
  const float by = ffma(td, fy_ym1, -bottom_y);
  const float bz = fsub(td, bd);
  // CrossProduct (cuda_util.cuh:69-73).
  float nx = ffma(ay, bz, -fmul(az, by));
  float ny = ffma(az, bx, -fmul(ax, bz));
  float nz = ffma(ax, by, -fmul(ay, bx));
  const float length = fsqrt_approx(ffma(nz, nz, ffma(nx, nx, fmul(ny, ny))));
  // Viewing direction (normalised with MUFU.RSQ).
  const float inv_dir_length = frsqrt_approx(fadd(ffma(fx_x, fx_x, fmul(fy_y, fy_y)), 1.0f));
  const float vy = fmul(fy_y, inv_dir_length);
  const float vx = fmul(fx_x, inv_dir_length);
  if (length > 1e-6f) {
    const float inv_length = fmul((a.fy_inv < 0.f) ? -1.0f : 1.0f, frcp(length));
    nx = fmul(nx, inv_length);
    ny = fmul(ny, inv_length);
    nz = fmul(nz, inv_length);
  } else {
    nx = 0.f; ny = 0.f; nz = -1.f;
  }
  *normal_out = make_float2(nx, ny);
  const float dot = ffma(inv_dir_length, nz, ffma(vx, nx, fmul(vy, ny)));
  return (dot >= a.normal_dot_threshold) ? static_cast<u16>(0) : center;
}

// ---------------------------------------------------------------------------------------
// a5: radii (cuda_depth_processing.cu:765-837)
// ---------------------------------------------------------------------------------------

struct RadiiArgs {
  float point_radius_extension_factor_squared;
  float clamp_factor_term;
  float inv_depth_scaling;
  float fx_inv, fy_inv, cx_inv, cy_inv;
};

// `get(y, x)`: normals-stage depth. Precondition: centre non-zero. Returns kept depth.
template <typename Get>
__device__ __forceinline__ u16 radii_pixel(const RadiiArgs& a, int x, int y, u16 center, Get get,
                                           float* radius_squared_out) {
  const float ids = a.inv_depth_scaling;
  const float depth = fmul(u2f(center), ids);
  const float local_x = fmul(depth, ffma(u2f(x), a.fx_inv, a.cx_inv));
  const float local_y = fmul(depth, ffma(u2f(y), a.fy_inv, a.cy_inv));
  int neighbor_count = 0;
  float radius_squared = 0.f;
  float min_neighbor_distance_squared = __int_as_float(0x7f800000);
#pragma unroll
  for (int dy = y - 1; dy <= y + 1; ++dy) {
#pragma unroll
    for (int dx = x - 1; dx <= x + 1; ++dx) {
      if (dx == x && dy == y) continue;
      const float ddepth = fmul(u2f(get(dy, dx)), ids);
      if (ddepth <= 0.f) continue;
      ++neighbor_count;
      const float oy = ffma(ffma(i2f(dy), a.fy_inv, a.cy_inv), ddepth, -local_y);
      const float ox = ffma(ddepth, ffma(i2f(dx), a.fx_inv, a.cx_inv), -local_x);
      const float oz = fsub(ddepth, depth);
      const float distance_squared = ffma(oz, oz, ffma(ox, ox, fmul(oy, oy)));
      if (distance_squared > radius_squared) radius_squared = distance_squared;
      if (distance_squared < min_neighbor_distance_squared) min_neighbor_distance_squared = distance_squared;
    }
  }
  radius_squared = fmul(radius_squared, a.point_radius_extension_factor_squared);
  const float distance_squared_clamp = fmul(min_neighbor_distance_squared, a.clamp_factor_term);
  if (radius_squared > distance_squared_clamp) radius_squared = distance_squared_clamp;
  *radius_squared_out = radius_squared;
  constexpr int kMinNeighborPixelsForRadiusComputation = 8;
  return (neighbor_count < kMinNeighborPixelsForRadiusComputation) ? static_cast<u16>(0) : center;
}

// ---------------------------------------------------------------------------------------
// fused a3+a4+a5 (+ clear of the association rasters, a6)
// ---------------------------------------------------------------------------------------

struct TailArgs {
  int width, height;
  int erosion_radius;
  NormalsArgs normals;
  RadiiArgs radii;
  const u16* in; size_t in_pitch;         // outlier-filtered depth (B)
  u16* out_depth; size_t out_depth_pitch;  // final depth (A)
  u16* out_depth_copy; size_t out_depth_copy_pitch;  // optional second copy (pre-blend depth of the pipeline)
  float2* out_normals; size_t out_normals_pitch;
  float* out_radius; size_t out_radius_pitch;
  // Optional: association rasters to reset for the following Integrate().
  uint4* assoc; float* first_depth; u8* supported;
  unsigned long long* timeline;  // device timeline slot of this launch or null (diagnostics)
  int skip;                      // placeholder launch of the frame graph: return at once
};

constexpr int kMaxErode = 3;

// TMA (cp.async.bulk.tensor) + mbarrier primitives for the tile fill below.
__device__ __forceinline__ u32 smem_u32(const void* p) { return static_cast<u32>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbarrier_init(unsigned long long* bar, u32 arrive_count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrive_count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // make the init visible to the async proxy
}
__device__ __forceinline__ void mbarrier_arrive_expect_tx(unsigned long long* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a transfer that never completes (it cannot with a valid descriptor) traps instead of
// hanging the GPU.
__device__ __forceinline__ void mbarrier_wait(unsigned long long* bar, u32 phase) {
  for (u32 attempt = 0; attempt < (1u << 22); ++attempt) {
    u32 done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    if (done) return;
  }
  __trap();
}
// 2-D tiled TMA load: box of the tensor map at element coordinates (x, y) -> dense shared-memory tile;
// elements outside the tensor arrive as zeros.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
               : "memory");
}

// Output tile 32 x 16 pixels, 256 threads (two output pixels per thread; the halo stages stride over
// the block), 8 blocks per SM: all 600 blocks of a VGA frame are resident at once. Halo recompute per
// output pixel: B 2.4 / E 1.4 / N 1.2 loads-or-pixels (32 x 8 tiles of round 1: 3.0 / 1.7 / 1.3).
constexpr int kTailTileH = 16;
constexpr int kTailHaloY = kMaxErode + 2;   // B tile rows above / below the output tile
constexpr int kTailHaloX = 8;               // >= kMaxErode + 2; 8 keeps every 8-pixel group 16-byte aligned
constexpr int kTailBW = kTileW + 2 * kTailHaloX;          // 48 pixels = 96 bytes per tile row
constexpr int kTailBH = kTailTileH + 2 * kTailHaloY;      // 26 rows
enum { kFillVector = 0, kFillTma = 1 };

// kFill selects how the outlier-filtered input tile (+ halo) reaches shared memory:
//   kFillTma    one thread issues ONE 2-D TMA box load (cp.async.bulk.tensor, 48 x 26 u16 = 2496 bytes,
//               out-of-image elements zero-filled by the hardware) and the block waits on an mbarrier:
//               no per-element index math or bounds tests at all;
//   kFillVector cooperative 128-bit loads (one aligned 8-pixel group per thread, scalar + bounds-checked
//               only at the image border).
// (Round 1 filled a 42 x 18 tile with scalar, bounds-checked u16 loads.)
template <int kFill>
__global__ void __launch_bounds__(256, 8)
k_erode_normals_radii(TailArgs a, const __grid_constant__ CUtensorMap in_map) {
  pdl_prologue();
  if (a.skip) return;
  const TimelineScope timeline_scope(a.timeline);
  // Tiles (origin relative to the 32 x 16 output tile): B (outlier-filtered input) (-8, -5), HV
  // (row-wise erosion validity) -2 / -(2 + r), E (eroded) -2, N (normals stage) -1.
  constexpr int TH = kTailTileH, HB = kTailHaloY, HBX = kTailHaloX, BW = kTailBW, BH = kTailBH;
  constexpr int EW = kTileW + 4, EH = TH + 4;                 // 36 x 20
  constexpr int HVH = EH + 2 * kMaxErode;                     // 26 rows
  constexpr int NW = kTileW + 2, NH = TH + 2;                 // 34 x 18
  __shared__ __align__(128) u16 sB[BH * BW];
  __shared__ __align__(8) unsigned long long fill_barrier;
  __shared__ u8 sHV[HVH * EW];
  __shared__ u16 sE[EH * EW];
  __shared__ u16 sN[NH * NW];

  const int r = a.erosion_radius;
  const int tile_x = blockIdx.x * kTileW;
  const int tile_y = blockIdx.y * TH;

  if (kFill == kFillTma) {
    if (threadIdx.x == 0) mbarrier_init(&fill_barrier, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbarrier_arrive_expect_tx(&fill_barrier, BH * BW * sizeof(u16));
      tma_load_2d(sB, &in_map, tile_x - HBX, tile_y - HB, &fill_barrier);
    }
    mbarrier_wait(&fill_barrier, 0);
  } else {
    const bool aligned = ((reinterpret_cast<uintptr_t>(a.in) | a.in_pitch) & 15) == 0;
    constexpr int kVecPerRow = BW / 8;
    for (int v = threadIdx.x; v < BH * kVecPerRow; v += 256) {
      const int ly = v / kVecPerRow, lx = (v - ly * kVecPerRow) * 8;
      const int gx = tile_x - HBX + lx, gy = tile_y - HB + ly;
      uint4 q = make_uint4(0u, 0u, 0u, 0u);
      if (gy >= 0 && gy < a.height) {
        if (aligned && gx >= 0 && gx + 8 <= a.width) {
          q = __ldg(reinterpret_cast<const uint4*>(row_ptr(a.in, a.in_pitch, gy) + gx));
        } else {
          union { uint4 v; u16 e[8]; } t;
          t.v = q;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (gx + k >= 0 && gx + k < a.width) t.e[k] = row_ptr(a.in, a.in_pitch, gy)[gx + k];
          q = t.v;
        }
      }
      *reinterpret_cast<uint4*>(&sB[ly * BW + lx]) = q;
    }
    __syncthreads();
  }

  // Erosion (cuda_depth_processing.cu:514-538), separable: row-wise validity, then columns.
  // E (ex, ey) is image pixel (tile_x - 2 + ex, tile_y - 2 + ey) = sB[(ey - 2 + HB) * BW + ex - 2 + HBX].
  if (r > 0) {
    for (int i = threadIdx.x; i < (EH + 2 * r) * EW; i += 256) {
      const int hy = i / EW, ex = i - hy * EW;
      const u16* row = &sB[(hy - r + HB - 2) * BW + ex + HBX - 2];
      bool valid = true;
      for (int dx = -r; dx <= r; ++dx) valid &= row[dx] != 0;
      sHV[i] = valid;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < EW * EH; i += 256) {
    const int ey = i / EW, ex = i - ey * EW;
    const u16 center = sB[(ey + HB - 2) * BW + ex + HBX - 2];
    u16 v = 0;
    if (center != 0) {
      if (r > 0) {
        bool valid = true;
        for (int dy = 0; dy <= 2 * r; ++dy) valid &= sHV[(ey + dy) * EW + ex] != 0;
        v = valid ? center : static_cast<u16>(0);
      } else {
        // CopyWithoutBorderCUDAKernel (:589-607): 1-pixel border zeroed
        const int gx = tile_x - 2 + ex, gy = tile_y - 2 + ey;
        v = (gx < 1 || gy < 1 || gx >= a.width - 1 || gy >= a.height - 1) ? static_cast<u16>(0) : center;
      }
    }
    sE[i] = v;
  }
  __syncthreads();

  // Normals stage on the N tile (image coordinates tile - 1 .. tile + 1).
  for (int i = threadIdx.x; i < NW * NH; i += 256) {
    const int ly = i / NW, lx = i - ly * NW;
    const int gx = tile_x - 1 + lx, gy = tile_y - 1 + ly;
    const int e = (ly + 1) * EW + lx + 1;  // position in the E tile
    const u16 center = sE[e];
    u16 v = 0;
    float2 normal = make_float2(0.f, 0.f);
    if (center != 0) {
      // The reference reads the four neighbours without bounds checks and relies on the
      // zero border left by the erosion (cuda_depth_processing.cuh:96-98); the E tile is
      // zero outside the image, which is the same thing.
      const u16 right = sE[e + 1], left = sE[e - 1], bottom = sE[e + EW], top = sE[e - EW];
      if (right != 0 && left != 0 && bottom != 0 && top != 0) {
        v = normals_pixel(a.normals, gx, gy, center, left, right, top, bottom, &normal);
      }
    }
    const bool interior = lx >= 1 && lx <= kTileW && ly >= 1 && ly <= TH;
    if (interior && gx < a.width && gy < a.height) row_ptr(a.out_normals, a.out_normals_pitch, gy)[gx] = normal;
    sN[i] = v;
  }
  __syncthreads();

  // Radii on the tile interior: two output pixels per thread.
#pragma unroll
  for (int half = 0; half < TH / 8; ++half) {
    const int lx = threadIdx.x & 31, ly = (threadIdx.x >> 5) + 8 * half;
    const int gx = tile_x + lx, gy = tile_y + ly;
    if (gx < a.width && gy < a.height) {
      const u16 center = sN[(ly + 1) * NW + lx + 1];
      u16 kept = 0;
      if (center != 0) {
        auto get = [&](int yy, int xx) -> u16 { return sN[(yy - tile_y + 1) * NW + (xx - tile_x + 1)]; };
        float radius_squared;
        kept = radii_pixel(a.radii, gx, gy, center, get, &radius_squared);
        row_ptr(a.out_radius, a.out_radius_pitch, gy)[gx] = radius_squared;
      }
      row_ptr(a.out_depth, a.out_depth_pitch, gy)[gx] = kept;
      if (a.out_depth_copy) row_ptr(a.out_depth_copy, a.out_depth_copy_pitch, gy)[gx] = kept;
      if (a.assoc) {
        const size_t p = static_cast<size_t>(gy) * a.width + gx;
        a.assoc[p] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
        a.first_depth[p] = __int_as_float(0x7f800000);
        a.supported[p] = 0;
      }
    }
  }
}

// ---- plain per-stage kernels (one thread per pixel) ------------------------------------

__global__ void __launch_bounds__(256)
k_erode(int radius, int width, int height, const u16* in, size_t in_pitch, u16* out, size_t out_pitch) {
  pdl_prologue();
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= width || y >= height) return;
  auto get = [&](int yy, int xx) -> u16 { return row_ptr(in, in_pitch, yy)[xx]; };
  row_ptr(out, out_pitch, y)[x] = erode_pixel(radius, x, y, width, height, get);
}

__global__ void __launch_bounds__(256)
k_normals(NormalsArgs a, int width, int height, const u16* in, size_t in_pitch, u16* out, size_t out_pitch,
          float2* normals, size_t normals_pitch) {
  pdl_prologue();
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= width || y >= height) return;
  const u16 center = row_ptr(in, in_pitch, y)[x];
  u16 kept = 0;
  float2 normal = make_float2(0.f, 0.f);
  // Unlike the reference this kernel stays in bounds when the caller did not zero the border.
  if (center != 0 && x >= 1 && y >= 1 && x < width - 1 && y < height - 1) {
    const u16 right = row_ptr(in, in_pitch, y)[x + 1];
    const u16 left = row_ptr(in, in_pitch, y)[x - 1];
    const u16 bottom = row_ptr(in, in_pitch, y + 1)[x];
    const u16 top = row_ptr(in, in_pitch, y - 1)[x];
    if (right != 0 && left != 0 && bottom != 0 && top != 0) {
      kept = normals_pixel(a, x, y, center, left, right, top, bottom, &normal);
    }
  }
  row_ptr(out, out_pitch, y)[x] = kept;
  row_ptr(normals, normals_pitch, y)[x] = normal;
}

__global__ void __launch_bounds__(256)
k_radii(RadiiArgs a, int width, int height, const u16* in, size_t in_pitch, float* radius, size_t radius_pitch,
        u16* out, size_t out_pitch) {
  pdl_prologue();
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= width || y >= height) return;
  const u16 center = row_ptr(in, in_pitch, y)[x];
  u16 kept = 0;
  if (center != 0) {
    auto get = [&](int yy, int xx) -> u16 {
      return (xx >= 0 && yy >= 0 && xx < width && yy < height) ? row_ptr(in, in_pitch, yy)[xx] : static_cast<u16>(0);
    };
    float radius_squared;
    kept = radii_pixel(a, x, y, center, get, &radius_squared);
    row_ptr(radius, radius_pitch, y)[x] = radius_squared;
  }
  row_ptr(out, out_pitch, y)[x] = kept;
}

// ---------------------------------------------------------------------------------------
// f2: MedianFilterAndDensifyDepthMap (APP/main.cc:207-252), one iteration per launch
// ---------------------------------------------------------------------------------------
// The reference runs this on the CPU inside its upload loop (main.cc:927-939, "TODO: Do this on the
// GPU"): 3x3 window clipped to the image, zeros excluded; with >= 2 valid values the output is
// their median - for an even count the middle element closer to the float average (IEEE division,
// host code is not fast-math), the upper one on a tie - otherwise the input pixel. Instead of
// sorting, every valid value gets its rank (ties by window position), which selects the same
// elements.
__global__ void __launch_bounds__(256)
k_median_densify(int width, int height, const u16* in, size_t in_pitch, u16* out, size_t out_pitch) {
  pdl_prologue();
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= width || y >= height) return;
  u32 v[9];
  int n = 0;
  u32 sum = 0;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int yy = y + dy, xx = x + dx;
      u32 value = 0;
      if (yy >= 0 && yy < height && xx >= 0 && xx < width) value = row_ptr(in, in_pitch, yy)[xx];
      v[(dy + 1) * 3 + dx + 1] = value;
      n += value != 0 ? 1 : 0;
      sum += value;
    }
  }
  u16 result = static_cast<u16>(v[4]);
  if (n >= 2) {
    u32 lower = 0, upper = 0;  // sorted[n / 2 - 1], sorted[n / 2]
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      if (v[i] == 0) continue;
      int rank = 0;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        if (j == i || v[j] == 0) continue;
        rank += (v[j] < v[i] || (v[j] == v[i] && j < i)) ? 1 : 0;
      }
      if (rank == n / 2 - 1) lower = v[i];
      if (rank == n / 2) upper = v[i];
    }
    if (n % 2 == 0) {
      const float average = __fdiv_rn(__uint2float_rn(sum), __int2float_rn(n));  // sum <= 8 * 65535: exact
      const float prev_diff = fabsf(__fsub_rn(__uint2float_rn(lower), average));
      const float next_diff = fabsf(__fsub_rn(__uint2float_rn(upper), average));
      result = static_cast<u16>(prev_diff < next_diff ? lower : upper);
    } else {
      result = static_cast<u16>(upper);
    }
  }
  row_ptr(out, out_pitch, y)[x] = result;
}

// ---- host-side argument construction (mirrors the reference's host wrappers) ------------

BilateralArgs MakeBilateralArgs(float sigma_xy, float sigma_value_factor, u16 value_to_ignore, float radius_factor,
                                u16 max_depth, float depth_valid_region_radius, int width, int height, const u16* in,
                                size_t in_pitch) {
  BilateralArgs a;
  const int radius = radius_factor * sigma_xy + 0.5f;  // cuda_depth_processing.cu:135
  a.denom_xy = 2.0f * sigma_xy * sigma_xy;             // :145
  a.sigma_value_factor = sigma_value_factor;
  a.radius = radius;
  a.radius_squared = radius * radius;
  a.value_to_ignore = value_to_ignore;
  a.max_depth = max_depth;
  a.valid_radius_squared = depth_valid_region_radius * depth_valid_region_radius;  // :151
  a.width = width; a.height = height;
  a.in = in; a.in_pitch = in_pitch;
  a.timeline = nullptr;
  a.skip = 0;
  return a;
}

void MakeUnprojection(float fx, float fy, float cx, float cy, float* fx_inv, float* fy_inv, float* cx_inv,
                      float* cy_inv) {
  // Unprojection intrinsics for pixel center convention (cuda_depth_processing.cu:258-264).
  *fx_inv = 1.0f / fx;
  *fy_inv = 1.0f / fy;
  const float cx_pixel_center = cx - 0.5f;
  const float cy_pixel_center = cy - 0.5f;
  *cx_inv = -cx_pixel_center / fx;
  *cy_inv = -cy_pixel_center / fy;
}

int MakeOutlierArgs(OutlierArgs* o, int other_count, int required_count, float tolerance, float fx, float fy,
                    float cx, float cy, int width, int height, const u16* const* other_depths,
                    const size_t* other_pitches, const float* others_TR_reference) {
  if (other_count != 2 && other_count != 4 && other_count != 6 && other_count != 8) {
    return SetError(SM_ERR_INVALID_ARGUMENT, "Unsupported value for outlier_filtering_frame_count (2,4,6,8)");
  }
  o->other_count = other_count;
  o->required_count = (required_count == -1 || required_count == other_count) ? -1 : required_count;  // main.cc:1061
  o->max_tolerance_factor = 1 + tolerance;  // cuda_depth_processing.cu:255-256
  o->min_tolerance_factor = 1 - tolerance;
  o->fx = fx; o->fy = fy; o->cx = cx; o->cy = cy;
  MakeUnprojection(fx, fy, cx, cy, &o->fx_inv, &o->fy_inv, &o->cx_inv, &o->cy_inv);
  o->width = width; o->height = height;
  for (int i = 0; i < other_count; ++i) {
    o->other_TR_reference[i] = MakeMat3x4(others_TR_reference + 12 * i);
    o->other_depths[i] = other_depths[i];
    o->other_pitches[i] = other_pitches[i];
  }
  return SM_OK;
}

NormalsArgs MakeNormalsArgs(float observation_angle_threshold_deg, float depth_scaling, float fx, float fy, float cx,
                            float cy) {
  NormalsArgs n;
  n.normal_dot_threshold = -1 * cosf(M_PI / 180.f * observation_angle_threshold_deg);  // :752
  n.inv_depth_scaling = 1.0f / depth_scaling;
  MakeUnprojection(fx, fy, cx, cy, &n.fx_inv, &n.fy_inv, &n.cx_inv, &n.cy_inv);
  return n;
}

RadiiArgs MakeRadiiArgs(float point_radius_extension_factor, float point_radius_clamp_factor, float depth_scaling,
                        float fx, float fy, float cx, float cy) {
  RadiiArgs r;
  r.point_radius_extension_factor_squared = point_radius_extension_factor * point_radius_extension_factor;  // :872
  r.clamp_factor_term = point_radius_clamp_factor * point_radius_clamp_factor * sqrtf(2) * sqrtf(2);       // :873
  r.inv_depth_scaling = 1.0f / depth_scaling;
  MakeUnprojection(fx, fy, cx, cy, &r.fx_inv, &r.fy_inv, &r.cx_inv, &r.cy_inv);
  return r;
}

dim3 TailGrid(int width, int height) { return dim3((width + kTileW - 1) / kTileW, (height + kTailTileH - 1) / kTailTileH); }
dim3 PixelGrid(int width, int height) { return dim3((width + 31) / 32, (height + 7) / 8); }

// The fused a1 + a2 launch (radius 6) as a descriptor.
void DescribeBilateralOutlier(KernelLaunch* k, const BilateralArgs& a, const OutlierArgs* o, u16* out, size_t out_pitch) {
  static const OutlierArgs kNoOutlier = {};
  const dim3 grid((a.width + kTileW - 1) / kTileW, (a.height + kBilateralTileH - 1) / kBilateralTileH);
  const dim3 block(32 * kBilateralTileH);
  // ignored samples get weight exp(-1 / (2 sigma^2) - ...): exactly 0 under ftz once 1 / (2 sigma^2) > 110
  const bool vanish = a.value_to_ignore == 0 && a.sigma_value_factor > 0.f &&
                      1.0f / (2.0f * a.sigma_value_factor * a.sigma_value_factor) > 110.0f;
  const void* func;
  if (o) func = vanish ? reinterpret_cast<const void*>(k_bilateral_outlier<6, true, true>)
                       : reinterpret_cast<const void*>(k_bilateral_outlier<6, true, false>);
  else func = vanish ? reinterpret_cast<const void*>(k_bilateral_outlier<6, false, true>)
                     : reinterpret_cast<const void*>(k_bilateral_outlier<6, false, false>);
  static_assert(sizeof(BilateralArgs) + sizeof(OutlierArgs) + 64 <= sizeof(k->storage), "KernelLaunch::storage too small");
  k->Reset(func, grid, block, 0, KID_BILATERAL_OUTLIER);
  k->Arg(a);
  k->Arg(o ? *o : kNoOutlier);
  k->Arg(out);
  k->Arg(out_pitch);
}

int LaunchBilateral(cudaStream_t stream, const BilateralArgs& a, const OutlierArgs* o, u16* out, size_t out_pitch) {
  if (a.radius == 6) {
    KernelLaunch k;
    DescribeBilateralOutlier(&k, a, o, out, out_pitch);
    LaunchOnStream(stream, k, false);
  } else {
    if (a.radius < 0) return SetError(SM_ERR_INVALID_ARGUMENT, "negative bilateral radius");
    { LaunchScope scope(stream, KID_BILATERAL_GENERIC); LaunchKernel(k_bilateral_generic, PixelGrid(a.width, a.height), dim3(256), 0, stream, a, out, out_pitch); }
    if (o) {
      { LaunchScope scope(stream, KID_OUTLIER); LaunchKernel(k_outlier, PixelGrid(a.width, a.height), dim3(256), 0, stream, *o, out, out_pitch, out, out_pitch); }
    }
  }
  return CheckLaunch("bilateral/outlier");
}

}  // namespace

// ---- entry points used by api.cu ----------------------------------------------------------

namespace {
// Arguments of the two fused launches for one frame (APP/main.cc:1015-1191).
int MakePreprocessArgs(const sm_preprocess_params& p, int width, int height, float fx, float fy, float cx, float cy,
                       const u16* raw, size_t raw_pitch, const u16* const* other_depths, const size_t* other_pitches,
                       const float* others_TR_reference, u16* scratch_B, size_t scratch_B_pitch, u16* out_depth,
                       size_t out_depth_pitch, float2* out_normals, size_t out_normals_pitch, float* out_radius,
                       size_t out_radius_pitch, uint4* clear_assoc, float* clear_first_depth, u8* clear_supported,
                       u16* out_depth_copy, size_t out_depth_copy_pitch, BilateralArgs* b, OutlierArgs* o, TailArgs* t) {
  if (p.depth_erosion_radius < 0 || p.depth_erosion_radius > kMaxErode) {
    return SetError(SM_ERR_INVALID_ARGUMENT, "depth_erosion_radius must be in [0, 3]");
  }
  *b = MakeBilateralArgs(p.bilateral_filter_sigma_xy, p.bilateral_filter_sigma_depth_factor, 0,
                         p.bilateral_filter_radius_factor,
                         static_cast<u16>(p.depth_scaling * p.max_depth),  // main.cc:1021
                         p.depth_valid_region_radius, width, height, raw, raw_pitch);
  const int status = MakeOutlierArgs(o, p.outlier_filtering_frame_count, p.outlier_filtering_required_inliers,
                                     p.outlier_filtering_depth_tolerance_factor, fx, fy, cx, cy, width, height,
                                     other_depths, other_pitches, others_TR_reference);
  if (status != SM_OK) return status;
  t->width = width; t->height = height;
  t->erosion_radius = p.depth_erosion_radius;
  t->normals = MakeNormalsArgs(p.observation_angle_threshold_deg, p.depth_scaling, fx, fy, cx, cy);
  t->radii = MakeRadiiArgs(p.point_radius_extension_factor, p.point_radius_clamp_factor, p.depth_scaling, fx, fy, cx, cy);
  t->in = scratch_B; t->in_pitch = scratch_B_pitch;
  t->out_depth = out_depth; t->out_depth_pitch = out_depth_pitch;
  t->out_depth_copy = out_depth_copy; t->out_depth_copy_pitch = out_depth_copy_pitch;
  t->out_normals = out_normals; t->out_normals_pitch = out_normals_pitch;
  t->out_radius = out_radius; t->out_radius_pitch = out_radius_pitch;
  t->assoc = clear_assoc; t->first_depth = clear_first_depth; t->supported = clear_supported;
  t->timeline = nullptr;
  t->skip = 0;
  return SM_OK;
}

// `in_map`: TMA descriptor of the raster t.in (MakeDepthTensorMap) or null -> 128-bit vector fill.
void DescribeTail(KernelLaunch* k, const TailArgs& t, const TensorMapStorage* in_map) {
  static_assert(sizeof(TailArgs) + sizeof(CUtensorMap) + 128 <= sizeof(k->storage), "KernelLaunch::storage too small");
  static_assert(sizeof(TensorMapStorage) == sizeof(CUtensorMap) && alignof(TensorMapStorage) == alignof(CUtensorMap), "TensorMapStorage");
  static const TensorMapStorage kNoMap = {};
  const bool tma = in_map != nullptr;
  k->Reset(tma ? reinterpret_cast<const void*>(k_erode_normals_radii<kFillTma>)
               : reinterpret_cast<const void*>(k_erode_normals_radii<kFillVector>),
           TailGrid(t.width, t.height), dim3(256), 0, KID_ERODE_NORMALS_RADII);
  k->Arg(t);
  k->Arg(tma ? *in_map : kNoMap);
}
}  // namespace

int PreprocessFused(cudaStream_t stream, const sm_preprocess_params& p, int width, int height, float fx, float fy,
                    float cx, float cy, const u16* raw, size_t raw_pitch, const u16* const* other_depths,
                    const size_t* other_pitches, const float* others_TR_reference, u16* scratch_B,
                    size_t scratch_B_pitch, u16* out_depth, size_t out_depth_pitch, float2* out_normals,
                    size_t out_normals_pitch, float* out_radius, size_t out_radius_pitch, uint4* clear_assoc,
                    float* clear_first_depth, u8* clear_supported, u16* out_depth_copy,
                    size_t out_depth_copy_pitch, unsigned long long* timeline_bilateral,
                    unsigned long long* timeline_tail, const TensorMapStorage* scratch_B_map) {
  BilateralArgs b;
  OutlierArgs o;
  TailArgs t;
  int status = MakePreprocessArgs(p, width, height, fx, fy, cx, cy, raw, raw_pitch, other_depths, other_pitches,
                                  others_TR_reference, scratch_B, scratch_B_pitch, out_depth, out_depth_pitch,
                                  out_normals, out_normals_pitch, out_radius, out_radius_pitch, clear_assoc,
                                  clear_first_depth, clear_supported, out_depth_copy, out_depth_copy_pitch, &b, &o, &t);
  if (status != SM_OK) return status;
  b.timeline = timeline_bilateral;
  t.timeline = timeline_tail;
  status = LaunchBilateral(stream, b, &o, scratch_B, scratch_B_pitch);
  if (status != SM_OK) return status;
  KernelLaunch k;
  DescribeTail(&k, t, scratch_B_map);
  LaunchOnStream(stream, k, true);
  return CheckLaunch("erode/normals/radii");
}

int DescribePreprocess(KernelLaunch* bilateral, KernelLaunch* tail, bool skip, const sm_preprocess_params& p, int width,
                       int height, float fx, float fy, float cx, float cy, const u16* raw, size_t raw_pitch,
                       const u16* const* other_depths, const size_t* other_pitches, const float* others_TR_reference,
                       u16* scratch_B, size_t scratch_B_pitch, u16* out_depth, size_t out_depth_pitch,
                       float2* out_normals, size_t out_normals_pitch, float* out_radius, size_t out_radius_pitch,
                       uint4* clear_assoc, float* clear_first_depth, u8* clear_supported, u16* out_depth_copy,
                       size_t out_depth_copy_pitch, unsigned long long* timeline_bilateral,
                       unsigned long long* timeline_tail, const TensorMapStorage* scratch_B_map) {
  BilateralArgs b;
  OutlierArgs o;
  TailArgs t;
  const int status = MakePreprocessArgs(p, width, height, fx, fy, cx, cy, raw, raw_pitch, other_depths, other_pitches,
                                        others_TR_reference, scratch_B, scratch_B_pitch, out_depth, out_depth_pitch,
                                        out_normals, out_normals_pitch, out_radius, out_radius_pitch, clear_assoc,
                                        clear_first_depth, clear_supported, out_depth_copy, out_depth_copy_pitch, &b,
                                        &o, &t);
  if (status != SM_OK) return status;
  if (b.radius != 6) return SetError(SM_ERR_INVALID_ARGUMENT, "the frame graph needs the fused bilateral kernel (radius 6)");
  b.timeline = timeline_bilateral;
  t.timeline = timeline_tail;
  b.skip = skip ? 1 : 0;
  t.skip = skip ? 1 : 0;
  DescribeBilateralOutlier(bilateral, b, &o, scratch_B, scratch_B_pitch);
  DescribeTail(tail, t, scratch_B_map);
  return SM_OK;
}

// TMA descriptor of a pitched u16 raster for the tile fill of k_erode_normals_radii: 2-D tiled, box
// kTailBW x kTailBH elements, no swizzle / interleave, out-of-bounds elements read as zero.
// cuTensorMapEncodeTiled is a driver entry point: resolved through the runtime (no libcuda link).
int MakeDepthTensorMap(TensorMapStorage* out, const u16* base, size_t pitch_bytes, int width, int height) {
  typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiled encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult query;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &query) != cudaSuccess ||
        query != cudaDriverEntryPointSuccess || fn == nullptr) {
      cudaGetLastError();
      return SetError(SM_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    }
    encode = reinterpret_cast<EncodeTiled>(fn);
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch_bytes & 15) != 0)
    return SetError(SM_ERR_INVALID_ARGUMENT, "TMA needs a 16-byte aligned raster and pitch");
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(width), static_cast<cuuint64_t>(height)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch_bytes)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kTailBW), static_cast<cuuint32_t>(kTailBH)};
  const cuuint32_t element_strides[2] = {1, 1};
  const CUresult res = encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<u16*>(base),
                              dims, strides, box, element_strides, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (res != CUDA_SUCCESS) return SetError(SM_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  return SM_OK;
}

int StageBilateral(cudaStream_t stream, float sigma_xy, float sigma_value_factor, u16 value_to_ignore,
                   float radius_factor, u16 max_depth, float depth_valid_region_radius, int width, int height,
                   const u16* in, size_t in_pitch, u16* out, size_t out_pitch) {
  const BilateralArgs b = MakeBilateralArgs(sigma_xy, sigma_value_factor, value_to_ignore, radius_factor, max_depth,
                                            depth_valid_region_radius, width, height, in, in_pitch);
  return LaunchBilateral(stream, b, nullptr, out, out_pitch);
}

int StageOutlier(cudaStream_t stream, int other_count, int required_count, float tolerance, float fx, float fy,
                 float cx, float cy, int width, int height, const u16* in, size_t in_pitch,
                 const u16* const* other_depths, const size_t* other_pitches, const float* others_TR_reference,
                 u16* out, size_t out_pitch) {
  OutlierArgs o;
  const int status = MakeOutlierArgs(&o, other_count, required_count, tolerance, fx, fy, cx, cy, width, height,
                                     other_depths, other_pitches, others_TR_reference);
  if (status != SM_OK) return status;
  { LaunchScope scope(stream, KID_OUTLIER); LaunchKernel(k_outlier, PixelGrid(width, height), dim3(256), 0, stream, o, in, in_pitch, out, out_pitch); }
  return CheckLaunch("outlier");
}

// `iterations` launches ping-ponging between `out` and `scratch` so that the last one writes `out`;
// iterations == 0 copies. `in` may alias neither.
int StageMedianDensify(cudaStream_t stream, int iterations, int width, int height, const u16* in, size_t in_pitch,
                       u16* out, size_t out_pitch, u16* scratch, size_t scratch_pitch) {
  if (iterations < 0) return SetError(SM_ERR_INVALID_ARGUMENT, "median_filter_and_densify_iterations < 0");
  if (iterations == 0) {
    if (cudaMemcpy2DAsync(out, out_pitch, in, in_pitch, width * sizeof(u16), height, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
      return SetError(SM_ERR_CUDA, "cudaMemcpy2DAsync (median, 0 iterations)");
    return SM_OK;
  }
  if (iterations > 1 && scratch == nullptr) return SetError(SM_ERR_INVALID_ARGUMENT, "median densify: scratch buffer needed for > 1 iteration");
  const u16* src = in;
  size_t src_pitch = in_pitch;
  for (int i = 0; i < iterations; ++i) {
    const bool to_out = ((iterations - 1 - i) % 2) == 0;  // the last iteration writes `out`
    u16* dst = to_out ? out : scratch;
    const size_t dst_pitch = to_out ? out_pitch : scratch_pitch;
    { LaunchScope scope(stream, KID_MEDIAN_DENSIFY); LaunchKernel(k_median_densify, PixelGrid(width, height), dim3(256), 0, stream, width, height, src, src_pitch, dst, dst_pitch); }
    src = dst;
    src_pitch = dst_pitch;
  }
  return CheckLaunch("median densify");
}

int StageErode(cudaStream_t stream, int radius, int width, int height, const u16* in, size_t in_pitch, u16* out,
               size_t out_pitch) {
  if (radius < 0 || radius > kMaxErode) return SetError(SM_ERR_INVALID_ARGUMENT, "radius value is not supported");
  { LaunchScope scope(stream, KID_ERODE); LaunchKernel(k_erode, PixelGrid(width, height), dim3(256), 0, stream, radius, width, height, in, in_pitch, out, out_pitch); }
  return CheckLaunch("erode");
}

int StageNormals(cudaStream_t stream, float observation_angle_threshold_deg, float depth_scaling, float fx, float fy,
                 float cx, float cy, int width, int height, const u16* in, size_t in_pitch, u16* out, size_t out_pitch,
                 float2* normals, size_t normals_pitch) {
  { LaunchScope scope(stream, KID_NORMALS); LaunchKernel(k_normals, PixelGrid(width, height), dim3(256), 0, stream, 
      MakeNormalsArgs(observation_angle_threshold_deg, depth_scaling, fx, fy, cx, cy), width, height, in, in_pitch, out,
      out_pitch, normals, normals_pitch); }
  return CheckLaunch("normals");
}

int StageRadii(cudaStream_t stream, float point_radius_extension_factor, float point_radius_clamp_factor,
               float depth_scaling, float fx, float fy, float cx, float cy, int width, int height, const u16* in,
               size_t in_pitch, float* radius, size_t radius_pitch, u16* out, size_t out_pitch) {
  { LaunchScope scope(stream, KID_RADII); LaunchKernel(k_radii, PixelGrid(width, height), dim3(256), 0, stream, 
      MakeRadiiArgs(point_radius_extension_factor, point_radius_clamp_factor, depth_scaling, fx, fy, cx, cy), width,
      height, in, in_pitch, radius, radius_pitch, out, out_pitch); }
  return CheckLaunch("radii");
}


// One shared-memory carve-out for every kernel of the file (see sm_create in api.cu).
int ConfigurePreprocessKernels(int carveout_percent) {
  if (carveout_percent < 0) return SM_OK;
  cudaFuncSetAttribute(k_bilateral_outlier<6, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
  cudaFuncSetAttribute(k_bilateral_outlier<6, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
  cudaFuncSetAttribute(k_bilateral_outlier<6, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
  cudaFuncSetAttribute(k_bilateral_outlier<6, false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
  cudaFuncSetAttribute(k_bilateral_generic, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
  cudaFuncSetAttribute(k_outlier, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
  cudaFuncSetAttribute(k_erode_normals_radii<kFillTma>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
  cudaFuncSetAttribute(k_erode_normals_radii<kFillVector>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
  cudaFuncSetAttribute(k_erode, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
  cudaFuncSetAttribute(k_normals, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
  cudaFuncSetAttribute(k_radii, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
  cudaFuncSetAttribute(k_median_densify, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
  cudaGetLastError();
  return SM_OK;
}

}  // namespace smb
