// regularize.cu — surfel regularisation for sm_100a (SURVEY §8 a14 + the second half of a12).
//
// Replaces RegularizeSurfelsCUDA (APP/cuda_surfel_reconstruction_kernels.cu:2099-2410: Clear,
// Accumulate, Step, Update = 4 sweeps over all slots) and
// UpdateNeighborsCUDARemoveReplacedNeighborsKernel (:1420-1437, a 5th sweep) by 2 sweeps:
//
//   k_reg_accumulate : [drop neighbour links to surfels with the detach flag] + Accumulate
//   k_reg_step       : gradient step (:2197-2290) from the current smooth buffer into the other one
//                      (= the reference's Update sweep, :2292-2308), accumulators reset to zero
//
// The gradient / weight accumulators (the reference's rows 11-13 and 23) live in one float4 record
// per slot (DeviceState::gradient) so that a neighbour contribution is one vector atomic.
// Invariant that makes the Clear sweep unnecessary: the records are zero between calls (new
// surfels are created with zeros, Accumulate only adds into surfels inside the regularisation
// window, and exactly those are reset by k_reg_step). The SoA rows 11-13 / 23 stay zero.
// Float atomics make the accumulated gradients order-dependent, as in the reference.

#include <algorithm>
#include <cstdlib>

#include "sm_kernels.cuh"

namespace smb {

namespace {

#define SM_S(row, i) d.surfels[static_cast<size_t>(row) * d.stride + (i)]
#define SM_SU(row, i) reinterpret_cast<u32*>(d.surfels)[static_cast<size_t>(row) * d.stride + (i)]
#define SM_SMOOTH(axis, i) d.smooth[static_cast<size_t>(axis) * d.stride + (i)]
#define SM_SMOOTH_NEXT(axis, i) d.smooth_next[static_cast<size_t>(axis) * d.stride + (i)]

constexpr int kBlock = 256;
// Minimum resident blocks per SM of the two sweeps (0 = whatever the register count gives: 4 and 5).
// A/B hook (tools/build_variant.sh): more resident warps against spills.
#ifndef SM_REG_ACCUMULATE_MIN_BLOCKS
#define SM_REG_ACCUMULATE_MIN_BLOCKS 1
#endif
#ifndef SM_REG_STEP_MIN_BLOCKS
#define SM_REG_STEP_MIN_BLOCKS 1
#endif

struct RegParams {
  u32 frame_index;
  int window;                 // regularization_frame_window_size
  float radius_factor_squared;
  float regularizer_weight;
  int count_slot;
  int remove_below_slot;      // -1: no detach-flag pass
  int skip;                   // placeholder launch of the frame graph: return at once
};

// `stamp < frame_index - window` evaluated like the reference: the subtraction in u32, the
// comparison in int (kernels.cu:2132,2206).
__device__ __forceinline__ bool outside_window(u32 stamp, const RegParams& p) {
  return static_cast<int>(stamp) < static_cast<int>(p.frame_index - static_cast<u32>(p.window));
}

__global__ void __launch_bounds__(kBlock, SM_REG_ACCUMULATE_MIN_BLOCKS) k_reg_accumulate(DeviceState d, RegParams p) {
  pdl_prologue();
  if (p.skip) return;
  const TimelineScope timeline_scope(d, p.frame_index, KID_REG_ACCUMULATE);
  const u32 n = d.counters->surfel_count[p.count_slot];
  const u32 n_remove = p.remove_below_slot >= 0 ? d.counters->surfel_count[p.remove_below_slot] : 0u;
  // The sweep is a chain of dependent gathers per surfel; the neighbour links (its first level)
  // are requested one round ahead, the first round's before the counts have arrived.
  const u32 step = gridDim.x * blockDim.x;
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  u32 nbr_ahead[4] = {kInvalidIndex, kInvalidIndex, kInvalidIndex, kInvalidIndex};
  if (i < d.stride) {
#pragma unroll
    for (int k = 0; k < 4; ++k) nbr_ahead[k] = SM_SU(SM_ROW_NEIGHBOR0 + k, i);
  }
  for (; i < n; i += step) {
    u32 nbr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) nbr[k] = nbr_ahead[k];
    if (i + step < n) {
#pragma unroll
      for (int k = 0; k < 4; ++k) nbr_ahead[k] = SM_SU(SM_ROW_NEIGHBOR0 + k, i + step);
    }
    if ((nbr[0] & nbr[1] & nbr[2] & nbr[3]) == kInvalidIndex) continue;  // no neighbours at all

    // batch 1: detach flag + stamp of the neighbours (one gather each: row kRowMeta), and this surfel's
    // own attributes
    u32 meta[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const u32 q = nbr[k] != kInvalidIndex ? nbr[k] : i;
      meta[k] = SM_SU(kRowMeta, q);
    }
    const float sx = SM_SMOOTH(0, i), sy = SM_SMOOTH(1, i), sz = SM_SMOOTH(2, i);
    const float nx = SM_S(SM_ROW_NORMAL_X, i), ny = SM_S(SM_ROW_NORMAL_Y, i), nz = SM_S(SM_ROW_NORMAL_Z, i);
    const float radius_squared = SM_S(SM_ROW_RADIUS_SQUARED, i);

    // UpdateNeighborsCUDARemoveReplacedNeighborsKernel (kernels.cu:1420-1437) for the slots
    // that existed before this frame, then the in-window count (kernels.cu:2125-2139).
    bool use[4];
    int neighbor_count = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i < n_remove && nbr[k] != kInvalidIndex && (meta[k] & kMetaDetachBit)) {
        nbr[k] = kInvalidIndex;
        SM_SU(SM_ROW_NEIGHBOR0 + k, i) = kInvalidIndex;
      }
      use[k] = nbr[k] != kInvalidIndex && !outside_window(meta[k] & ~kMetaDetachBit, p);
      neighbor_count += use[k] ? 1 : 0;
    }
    if (neighbor_count == 0) continue;

    // batch 2: smooth positions of the neighbours
    float qx[4], qy[4], qz[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const u32 q = use[k] ? nbr[k] : i;
      qx[k] = SM_SMOOTH(0, q);
      qy[k] = SM_SMOOTH(1, q);
      qz[k] = SM_SMOOTH(2, q);
    }
    const float max_distance_squared = fmul(radius_squared, p.radius_factor_squared);
    const float rcp_count = frcp(i2f(neighbor_count));
    const float factor = fmul(fadd(p.regularizer_weight, p.regularizer_weight), rcp_count);  // 2 * w / count
    const float weight_term = fmul(rcp_count, p.regularizer_weight);                         // w / count
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!use[k]) continue;
      const u32 q = nbr[k];
      const float dx = fsub(qx[k], sx);
      const float dy = fsub(qy[k], sy);
      const float dz = fsub(qz[k], sz);
      const float f = fmul(factor, ffma(nz, dz, ffma(nx, dx, fmul(ny, dy))));
      // kernels.cu:2173-2176: four float atomicAdds; here one 16-byte vector atomic (sm_90+), the
      // same four fp32 additions in the same (arbitrary) arrival order
      atomicAdd(&d.gradient[q], make_float4(fmul(nx, f), fmul(ny, f), fmul(nz, f), weight_term));
      // If the neighbour is too far away, remove it (kernels.cu:2184-2192).
      if (squared_norm(dx, dy, dz) > max_distance_squared) SM_SU(SM_ROW_NEIGHBOR0 + k, i) = kInvalidIndex;
    }
  }
}

__global__ void __launch_bounds__(kBlock, SM_REG_STEP_MIN_BLOCKS) k_reg_step(DeviceState d, RegParams p) {
  pdl_prologue();
  if (p.skip) return;
  const TimelineScope timeline_scope(d, p.frame_index, KID_REG_STEP);
  const u32 n = d.counters->surfel_count[p.count_slot];
  // Stamp and neighbour links (first level of the gather chain) are requested one round ahead.
  const u32 step = gridDim.x * blockDim.x;
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  u32 stamp_ahead = 0, nbr_ahead[4] = {kInvalidIndex, kInvalidIndex, kInvalidIndex, kInvalidIndex};
  if (i < d.stride) {
    stamp_ahead = SM_SU(SM_ROW_LAST_UPDATE_STAMP, i);
#pragma unroll
    for (int k = 0; k < 4; ++k) nbr_ahead[k] = SM_SU(SM_ROW_NEIGHBOR0 + k, i);
  }
  for (; i < n; i += step) {
    const u32 stamp = stamp_ahead;
    u32 nbr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) nbr[k] = nbr_ahead[k];
    if (i + step < n) {
      stamp_ahead = SM_SU(SM_ROW_LAST_UPDATE_STAMP, i + step);
#pragma unroll
      for (int k = 0; k < 4; ++k) nbr_ahead[k] = SM_SU(SM_ROW_NEIGHBOR0 + k, i + step);
    }
    const float sx = SM_SMOOTH(0, i), sy = SM_SMOOTH(1, i), sz = SM_SMOOTH(2, i);
    if (outside_window(stamp, p)) {
      // not regularised (kernels.cu:2206): the smooth position carries over to the next buffer
      SM_SMOOTH_NEXT(0, i) = sx; SM_SMOOTH_NEXT(1, i) = sy; SM_SMOOTH_NEXT(2, i) = sz;
      continue;
    }
    const float nx = SM_S(SM_ROW_NORMAL_X, i), ny = SM_S(SM_ROW_NORMAL_Y, i), nz = SM_S(SM_ROW_NORMAL_Z, i);
    // Data term (factor 2) + neighbour-induced terms.
    const float4 accumulated = d.gradient[i];
    float gx = ffma(fsub(sx, SM_S(SM_ROW_X, i)), 2.0f, accumulated.x);
    float gy = ffma(fsub(sy, SM_S(SM_ROW_Y, i)), 2.0f, accumulated.y);
    float gz = ffma(fsub(sz, SM_S(SM_ROW_Z, i)), 2.0f, accumulated.z);
    int neighbor_count = 0;
    float rx = 0.f, ry = 0.f, rz = 0.f;
    float qx[4], qy[4], qz[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const u32 q = nbr[k] != kInvalidIndex ? nbr[k] : i;
      qx[k] = SM_SMOOTH(0, q);
      qy[k] = SM_SMOOTH(1, q);
      qz[k] = SM_SMOOTH(2, q);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (nbr[k] == kInvalidIndex) continue;
      ++neighbor_count;
      const float dx = fsub(qx[k], sx);
      const float dy = fsub(qy[k], sy);
      const float dz = fsub(qz[k], sz);
      const float normal_dot_difference = ffma(nz, dz, ffma(nx, dx, fmul(ny, dy)));
      rx = ffma(-nx, normal_dot_difference, rx);
      ry = ffma(-ny, normal_dot_difference, ry);
      rz = ffma(-nz, normal_dot_difference, rz);
    }
    if (neighbor_count > 0) {
      const float factor = fmul(fadd(p.regularizer_weight, p.regularizer_weight), frcp(i2f(neighbor_count)));
      gx = ffma(factor, rx, gx);
      gy = ffma(factor, ry, gy);
      gz = ffma(factor, rz, gz);
    }
    const float gradient_length = fsqrt_approx(ffma(gz, gz, ffma(gx, gx, fmul(gy, gy))));
    const float residual_terms_weight_sum = fadd(fadd(p.regularizer_weight, 1.0f), accumulated.w);
    float step_factor = fmul(frcp(residual_terms_weight_sum), 0.5f);
    const float max_step_length = fsqrt_approx(SM_S(SM_ROW_RADIUS_SQUARED, i));
    const float step_length = fmul(step_factor, gradient_length);
    if (step_length > max_step_length) step_factor = fmul(step_factor, fmul(max_step_length, frcp(step_length)));
    // The new smooth position goes to the other buffer (the neighbours still read the old one);
    // this surfel's accumulator, read by nobody else in this sweep, is reset for the next call.
    SM_SMOOTH_NEXT(0, i) = ffma(step_factor, -gx, sx);
    SM_SMOOTH_NEXT(1, i) = ffma(step_factor, -gy, sy);
    SM_SMOOTH_NEXT(2, i) = ffma(step_factor, -gz, sz);
    d.gradient[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// RegularizeSurfelsCUDACopyOnlyKernel (kernels.cu:2310-2327) [+ detach-flag pass].
__global__ void __launch_bounds__(kBlock) k_reg_copy_only(DeviceState d, RegParams p) {
  pdl_prologue();
  if (p.skip) return;
  const TimelineScope timeline_scope(d, p.frame_index, KID_REG_COPY_ONLY);
  const u32 n = d.counters->surfel_count[p.count_slot];
  const u32 n_remove = p.remove_below_slot >= 0 ? d.counters->surfel_count[p.remove_below_slot] : 0u;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i < n_remove) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const u32 q = SM_SU(SM_ROW_NEIGHBOR0 + k, i);
        if (q != kInvalidIndex && (SM_SU(SM_ROW_COLOR, q) >> 24) == 1u) SM_SU(SM_ROW_NEIGHBOR0 + k, i) = kInvalidIndex;
      }
    }
    if (outside_window(SM_SU(SM_ROW_LAST_UPDATE_STAMP, i), p)) continue;
    SM_SMOOTH(0, i) = SM_S(SM_ROW_X, i);
    SM_SMOOTH(1, i) = SM_S(SM_ROW_Y, i);
    SM_SMOOTH(2, i) = SM_S(SM_ROW_Z, i);
  }
}

}  // namespace

namespace {
__global__ void __launch_bounds__(kBlock) k_rebuild_meta(DeviceState d, u32 count) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const u32 flag = (SM_SU(SM_ROW_COLOR, i) >> 24) == 1u ? kMetaDetachBit : 0u;
    SM_SU(kRowMeta, i) = (SM_SU(SM_ROW_LAST_UPDATE_STAMP, i) & ~kMetaDetachBit) | flag;
  }
}
}  // namespace

int RebuildMetaRow(cudaStream_t stream, const DeviceState& d, u32 count, int sm_count) {
  if (count == 0) return SM_OK;
  k_rebuild_meta<<<sm_count * 4, kBlock, 0, stream>>>(d, count);
  return CheckLaunch("rebuild meta row");
}

int DescribeRegularize(KernelLaunch* first, KernelLaunch* second, bool skip, const LaunchPlan& plan,
                       const DeviceState& d, bool disable_denoising, u32 frame_index,
                       float radius_factor_for_regularization_neighbors, float regularizer_weight,
                       int regularization_frame_window_size, int count_slot, int remove_replaced_below_slot) {
  RegParams p;
  p.frame_index = frame_index;
  p.window = regularization_frame_window_size;
  p.radius_factor_squared =
      radius_factor_for_regularization_neighbors * radius_factor_for_regularization_neighbors;  // kernels.cu:2379
  p.regularizer_weight = regularizer_weight;
  p.count_slot = count_slot;
  p.remove_below_slot = remove_replaced_below_slot;
  p.skip = skip ? 1 : 0;
  static_assert(sizeof(DeviceState) + sizeof(RegParams) + 32 <= sizeof(first->storage), "KernelLaunch::storage too small");
  if (disable_denoising) {
    first->Reset(reinterpret_cast<const void*>(k_reg_copy_only), dim3(plan.reg_copy), dim3(kBlock), 0, KID_REG_COPY_ONLY);
    first->Arg(d);
    first->Arg(p);
    return 1;
  }
  first->Reset(reinterpret_cast<const void*>(k_reg_accumulate), dim3(plan.reg_accumulate), dim3(kBlock), 0, KID_REG_ACCUMULATE);
  first->Arg(d);
  first->Arg(p);
  second->Reset(reinterpret_cast<const void*>(k_reg_step), dim3(plan.reg_step), dim3(kBlock), 0, KID_REG_STEP);
  second->Arg(d);
  second->Arg(p);
  return 2;
}

int RegularizeSurfels(cudaStream_t stream, DeviceState& d, bool disable_denoising, u32 frame_index,
                      float radius_factor_for_regularization_neighbors, float regularizer_weight,
                      int regularization_frame_window_size, int count_slot, int remove_replaced_below_slot,
                      const LaunchPlan& plan) {
  KernelLaunch first, second;
  const int n = DescribeRegularize(&first, &second, false, plan, d, disable_denoising, frame_index,
                                   radius_factor_for_regularization_neighbors, regularizer_weight,
                                   regularization_frame_window_size, count_slot, remove_replaced_below_slot);
  LaunchOnStream(stream, first, false);
  if (n == 1) return CheckLaunch("regularize (copy only)");
  LaunchOnStream(stream, second, true);
  // k_reg_step wrote every slot of the other smooth buffer: it is the current one from here on
  float* const filled = d.smooth_next;
  d.smooth_next = d.smooth;
  d.smooth = filled;
  return CheckLaunch("regularize");
}

// Per-device configuration: one shared-memory carve-out for every kernel of the file (see
// sm_create in api.cu) and grids = the blocks resident at once (see ConfigureIntegrateKernels).
int ConfigureRegularizeKernels(int carveout_percent, LaunchPlan* plan) {
  if (carveout_percent >= 0) {
    cudaFuncSetAttribute(k_reg_accumulate, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
    cudaFuncSetAttribute(k_reg_step, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
    cudaFuncSetAttribute(k_reg_copy_only, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
    cudaGetLastError();
  }
  const char* e = std::getenv("SM_B200_RESIDENT_GRIDS");
  auto resident = [&](auto kernel) {
    int per_sm = 0;
    if ((e && e[0] == '0') || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, 0) != cudaSuccess || per_sm < 1) {
      cudaGetLastError();
      per_sm = 8;
    }
    return ScaleGrid(plan->sm_count * per_sm);
  };
  plan->reg_accumulate = resident(k_reg_accumulate);
  plan->reg_step = resident(k_reg_step);
  plan->reg_copy = resident(k_reg_copy_only);
  // SM_B200_OFFCHAIN_GRID_PERCENT (A/B hook): fraction of the resident grid for the sweeps that are not on
  // the dependency chain that ends a frame step, so that the chain's kernels find free SM resources.
  if (const char* pe = std::getenv("SM_B200_OFFCHAIN_GRID_PERCENT")) {
    const int percent = std::atoi(pe);
    if (percent > 0 && percent < 100) {
      plan->reg_accumulate = std::max(plan->sm_count, plan->reg_accumulate * percent / 100);
      plan->reg_step = std::max(plan->sm_count, plan->reg_step * percent / 100);
    }
  }
  return SM_OK;
}

}  // namespace smb
