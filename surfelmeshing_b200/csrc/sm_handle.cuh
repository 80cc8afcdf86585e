// sm_handle.cuh — the handle behind the C ABI (struct sm_reconstruction), shared by api.cu
// (entry points) and pipeline.cu (the RGB-D stream runner and its frame graph).
#pragma once

#include <vector>

#include "sm_kernels.cuh"

namespace smb {

constexpr int kSets = 3;  // buffer sets of the frame pipeline: frames f, f + 1, f + 2 are in flight

// Supporting-surfel tie-break configuration (sm_configure "tiebreak_*"; TieBreak in sm_kernels.cuh
// holds the per-frame values derived from it).
struct TieBreakConfig {
  u32 wave;              // slots per launch wave of the modelled race (0: plain "primary, then lowest index")
  double early_fraction; // fraction of secondary associations that compete like primary ones
  double index_order_fraction;  // fraction of the pixels that order the supporters of a wave by slot index
  double early_fraction_later, index_order_fraction_later;  // the same two for the waves after the first (< 0: same as the first)
  double early_fraction_second;  // early fraction of the second wave alone (< 0: same as the later waves)
  u32 wave_offset;       // 1: per-pixel random phase of the wave boundaries (sm_kernels.cuh: tb_phase)
  u32 lane_request;      // log2 of the slots that keep their order in the shuffled order (5: a warp of the reference)
  u32 lane_shift;        // what is in effect: lane_request, or 0 when the wave is not a multiple of it
  u32 mul, mul_inv;      // derived from wave and lane_shift
};
// Defaults (DESIGN.md section 4 / profiles/r02_race_stats.md have the measurements that picked them):
// wave = the reference's AssociateSurfels launch wave on a B200 (1024-thread blocks, 31 registers ->
// 2 blocks x 148 SMs = 296 blocks of slots); inside a wave and a kind the lower slot won 100 % of the pairs
// that sit in one warp, ~49 % across the warps of one block and 52 - 70 % across blocks (the warps of a wave in a
// shuffled order that keeps the lanes of a warp in order, and a quarter of the pixels in plain slot order). The
// blocks of the FIRST wave start together; those of the later waves start one by one as earlier blocks retire, so
// there arrival follows the slot order more closely (lower slot wins 80 % instead of 70 %) and a secondary
// association beats a primary one of the same wave 5.4 % of the time instead of 0.6 %: separate fractions for the
// first wave, the second (at VGA sizes only partly filled) and the later ones (1 % / 1.5 % / 3 % early secondaries,
// 25 % / 45 % / 45 % of the pixels in slot order) put the per-frame merge flags at 1.2x the reference's own run-to-run
// difference and the free-running totals of the 500- and 1000-frame VGA streams inside the reference's spread
// (1280x960: -0.1 % slots, -0.06 % merges).
constexpr u32 kDefaultTieBreakWave = 296 * 1024;
constexpr u32 kDefaultTieBreakLaneShift = 5;
constexpr u32 kDefaultTieBreakWaveOffset = 0;
constexpr double kDefaultTieBreakEarlyFraction = 0.01;
constexpr double kDefaultTieBreakIndexOrderFraction = 0.25;
constexpr double kDefaultTieBreakEarlyFractionLater = 0.03;
constexpr double kDefaultTieBreakIndexOrderFractionLater = 0.45;
constexpr double kDefaultTieBreakEarlyFractionSecond = 0.015;
TieBreak MakeTieBreak(const TieBreakConfig& cfg, u32 frame_index);
int SetTieBreakWave(TieBreakConfig* cfg, u32 wave, u32 capacity);   // uses cfg->lane_request

struct FrameGraph;  // pipeline.cu
void DestroyFrameGraph(FrameGraph* g);

}  // namespace smb

struct sm_reconstruction {
  smb::DeviceState d{};
  int device = 0;
  int sm_count = 0;
  smb::LaunchPlan plan{};
  float fx = 0, fy = 0, cx = 0, cy = 0;
  int count_slot = 0;             // Counters::surfel_count slot holding the current count
  bool rasters_cleared = false;   // the fused pre-processing tail already reset the rasters
  smb::IntegrateEvents events{};
  smb::TieBreakConfig tiebreak{};
  smb::TieBreak last_tiebreak{};  // of the most recent frame (decodes the supporting raster, sm_download_rasters)
  // host mirror of the counters (pinned) for on-demand queries
  smb::Counters* host_counters = nullptr;
  // Stream of the most recently submitted work: the count queries, which have no stream argument,
  // synchronise with it (work on a non-blocking stream is not ordered with the NULL stream).
  cudaStream_t last_stream = nullptr;
  // pre-processing scratch (APP/main.cc filtered_depth_buffer_B)
  smb::u16* scratch_B = nullptr; size_t scratch_B_pitch = 0;
  smb::TensorMapStorage scratch_B_map{};   // TMA descriptor of scratch_B (tile fill of the pre-processing tail)
  bool scratch_B_map_valid = false;
  const smb::TensorMapStorage* ScratchBMap() const { return scratch_B_map_valid ? &scratch_B_map : nullptr; }
  // sm_integrate: snapshot of the depth before the measurement blending (k_blend reads the snapshot and
  // writes the caller's buffer: see the kernel)
  smb::u16* blend_src = nullptr; size_t blend_src_pitch = 0;
  // Association rasters / lists per buffer set: the pre-processing tail of a later frame resets one set
  // while Integrate() of an earlier frame still works on another (sm_stream_run).
  smb::PixelAssoc* assoc_set[smb::kSets] = {};
  float* first_depth_set[smb::kSets] = {};
  smb::u8* supported_set[smb::kSets] = {};
  smb::VisEntry* vis_set[smb::kSets] = {};
  smb::u32* seg_count_set[smb::kSets] = {};
  smb::u8* merge_flag_set[smb::kSets] = {};
  // stream-runner buffers (pre-processing outputs per buffer set)
  smb::u16* run_depth[smb::kSets] = {}; size_t run_depth_pitch = 0;
  float2* run_normals[smb::kSets] = {}; size_t run_normals_pitch = 0;
  float* run_radius[smb::kSets] = {}; size_t run_radius_pitch = 0;
  smb::u16* run_depth_pre[smb::kSets] = {};   // pre-blend copy of run_depth (merge runs next to blend)
  float* smooth_alt = nullptr;    // second smooth-position buffer (DeviceState::smooth / smooth_next)
  // multi-stream pipeline of round 1 (SM_B200_GRAPH=0)
  smb::PipelineCtx pipe{};
  cudaStream_t pre_stream = nullptr;
  cudaEvent_t pre_done[2] = {nullptr, nullptr}, int_done[2] = {nullptr, nullptr}, entry_event = nullptr;
  // host-resident streams: raw-depth / colour rings filled by the upload stream
  std::vector<smb::u16*> ring_depth; size_t ring_depth_pitch = 0;
  std::vector<uchar3*> ring_color; size_t ring_color_pitch = 0;
  cudaStream_t upload_stream = nullptr;
  cudaEvent_t upload_done = nullptr;
  // MedianFilterAndDensifyDepthMap passes applied to every raw depth map entering the ring
  // (APP/main.cc:435 --median_filter_and_densify_iterations, default 0) and their staging buffers
  int median_iterations = 0;
  smb::u16* median_stage[2] = {nullptr, nullptr}; size_t median_stage_pitch = 0;
  std::vector<cudaEvent_t> iteration_done;   // frame graph: one event per iteration slot (ring reuse)
  // delta transfer (transfer.cu): operation counter, per-operation regularisation thresholds, staging
  smb::u32 op_epoch = 0;
  uint64_t state_generation = 1;   // bumped by sm_reset / sm_load_state
  struct Operation { smb::u32 epoch; int stamp_threshold; };
  std::vector<Operation> op_history;
  smb::u32* delta_index = nullptr; float* delta_values = nullptr; smb::u32* delta_cursor = nullptr;
  smb::u32 delta_capacity = 0;
  smb::u32* delta_host = nullptr; size_t delta_host_capacity = 0;   // pinned, in 32-bit words
  // frame graph (pipeline.cu)
  cudaStream_t graph_stream = nullptr;
  cudaEvent_t graph_exit = nullptr;
  smb::FrameGraph* graph = nullptr;
};

namespace smb {
// api.cu
int FetchCounters(sm_reconstruction* r, cudaStream_t stream);
FrameParams MakeFrameParams(const sm_reconstruction* r, u32 frame_index, int count_slot, const sm_integrate_params& p,
                            u16* depth, size_t depth_pitch, const u16* depth_pre, size_t depth_pre_pitch,
                            const float* normals, size_t normals_pitch, const float* radius, size_t radius_pitch,
                            const uint8_t* color, size_t color_pitch, const float* global_T_local,
                            const float* local_T_global);
int IntegrateImpl(sm_reconstruction* r, cudaStream_t stream, u32 frame_index, const sm_integrate_params& p, u16* depth,
                  size_t depth_pitch, const float* normals, size_t normals_pitch, const float* radius,
                  size_t radius_pitch, const uint8_t* color, size_t color_pitch, const float* global_T_local,
                  const float* local_T_global);
void CountLaunches(unsigned long long n);
unsigned long long LaunchCount();
// transfer.cu
void RecordOperation(sm_reconstruction* r, int stamp_threshold);   // one Integrate() / Regularize(): ++op_epoch
void FreeTransferBuffers(sm_reconstruction* r);
int TransferDelta(sm_reconstruction* r, cudaStream_t stream, uint32_t frame_index, sm_transfer_token* token, float* x,
                  float* y, float* z, float* radius_squared, float* nx, float* ny, float* nz,
                  uint32_t* last_update_stamp, sm_transfer_stats* stats);
int UpdateVisualizationBuffers(sm_reconstruction* r, cudaStream_t stream, const sm_visualization_params& p, float* vertex,
                               uint32_t* neighbor_index, float* normal_vertex);
// pipeline.cu
int StreamRun(sm_reconstruction* r, cudaStream_t stream, const sm_stream_desc* s, const sm_preprocess_params* pp,
              const sm_integrate_params* ip, int first_frame, int last_frame, sm_stream_stats* stats);
}  // namespace smb
