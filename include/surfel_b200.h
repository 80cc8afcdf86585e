/*
 * surfel_b200.h — C ABI of libsurfel_b200.so, the Blackwell-native (sm_100a)
 * replacement for the per-frame surfel reconstruction hot path of
 * puzzlepaint/surfelmeshing.
 *
 * Every entry point cites the reference interface it replaces (paths relative
 * to the reference tree; APP = applications/surfel_meshing/src/surfel_meshing).
 *
 * Conventions
 *  - Plain pointers and sizes only; `stream` is a cudaStream_t passed as void*.
 *  - Rasters are row-pitched device buffers (pitch in BYTES, as handed out by
 *    cudaMallocPitch / libvis CUDABuffer<T>, libvis/src/libvis/cuda/cuda_buffer_inl.h:36-48).
 *  - Rigid transforms are 3x4 row-major float[12] (= libvis CUDAMatrix3x4 rows,
 *    libvis/src/libvis/cuda/cuda_matrix.cuh:67-116).
 *  - Camera intrinsics are the reference's PinholeCamera4f::parameters()
 *    {fx, fy, cx, cy} in pixel-CORNER convention (cx_file + 0.5), APP/
 *    cuda_surfel_reconstruction_kernels.cc:63-74.
 *  - Every function returns SM_OK (0) or a negative SM_ERR_* code; the message is
 *    available from sm_last_error(). (Reference: no return values, CUDA errors
 *    abort through LOG(FATAL), libvis/src/libvis/cuda/cuda_util.h:35-49. The
 *    C++ adapter in surfel_b200_adapter.h maps non-zero to LOG(FATAL).)
 *  - There is NO CPU fallback: if the CUDA runtime or an sm_100 device is not
 *    available the calls fail with SM_ERR_CUDA.
 */
#ifndef SURFEL_B200_H_
#define SURFEL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SM_OK 0
#define SM_ERR_CUDA (-1)
#define SM_ERR_INVALID_ARGUMENT (-2)
#define SM_ERR_CAPACITY (-3) /* surfel cap exceeded (reference: unchecked overflow) */

/* Row indices of the surfel SoA (row = attribute, column = surfel), identical
 * to APP/cuda_surfel_reconstruction_kernels.cuh:48-78. */
enum {
  SM_ROW_X = 0, SM_ROW_Y = 1, SM_ROW_Z = 2,
  SM_ROW_SMOOTH_X = 3, SM_ROW_SMOOTH_Y = 4, SM_ROW_SMOOTH_Z = 5,
  SM_ROW_CONFIDENCE = 6, SM_ROW_RADIUS_SQUARED = 7,
  SM_ROW_NORMAL_X = 8, SM_ROW_NORMAL_Y = 9, SM_ROW_NORMAL_Z = 10,
  SM_ROW_GRADIENT_X = 11, SM_ROW_GRADIENT_Y = 12, SM_ROW_GRADIENT_Z = 13,
  SM_ROW_ACCUM_X = 14, SM_ROW_ACCUM_Y = 15, SM_ROW_ACCUM_Z = 16,
  SM_ROW_CREATION_STAMP = 17, SM_ROW_LAST_UPDATE_STAMP = 18,
  SM_ROW_NEIGHBOR0 = 19, /* ..22 */
  SM_ROW_GRADIENT_COUNT = 23,
  SM_ROW_COLOR = 24,
  SM_ROW_COUNT = 25
};
#define SM_INVALID_SURFEL_INDEX 0xFFFFFFFFu /* APP/surfel.h:63, kernels.cu:74 */

/* Opaque handle: owns the surfel SoA, the scratch rasters, the compact index
 * lists and the timing events. Replaces class vis::CUDASurfelReconstruction,
 * APP/cuda_surfel_reconstruction.h:44-176. One handle per device/stream;
 * re-entrant per handle, no globals (the reference keeps a function-static
 * device buffer, APP/cuda_surfel_reconstruction_kernels.cc:479). */
typedef struct sm_reconstruction sm_reconstruction;

/* Arguments of CUDASurfelReconstruction::Integrate() that are not buffers,
 * APP/cuda_surfel_reconstruction.h:59-77; defaults APP/main.cc:279-371. */
typedef struct sm_integrate_params {
  float depth_scaling;                                       /* 5000 */
  float sensor_noise_factor;                                 /* 0.05 */
  float max_surfel_confidence;                               /* 5 */
  float regularizer_weight;                                  /* 10 */
  int32_t regularization_frame_window_size;                  /* 30 */
  int32_t do_blending;                                       /* 1 */
  int32_t measurement_blending_radius;                       /* 12 */
  int32_t regularization_iterations_per_integration_iteration; /* 1 */
  float radius_factor_for_regularization_neighbors;          /* 2 */
  float normal_compatibility_threshold_deg;                  /* 40 */
  int32_t surfel_integration_active_window_size;             /* INT_MAX */
} sm_integrate_params;

/* Arguments of the depth pre-processing call sequence APP/main.cc:1015-1191;
 * defaults APP/main.cc:415-478. */
typedef struct sm_preprocess_params {
  float depth_scaling;                          /* 5000 */
  float max_depth;                              /* 3 (metres) */
  float depth_valid_region_radius;              /* 333 (pixels) */
  float bilateral_filter_sigma_xy;              /* 3 */
  float bilateral_filter_radius_factor;         /* 2 */
  float bilateral_filter_sigma_depth_factor;    /* 0.05 */
  int32_t outlier_filtering_frame_count;        /* 8 (2,4,6,8) */
  int32_t outlier_filtering_required_inliers;   /* -1 = all */
  float outlier_filtering_depth_tolerance_factor; /* 0.02 */
  int32_t depth_erosion_radius;                 /* 2 (0..3) */
  float observation_angle_threshold_deg;        /* 85 */
  float point_radius_extension_factor;          /* 1.5 */
  float point_radius_clamp_factor;              /* +inf */
} sm_preprocess_params;

void sm_default_integrate_params(sm_integrate_params* p);
void sm_default_preprocess_params(sm_preprocess_params* p);

const char* sm_last_error(void);
/* Library/arch identification string, e.g. "surfel_b200 sm_100a". */
const char* sm_version(void);

/* ---- lifecycle ----------------------------------------------------------
 * sm_create replaces the CUDASurfelReconstruction constructor,
 * APP/cuda_surfel_reconstruction.cc:44-91 (max_surfel_count, camera; the three
 * GL resources and the render window are GUI-only and have no equivalent).
 * Uses the current CUDA device. */
int sm_create(sm_reconstruction** out, uint64_t max_surfel_count,
              int32_t width, int32_t height,
              float fx, float fy, float cx, float cy);
int sm_destroy(sm_reconstruction* r);
/* Empties the surfel cloud (surfel_count_ = merge_count_ = 0), stream-ordered. */
int sm_reset(sm_reconstruction* r, void* stream);

/* ---- depth pre-processing ------------------------
 * One call = the five launches of APP/main.cc:1015-1191:
 *   BilateralFilteringAndDepthCutoffCUDA  (APP/cuda_depth_processing.cu:120-158)
 *   OutlierDepthMapFusionCUDA<K+1,u16>    (:229-285 / :399-457)
 *   ErodeDepthMapCUDA | CopyWithoutBorder (:540-579 / :609-633)
 *   ComputeNormalsAndDropBadPixelsCUDA    (:720-762)
 *   ComputePointRadiiAndRemoveIsolatedPixelsCUDA (:839-883)
 * raw_depth: this frame's uploaded u16 depth. other_depths[k] /
 * other_pitches[k] / others_TR_reference[12*k..]: the K = outlier_filtering_
 * frame_count other RAW depth maps and (ref_T_global_scaled *
 * global_T_other_scaled)^-1 exactly as built at APP/main.cc:1039-1058 (host
 * arrays; device pointers inside). Outputs: out_depth (the reference's
 * filtered_depth_buffer_A handed to Integrate), out_normals (float2 per
 * pixel), out_radius (SQUARED radius per pixel; like the reference it is only
 * written where the normals stage kept a depth). */
int sm_preprocess(sm_reconstruction* r, void* stream,
                  const sm_preprocess_params* p,
                  const uint16_t* raw_depth, size_t raw_pitch,
                  const uint16_t* const* other_depths, const size_t* other_pitches,
                  const float* others_TR_reference,
                  uint16_t* out_depth, size_t out_depth_pitch,
                  float* out_normals, size_t out_normals_pitch,
                  float* out_radius, size_t out_radius_pitch);

/* The five stages individually (same kernels the fused call is built from);
 * these are what the vis::-named link shims forward to. */
int sm_bilateral_filter_and_depth_cutoff(
    void* stream, float sigma_xy, float sigma_value_factor, uint16_t value_to_ignore,
    float radius_factor, uint16_t max_depth, float depth_valid_region_radius,
    int32_t width, int32_t height,
    const uint16_t* in_depth, size_t in_pitch, uint16_t* out_depth, size_t out_pitch);
int sm_outlier_depth_map_fusion(
    void* stream, int32_t other_count, int32_t required_count /* -1 = all */,
    float tolerance, float fx, float fy, float cx, float cy,
    int32_t width, int32_t height,
    const uint16_t* in_depth, size_t in_pitch,
    const uint16_t* const* other_depths, const size_t* other_pitches,
    const float* others_TR_reference,
    uint16_t* out_depth, size_t out_pitch);
/* MedianFilterAndDensifyDepthMap(), APP/main.cc:207-252, which the reference runs on the CPU
 * inside its upload loop (main.cc:927-939): `iterations` passes (main.cc:435,
 * --median_filter_and_densify_iterations) of the 3x3 zero-excluding median that also fills
 * holes with >= 2 valid neighbours. Device buffers; `scratch` (same size) is needed for more
 * than one pass; the result is in out_depth. iterations == 0 copies. */
int sm_median_filter_and_densify_depth_map(
    void* stream, int32_t iterations, int32_t width, int32_t height,
    const uint16_t* in_depth, size_t in_pitch, uint16_t* out_depth, size_t out_pitch,
    uint16_t* scratch, size_t scratch_pitch);
int sm_erode_depth_map(void* stream, int32_t radius /* 0 = copy w/o border */,
                       int32_t width, int32_t height,
                       const uint16_t* in_depth, size_t in_pitch,
                       uint16_t* out_depth, size_t out_pitch);
int sm_compute_normals_and_drop_bad_pixels(
    void* stream, float observation_angle_threshold_deg, float depth_scaling,
    float fx, float fy, float cx, float cy, int32_t width, int32_t height,
    const uint16_t* in_depth, size_t in_pitch, uint16_t* out_depth, size_t out_pitch,
    float* out_normals, size_t normals_pitch);
int sm_compute_point_radii_and_remove_isolated_pixels(
    void* stream, float point_radius_extension_factor, float point_radius_clamp_factor,
    float depth_scaling, float fx, float fy, float cx, float cy,
    int32_t width, int32_t height,
    const uint16_t* in_depth, size_t in_pitch, float* out_radius, size_t radius_pitch,
    uint16_t* out_depth, size_t out_pitch);

/* ---- Integrate / Regularize ---------------------------
 * sm_integrate replaces CUDASurfelReconstruction::Integrate(),
 * APP/cuda_surfel_reconstruction.cc:112-320. `depth` is in/out (blended in
 * place, :201-213); `color` is packed uchar3 rows. global_T_local and
 * local_T_global = global_T_local^-1 are both supplied so that both sides of
 * a parity test consume bit-identical matrices (the reference computes the
 * inverse with Sophus on the host, :144,:156,:181,:251).
 * Unlike the reference the call does NOT block the host (the reference
 * synchronises twice, cuda_surfel_reconstruction_kernels.cc:509 and
 * cuda_surfel_reconstruction.cc:290): counts stay device-resident and are
 * fetched by sm_surfel_count()/sm_surfels_size() on demand. */
int sm_integrate(sm_reconstruction* r, void* stream, uint32_t frame_index,
                 const sm_integrate_params* p,
                 uint16_t* depth, size_t depth_pitch,
                 const float* normals, size_t normals_pitch,
                 const float* radius, size_t radius_pitch,
                 const uint8_t* color, size_t color_pitch,
                 const float global_T_local[12], const float local_T_global[12]);

/* Replaces CUDASurfelReconstruction::Regularize(), cuda_surfel_reconstruction.cc:322-337. */
int sm_regularize(sm_reconstruction* r, void* stream, uint32_t frame_index,
                  float regularizer_weight,
                  float radius_factor_for_regularization_neighbors,
                  int32_t regularization_frame_window_size);

/* surfel_count() = entries - merged, surfels_size() = entries in use
 * (cuda_surfel_reconstruction.h:125-128). Both synchronise with the stream of
 * the most recently submitted call on this handle and return the counters of
 * that call. */
int sm_surfel_count(sm_reconstruction* r, uint32_t* out);
int sm_surfels_size(sm_reconstruction* r, uint32_t* out);

/* Replaces TransferAllToCPU(), cuda_surfel_reconstruction.cc:339-359: fills the
 * eight CUDASurfelBuffersCPU arrays (APP/cuda_surfels_cpu.h:40-73: SMOOTH x,y,z,
 * radius^2, normal x,y,z, last-update stamp), surfels_size() entries each. The
 * host arrays may be pageable (as in the reference) or pinned. *out_count
 * receives surfels_size(). Stream-ordered; the caller synchronises the stream
 * before reading, as APP/main.cc:1261-1287 does. */
int sm_transfer_all_to_cpu(sm_reconstruction* r, void* stream, uint32_t frame_index,
                           float* x, float* y, float* z, float* radius_squared,
                           float* nx, float* ny, float* nz, uint32_t* last_update_stamp,
                           uint64_t* out_count);

/* Delta form of TransferAllToCPU. The reference copies all eight rows on
 * every transfer (cuda_surfel_reconstruction.cc:339-359; 160 MB at 5 M surfels, into pageable
 * memory) and its consumer then compares every CPU surfel with the arrays
 * (SurfelMeshing::IntegrateCUDABuffers, APP/surfel_meshing.cc:189-288). This call brings arrays
 * that hold an EARLIER transfer up to date: the slots whose transferred attributes can have
 * changed since then (new, integrated, regularised, merged) are compacted on the GPU, moved with
 * one copy through pinned staging and scattered into the untouched CUDASurfelBuffersCPU layout;
 * afterwards the arrays are identical to what sm_transfer_all_to_cpu would have produced.
 * `token` identifies the transfer that last filled THESE arrays (zero-initialise it for arrays
 * that were never filled: the call then does a full transfer) and is updated; with the
 * reference's write/read double buffer (cuda_surfels_cpu.h:83-124) keep one token per buffer.
 * After sm_reset / sm_load_state, or when most of the cloud changed, the call falls back to the
 * full transfer. Synchronises `stream`. */
typedef struct sm_transfer_token {
  uint64_t generation;    /* 0 = never */
  uint64_t epoch;
  uint64_t surfel_count;
} sm_transfer_token;
typedef struct sm_transfer_stats {
  uint64_t surfel_count;   /* surfels_size() */
  uint64_t changed_count;  /* records moved (= surfel_count for a full transfer) */
  uint64_t d2h_bytes;
  int32_t full_transfer;
  int32_t reserved;
} sm_transfer_stats;
int sm_transfer_delta_to_cpu(sm_reconstruction* r, void* stream, uint32_t frame_index,
                             sm_transfer_token* token,
                             float* x, float* y, float* z, float* radius_squared,
                             float* nx, float* ny, float* nz, uint32_t* last_update_stamp,
                             sm_transfer_stats* stats /* may be NULL */);

/* Replaces ExportVertices(), cuda_surfel_reconstruction.cc:405-410 /
 * kernels.cu:2412-2464: packed xyz (NaN for merged) and rgb, device buffers of
 * 3*surfels_size() elements each. */
int sm_export_vertices(sm_reconstruction* r, void* stream,
                       float* position_buffer, uint8_t* color_buffer);

/* Replaces UpdateVisualizationBuffers(), cuda_surfel_reconstruction.cc:361-403, and the three
 * kernels behind it (UpdateSurfelVertexBufferCUDA, UpdateNeighborIndexBufferCUDA,
 * UpdateNormalVertexBufferCUDA, kernels.cu:274-560) with one sweep. The reference writes CUDA-mapped
 * OpenGL buffers (cudaGraphicsResource_t); here the caller passes the mapped device pointers
 * (cudaGraphicsResourceGetMappedPointer) or any plain device buffers; a NULL pointer skips that
 * buffer, as the reference skips a null resource:
 *   vertex_buffer          surfels_size() x point_size_in_floats floats: x (NaN hides a surfel that
 *                          was replaced after the last triangulation), y, z, rgba bits
 *   neighbor_index_buffer  surfels_size() x 4 x {surfel, neighbour-or-surfel} u32 (16-byte aligned)
 *   normal_vertex_buffer   surfels_size() x {p, p + radius * normal} floats */
typedef struct sm_visualization_params {
  uint32_t frame_index;
  uint32_t latest_triangulated_frame_index;
  uint32_t latest_mesh_surfel_count;
  int32_t surfel_integration_active_window_size;
  uint32_t point_size_in_floats;               /* sizeof(Point3fC3u8) / sizeof(float) = 4 */
  int32_t visualize_last_update_timestamp;
  int32_t visualize_creation_timestamp;
  int32_t visualize_radii;
  int32_t visualize_normals;
} sm_visualization_params;
int sm_update_visualization_buffers(sm_reconstruction* r, void* stream, const sm_visualization_params* p,
                                    float* vertex_buffer, uint32_t* neighbor_index_buffer,
                                    float* normal_vertex_buffer);

/* ---- radius-limited k-nearest-neighbour queries for the meshing thread ----
 * Replaces CompressedOctree::FindNearestSurfelsWithinRadius<include_completed_surfels, include_free_surfels>
 * (octree.h:471, octree.cc:313-470; callers surfel_meshing.cc:421 <false, true> and :821 <true, false>) for a BATCH of
 * queries against one snapshot of the cloud: per query the <= max_result_count (<= 64) nearest points with
 * squared distance <= radius_squared, ascending, the meshing-state filter applied before the cap, exactly what
 * the octree returns (equal distances, which the octree leaves to its traversal order, come out by ascending
 * index). All pointers are DEVICE pointers; the calls are asynchronous on `stream`. An index lives on the CUDA device that
 * was current at sm_knn_create; that device must be current for every later call on it (as for a reconstruction handle).
 *
 *   sm_knn_create   index for up to max_points points (<= 2^26)
 *   sm_knn_build    bins points [0, point_count) into a hashed uniform grid of `cell_size` (choose it near the
 *                   largest query radius: a query visits the (2 r / cell_size + 1)^3 cells its ball touches).
 *                   A point is left out if radius_squared (optional) is <= 0 at its index (merged surfels,
 *                   kernels.cu:1987) or state (optional) is 255 there.
 *   sm_knn_build_from_reconstruction   the same over the handle's current surfels: smooth positions (what
 *                   TransferAllToCPU hands to the meshing thread, cuda_surfel_reconstruction.cc:345-347) of
 *                   all slots with radius_squared > 0.
 *   sm_knn_query    state (optional, one byte per point index: 0 free, 1 front, 2 completed, 255 absent; Surfel::
 *                   MeshingState, surfel.h:67-71) is read at query time, so one index serves both callers.
 *                   Outputs: [query_count][max_result_count] squared distances (+inf past the count) and indices
 *                   (0xFFFFFFFF past the count), [query_count] counts. */
typedef struct sm_knn_index sm_knn_index;
int sm_knn_create(sm_knn_index** out, uint32_t max_points);
void sm_knn_destroy(sm_knn_index* k);
int sm_knn_build(sm_knn_index* k, void* stream, uint32_t point_count, const float* x, const float* y, const float* z,
                 const float* radius_squared /* may be NULL */, const uint8_t* state /* may be NULL */, float cell_size);
int sm_knn_build_from_reconstruction(sm_knn_index* k, sm_reconstruction* r, void* stream, float cell_size,
                                     uint32_t* out_point_count /* may be NULL */);
int sm_knn_query(sm_knn_index* k, void* stream, uint32_t query_count, const float* qx, const float* qy, const float* qz,
                 const float* radius_squared, const uint8_t* state /* may be NULL */, int32_t include_completed_surfels,
                 int32_t include_free_surfels, int32_t max_result_count, float* out_distance_squared,
                 uint32_t* out_index, int32_t* out_count);
/* One neighbour batch for a meshing iteration with HOST arrays on both sides (the CUDASurfelBuffersCPU arrays of the
 * last TransferAllToCPU, cuda_surfels_cpu.h:40-73): uploads x, y, z, radius_squared [point_count], indexes the points with
 * radius_squared > 0, asks for every point its <= max_result_count nearest members within radius_factor_squared *
 * radius_squared[i] (all meshing states; max_neighbor_search_range_increase_factor^2 covers every radius
 * TriangulateSurfel can ask for, surfel_meshing.cc:323-413) and writes the [point_count][max_result_count] rows and
 * [point_count] counts to host memory (count 0 for points that are not indexed). cell_size <= 0: twice the largest
 * query radius. Synchronous; the device staging lives in the index and is reused. */
int sm_knn_batch_host(sm_knn_index* k, void* stream, uint32_t point_count, const float* x, const float* y, const float* z,
                      const float* radius_squared, float radius_factor_squared, float cell_size, int32_t max_result_count,
                      float* out_distance_squared, uint32_t* out_index, int32_t* out_count);

/* Replaces GetTimings(), cuda_surfel_reconstruction.cc:412-429 (milliseconds of
 * the last Integrate: data association, merging, blending, integration,
 * neighbour update, new-surfel creation, regularisation). */
int sm_get_timings(sm_reconstruction* r, float out_ms[7]);
/* Event timing costs a few microseconds per frame; off by default. */
int sm_enable_timings(sm_reconstruction* r, int32_t enable);

/* ---- state access for parity tests / checkpointing ----------------------
 * The reference has no save/load; these move the 25-row SoA and
 * the two counters. rows: SM_ROW_COUNT x surfels_size floats, row-major. */
int sm_dump_state(sm_reconstruction* r, void* stream, float* host_rows,
                  uint64_t host_row_stride_elems, uint32_t* surfels_size, uint32_t* merge_count);
int sm_load_state(sm_reconstruction* r, void* stream, const float* host_rows,
                  uint64_t host_row_stride_elems, uint32_t surfels_size, uint32_t merge_count);

/* Scratch rasters of the last Integrate, de-interleaved into the reference's
 * per-pixel buffers (cuda_surfel_reconstruction.h:133-144), tightly packed
 * W*H host arrays; any pointer may be NULL. */
int sm_download_rasters(sm_reconstruction* r, void* stream,
                        uint32_t* supporting_surfels, uint32_t* supporting_surfel_counts,
                        float* supporting_surfel_depth_sums, uint32_t* conflicting_surfels,
                        float* first_surfel_depth, uint8_t* new_surfel_flag_vector,
                        uint32_t* new_surfel_indices);

/* ---- RGB-D stream runner (the frame loop of APP/main.cc:885-1223) ---------
 * Runs preprocess + Integrate over frames [first_frame, last_frame) of a
 * stream whose other-frame transforms were precomputed by the caller. With
 * frames_on_host != 0 the depth/colour pointers are (pinned) host memory and
 * each raw depth map / colour image is uploaded once on an internal copy
 * stream, overlapped with compute, exactly like the reference's upload_stream
 * (APP/main.cc:902-995); otherwise they are device-resident.
 * The call never synchronises with the host inside the loop and overlaps the
 * kernels of three consecutive frames (one instantiated CUDA graph per frame
 * step on an internal stream; SM_B200_GRAPH=0 selects the
 * multi-stream event pipeline instead); `stream` only brackets the call: work enqueued on it
 * before the call is complete before the first frame starts, work enqueued
 * after the call sees all frames integrated. The call itself returns after
 * one synchronisation at the end (to fetch the counters for `stats`). With
 * sm_enable_timings(r, 1) or sm_profile_kernels(1) the frames run one kernel
 * after the other on `stream`. */
typedef struct sm_stream_desc {
  int32_t width, height, frame_count;
  int32_t frames_on_host;
  const uint16_t* depth;         /* frame_count x H x W, tightly packed */
  const uint8_t* color;          /* frame_count x H x W x 3 */
  const float* global_T_frame;   /* frame_count x 12, host */
  const float* frame_T_global;   /* frame_count x 12, host */
  const float* others_TR_reference; /* frame_count x K x 12, host (K = outlier_filtering_frame_count) */
} sm_stream_desc;

typedef struct sm_stream_stats {
  uint32_t frames_integrated;
  uint32_t surfels_size;      /* entries in use after the last frame */
  uint32_t surfel_count;      /* entries - merged */
  uint64_t kernel_launches;   /* kernels this library launched during the call */
  uint64_t h2d_bytes;         /* bytes uploaded inside the call */
  uint64_t d2h_bytes;         /* bytes downloaded inside the call */
  double host_enqueue_ms;     /* host time spent enqueuing the frames (before the final synchronisation) */
} sm_stream_stats;

int sm_stream_run(sm_reconstruction* r, void* stream, const sm_stream_desc* s,
                  const sm_preprocess_params* pp, const sm_integrate_params* ip,
                  int32_t first_frame, int32_t last_frame, sm_stream_stats* stats);

/* Named tuning knobs of a handle (no counterpart in the reference). Keys:
 *   "tiebreak_wave" (slots per launch wave of the reference's association kernel; 0 = the plain rule "primary-pixel
 *   association before secondary, then lowest index"), "tiebreak_lanes" (consecutive slots that keep their order: a
 *   warp), "tiebreak_early_fraction" / "tiebreak_index_order_fraction" (first wave), "tiebreak_early_fraction_second"
 *   (second wave), "tiebreak_early_fraction_later" / "tiebreak_index_order_fraction_later" (later waves; negative =
 *   inherit), "tiebreak_wave_offset" (measurement hook): the reproducible rule that picks the supporting surfel of a
 *   pixel with several supporters, where the reference lets the first atomicCAS win
 *   (APP/cuda_surfel_reconstruction_kernels.cu:1688); DESIGN.md section 4 has the measurements behind the defaults.
 *   "median_filter_and_densify_iterations": sm_stream_run applies that many
 *   MedianFilterAndDensifyDepthMap passes (APP/main.cc:207-252, 927-939) to every raw depth map
 *   as it enters the device-side frame ring (default 0, as in the reference). */
int sm_configure(sm_reconstruction* r, const char* key, double value);

/* Number of kernel launches issued by this library since load (all handles). */
uint64_t sm_kernel_launch_count(void);

/* Per-kernel device timing for the roofline report of bench.py: while enabled, every kernel
 * launch of the library is bracketed by CUDA events on its launching stream.
 * sm_profile_report synchronises the device and returns, per kernel id, the accumulated
 * elapsed milliseconds and the launch count since the last report. Never enabled inside a
 * timed throughput region (the extra events serialise the launches). */
int sm_profile_kernels(int32_t enable);
int32_t sm_profile_kernel_count(void);
const char* sm_profile_kernel_name(int32_t id);
int sm_profile_report(double* total_ms, uint64_t* launches, int32_t n);

/* Work counters of the last Integrate(): surfel slots swept (N), surfels that projected
 * into the image (V), sum of the supporting counts (S), new surfels (M). Synchronises. */
int sm_frame_counters(sm_reconstruction* r, void* stream, uint64_t out[4]);

/* Diagnostics: device timeline. After sm_timeline_enable(r, frames) every kernel of the frame
 * pipeline stamps the start of its first block and the end of its last warp (%globaltimer,
 * nanoseconds) into slot [frame_index % frames][kernel id]; this shows the pipeline as it runs on
 * the GPU (events and profilers serialise it). sm_timeline_read copies frames x
 * sm_profile_kernel_count() x {start, end} values (start = UINT64_MAX: not launched).
 * frames = 0 disables and frees the buffer. No counterpart in the reference. */
int sm_timeline_enable(sm_reconstruction* r, int32_t frames);
int sm_timeline_read(sm_reconstruction* r, uint64_t* out, int32_t frames);

#ifdef __cplusplus
}
#endif
#endif  /* SURFEL_B200_H_ */
