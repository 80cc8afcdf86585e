"""Host-side mirror of the reference interface for the surfel reconstruction hot path.

`CUDASurfelReconstruction` has the public surface of the reference class of the same
name (applications/surfel_meshing/src/surfel_meshing/cuda_surfel_reconstruction.h:44-176):
Integrate / Regularize / TransferAllToCPU / ExportVertices / GetTimings / surfel_count /
surfels_size, same argument order and meaning. The depth pre-processing free functions
keep the reference's names (cuda_depth_processing.cuh:43-122). Everything forwards to the
C ABI (include/surfel_b200.h); PyTorch only provides device memory and streams.

The same class drives the parity oracle when constructed with
`lib=load_reference_oracle()` (tests / bench reference arm only).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import (IntegrateParams, Library, PreprocessParams, StreamDesc, StreamStats, TransferStats, TransferToken,
                   VisualizationParams)

BUFFER_NAMES = ["surfel_x_buffer", "surfel_y_buffer", "surfel_z_buffer", "surfel_radius_squared_buffer",
                "surfel_normal_x_buffer", "surfel_normal_y_buffer", "surfel_normal_z_buffer",
                "surfel_last_update_stamp_buffer"]


def make_cpu_buffers(max_surfel_count: int, pinned: bool = False) -> dict:
    """The eight arrays of CUDASurfelBuffersCPU (APP/cuda_surfels_cpu.h:40-73), max_surfel_count long."""
    out = {}
    for k in BUFFER_NAMES:
        dtype = np.uint32 if "stamp" in k else np.float32
        if pinned:
            t = torch.empty(max(max_surfel_count, 1), dtype=torch.int32 if "stamp" in k else torch.float32).pin_memory()
            out[k] = t.numpy().view(dtype)
            out["_keep_" + k] = t
        else:
            out[k] = np.zeros(max(max_surfel_count, 1), dtype=dtype)
    return out

ROW_NAMES = [
    "x", "y", "z", "smooth_x", "smooth_y", "smooth_z", "confidence", "radius_squared",
    "normal_x", "normal_y", "normal_z", "gradient_x", "gradient_y", "gradient_z",
    "accum_x", "accum_y", "accum_z", "creation_stamp", "last_update_stamp",
    "neighbor0", "neighbor1", "neighbor2", "neighbor3", "gradient_count", "color",
]
# Rows that hold scratch data between calls and are excluded from state comparisons:
# gradient (11-13), accum (14-16, never written), gradient weight sum (23).
SCRATCH_ROWS = (11, 12, 13, 14, 15, 16, 23)


def _stream_handle(stream) -> int:
    if stream is None:
        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, torch.cuda.Stream):
        return stream.cuda_stream
    return int(stream)


def _raster(t: torch.Tensor, channels: int = 1):
    """(device pointer, pitch in bytes) of a row-pitched raster tensor [H, W(, C)]."""
    if not t.is_cuda:
        raise ValueError("rasters must be CUDA tensors (there is no CPU path)")
    if channels == 1:
        assert t.dim() == 2 and t.stride(1) == 1, "expected [H, W] raster with unit pixel stride"
    else:
        assert t.dim() == 3 and t.shape[2] == channels and t.stride(2) == 1 and t.stride(1) == channels
    return C.c_void_p(t.data_ptr()), t.stride(0) * t.element_size()


def _mat12(m) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(m, dtype=np.float32).reshape(-1)[:12])
    assert a.size == 12
    return a


def invert_rigid(m) -> np.ndarray:
    """Inverse of a 3x4 rigid transform, computed in float64 and rounded to float32 once."""
    a = np.asarray(m, dtype=np.float64).reshape(3, 4)
    R, t = a[:, :3], a[:, 3]
    return np.concatenate([R.T, (-R.T @ t)[:, None]], axis=1).astype(np.float32)


# ---------------------------------------------------------------------------------------
# depth pre-processing (names of cuda_depth_processing.cuh)
# ---------------------------------------------------------------------------------------

def BilateralFilteringAndDepthCutoffCUDA(stream, sigma_xy, sigma_value_factor, value_to_ignore, radius_factor,
                                         max_depth, depth_valid_region_radius, input_depth, output_depth,
                                         lib: Optional[Library] = None):
    lib = lib or _lib.load_product()
    ip, ipitch = _raster(input_depth)
    op, opitch = _raster(output_depth)
    H, W = input_depth.shape
    lib.call("bilateral_filter_and_depth_cutoff", _stream_handle(stream), sigma_xy, sigma_value_factor,
             int(value_to_ignore), radius_factor, int(max_depth), depth_valid_region_radius, W, H, ip, ipitch,
             op, opitch)


def _others_args(other_depths: Sequence[torch.Tensor], others_TR_reference):
    K = len(other_depths)
    ptrs = (C.c_void_p * K)(*[t.data_ptr() for t in other_depths])
    pitches = (C.c_size_t * K)(*[t.stride(0) * t.element_size() for t in other_depths])
    mats = np.ascontiguousarray(np.asarray(others_TR_reference, dtype=np.float32).reshape(K, 12))
    return K, ptrs, pitches, mats


def OutlierDepthMapFusionCUDA(stream, tolerance, input_depth, depth_fx, depth_fy, depth_cx, depth_cy, other_depths,
                              others_TR_reference, output_depth, required_count: int = -1,
                              lib: Optional[Library] = None):
    """Both reference overloads: required_count = -1 means 'all other frames must agree'."""
    lib = lib or _lib.load_product()
    K, ptrs, pitches, mats = _others_args(other_depths, others_TR_reference)
    ip, ipitch = _raster(input_depth)
    op, opitch = _raster(output_depth)
    H, W = input_depth.shape
    lib.call("outlier_depth_map_fusion", _stream_handle(stream), K, required_count, tolerance, depth_fx, depth_fy,
             depth_cx, depth_cy, W, H, ip, ipitch, ptrs, pitches, mats.ctypes.data_as(C.c_void_p), op, opitch)


def MedianFilterAndDensifyDepthMap(stream, iterations, input_depth, output_depth=None, lib: Optional[Library] = None):
    """APP/main.cc:207-252 (there on the CPU, main.cc:927-939): `iterations` passes of the 3x3
    zero-excluding median that also fills holes. Returns the output tensor."""
    lib = lib or _lib.load_product()
    H, W = input_depth.shape
    if output_depth is None:
        output_depth = torch.zeros((H, W), dtype=torch.uint16, device=input_depth.device)
    scratch = torch.zeros((H, W), dtype=torch.uint16, device=input_depth.device)
    ip, ipitch = _raster(input_depth)
    op, opitch = _raster(output_depth)
    sp, spitch = _raster(scratch)
    lib.call("median_filter_and_densify_depth_map", _stream_handle(stream), int(iterations), W, H, ip, ipitch, op, opitch,
             sp, spitch)
    return output_depth


def ErodeDepthMapCUDA(stream, radius, input_depth, output_depth, lib: Optional[Library] = None):
    lib = lib or _lib.load_product()
    ip, ipitch = _raster(input_depth)
    op, opitch = _raster(output_depth)
    H, W = input_depth.shape
    lib.call("erode_depth_map", _stream_handle(stream), radius, W, H, ip, ipitch, op, opitch)


def CopyWithoutBorderCUDA(stream, input_depth, output_depth, lib: Optional[Library] = None):
    ErodeDepthMapCUDA(stream, 0, input_depth, output_depth, lib=lib)


def ComputeNormalsAndDropBadPixelsCUDA(stream, observation_angle_threshold_deg, depth_scaling, depth_fx, depth_fy,
                                       depth_cx, depth_cy, in_depth, out_depth, out_normals,
                                       lib: Optional[Library] = None):
    lib = lib or _lib.load_product()
    ip, ipitch = _raster(in_depth)
    op, opitch = _raster(out_depth)
    np_, npitch = _raster(out_normals, 2)
    H, W = in_depth.shape
    lib.call("compute_normals_and_drop_bad_pixels", _stream_handle(stream), observation_angle_threshold_deg,
             depth_scaling, depth_fx, depth_fy, depth_cx, depth_cy, W, H, ip, ipitch, op, opitch, np_, npitch)


def ComputePointRadiiAndRemoveIsolatedPixelsCUDA(stream, point_radius_extension_factor, point_radius_clamp_factor,
                                                 depth_scaling, depth_fx, depth_fy, depth_cx, depth_cy, depth_buffer,
                                                 radius_buffer, out_depth, lib: Optional[Library] = None):
    lib = lib or _lib.load_product()
    ip, ipitch = _raster(depth_buffer)
    rp, rpitch = _raster(radius_buffer)
    op, opitch = _raster(out_depth)
    H, W = depth_buffer.shape
    lib.call("compute_point_radii_and_remove_isolated_pixels", _stream_handle(stream),
             point_radius_extension_factor, point_radius_clamp_factor, depth_scaling, depth_fx, depth_fy, depth_cx,
             depth_cy, W, H, ip, ipitch, rp, rpitch, op, opitch)


# ---------------------------------------------------------------------------------------
# CUDASurfelReconstruction
# ---------------------------------------------------------------------------------------

class CUDASurfelReconstruction:
    """Mirror of vis::CUDASurfelReconstruction (cuda_surfel_reconstruction.h:44-176).

    The constructor takes the camera as (width, height, fx, fy, cx, cy) with cx, cy in the
    reference's pixel-corner convention (PinholeCamera4f::parameters()); the three OpenGL
    resources and the render window of the reference constructor are GUI-only and omitted.
    """

    def __init__(self, max_surfel_count: int, width: int, height: int, fx: float, fy: float, cx: float, cy: float,
                 lib: Optional[Library] = None):
        self.lib = lib or _lib.load_product()
        self.width, self.height = int(width), int(height)
        self.fx, self.fy, self.cx, self.cy = float(fx), float(fy), float(cx), float(cy)
        self.max_surfel_count = int(max_surfel_count)
        handle = C.c_void_p()
        self.lib.call("create", C.byref(handle), self.max_surfel_count, self.width, self.height, self.fx, self.fy,
                      self.cx, self.cy)
        self._h = handle

    def close(self):
        if getattr(self, "_h", None):
            self.lib.fn["destroy"](self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference API ------------------------------------------------------------------
    def Integrate(self, stream, frame_index, depth_scaling, depth_buffer, normals_buffer, radius_buffer,
                  color_buffer, global_T_local, sensor_noise_factor, max_surfel_confidence, regularizer_weight,
                  regularization_frame_window_size, do_blending, measurement_blending_radius,
                  regularization_iterations_per_integration_iteration, radius_factor_for_regularization_neighbors,
                  normal_compatibility_threshold_deg, surfel_integration_active_window_size,
                  local_T_global=None):
        """cuda_surfel_reconstruction.cc:112-320. `depth_buffer` is blended in place.

        `local_T_global` defaults to the float64 inverse of `global_T_local` rounded to fp32
        (the reference inverts with Sophus on the host)."""
        p = IntegrateParams(depth_scaling, sensor_noise_factor, max_surfel_confidence, regularizer_weight,
                            regularization_frame_window_size, 1 if do_blending else 0, measurement_blending_radius,
                            regularization_iterations_per_integration_iteration,
                            radius_factor_for_regularization_neighbors, normal_compatibility_threshold_deg,
                            surfel_integration_active_window_size)
        self.integrate(stream, frame_index, p, depth_buffer, normals_buffer, radius_buffer, color_buffer,
                       global_T_local, local_T_global)

    def integrate(self, stream, frame_index, params: IntegrateParams, depth_buffer, normals_buffer, radius_buffer,
                  color_buffer, global_T_local, local_T_global=None):
        g = _mat12(global_T_local)
        l = _mat12(local_T_global if local_T_global is not None else invert_rigid(g))
        dp, dpitch = _raster(depth_buffer)
        np_, npitch = _raster(normals_buffer, 2)
        rp, rpitch = _raster(radius_buffer)
        cp, cpitch = _raster(color_buffer, 3)
        self.lib.call("integrate", self._h, _stream_handle(stream), int(frame_index), C.byref(params), dp, dpitch,
                      np_, npitch, rp, rpitch, cp, cpitch, g.ctypes.data_as(C.c_void_p),
                      l.ctypes.data_as(C.c_void_p))

    def Regularize(self, stream, frame_index, regularizer_weight, radius_factor_for_regularization_neighbors,
                   regularization_frame_window_size):
        """cuda_surfel_reconstruction.cc:322-337."""
        self.lib.call("regularize", self._h, _stream_handle(stream), int(frame_index), regularizer_weight,
                      radius_factor_for_regularization_neighbors, regularization_frame_window_size)

    def TransferAllToCPU(self, stream, frame_index, buffers: Optional[dict] = None) -> dict:
        """cuda_surfel_reconstruction.cc:339-359: fills the CUDASurfelBuffersCPU arrays
        (cuda_surfels_cpu.h:40-73) and returns them (synchronises the stream)."""
        n = self.surfels_size()
        names = ["surfel_x_buffer", "surfel_y_buffer", "surfel_z_buffer", "surfel_radius_squared_buffer",
                 "surfel_normal_x_buffer", "surfel_normal_y_buffer", "surfel_normal_z_buffer",
                 "surfel_last_update_stamp_buffer"]
        if buffers is None:
            buffers = {k: np.empty(max(n, 1), dtype=np.uint32 if "stamp" in k else np.float32) for k in names}
        count = C.c_uint64()
        ptrs = [buffers[k].ctypes.data_as(C.c_void_p) for k in names]
        self.lib.call("transfer_all_to_cpu", self._h, _stream_handle(stream), int(frame_index), *ptrs,
                      C.byref(count))
        torch.cuda.synchronize()
        buffers["frame_index"] = int(frame_index)
        buffers["surfel_count"] = int(count.value)
        return buffers

    def TransferDeltaToCPU(self, stream, frame_index, buffers: dict, token: TransferToken) -> TransferStats:
        """sm_transfer_delta_to_cpu: brings `buffers` (filled by the transfer `token` stands for; a fresh
        TransferToken() = never) up to date; afterwards they equal a full TransferAllToCPU. The arrays
        must be at least surfels_size() long (the reference allocates max_surfel_count)."""
        stats = TransferStats()
        ptrs = [buffers[k].ctypes.data_as(C.c_void_p) for k in BUFFER_NAMES]
        self.lib.call("transfer_delta_to_cpu", self._h, _stream_handle(stream), int(frame_index), C.byref(token), *ptrs,
                      C.byref(stats))
        buffers["frame_index"] = int(frame_index)
        buffers["surfel_count"] = int(stats.surfel_count)
        return stats

    def UpdateVisualizationBuffers(self, stream, frame_index, latest_triangulated_frame_index, latest_mesh_surfel_count,
                                   surfel_integration_active_window_size, visualize_last_update_timestamp,
                                   visualize_creation_timestamp, visualize_radii, visualize_normals,
                                   vertex_buffer=None, neighbor_index_buffer=None, normal_vertex_buffer=None,
                                   point_size_in_floats: int = 4):
        """cuda_surfel_reconstruction.cc:361-403; the three OpenGL buffers of the reference are plain
        device tensors here (None = not wanted)."""
        p = VisualizationParams(int(frame_index), int(latest_triangulated_frame_index), int(latest_mesh_surfel_count),
                                int(surfel_integration_active_window_size), int(point_size_in_floats),
                                int(bool(visualize_last_update_timestamp)), int(bool(visualize_creation_timestamp)),
                                int(bool(visualize_radii)), int(bool(visualize_normals)))
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        self.lib.call("update_visualization_buffers", self._h, _stream_handle(stream), C.byref(p), ptr(vertex_buffer),
                      ptr(neighbor_index_buffer), ptr(normal_vertex_buffer))

    def ExportVertices(self, stream, position_buffer: torch.Tensor, color_buffer: torch.Tensor):
        """cuda_surfel_reconstruction.cc:405-410."""
        self.lib.call("export_vertices", self._h, _stream_handle(stream), C.c_void_p(position_buffer.data_ptr()),
                      C.c_void_p(color_buffer.data_ptr()))

    def GetTimings(self):
        """cuda_surfel_reconstruction.cc:412-429: (data_association, surfel_merging,
        measurement_blending, integration, neighbor_update, new_surfel_creation,
        regularization) in milliseconds."""
        out = (C.c_float * 7)()
        self.lib.call("get_timings", self._h, C.byref(out))
        return tuple(out)

    def enable_timings(self, enable=True):
        self.lib.call("enable_timings", self._h, 1 if enable else 0)

    def surfel_count(self) -> int:
        v = C.c_uint32()
        self.lib.call("surfel_count", self._h, C.byref(v))
        return v.value

    def surfels_size(self) -> int:
        v = C.c_uint32()
        self.lib.call("surfels_size", self._h, C.byref(v))
        return v.value

    # -- extras: fused pre-processing, state access, stream runner --------------------------
    def configure(self, key: str, value: float):
        """sm_configure: named tuning knobs (product only), e.g. "tiebreak_wave"."""
        self.lib.call("configure", self._h, key.encode(), float(value))

    def reset(self, stream=None):
        self.lib.call("reset", self._h, _stream_handle(stream))

    def preprocess(self, stream, params: PreprocessParams, raw_depth, other_depths, others_TR_reference, out_depth,
                   out_normals, out_radius):
        """The pre-processing call sequence of APP/main.cc:1015-1191 in one call."""
        K, ptrs, pitches, mats = _others_args(other_depths, others_TR_reference)
        assert K == params.outlier_filtering_frame_count
        rp, rpitch = _raster(raw_depth)
        op, opitch = _raster(out_depth)
        np_, npitch = _raster(out_normals, 2)
        radp, radpitch = _raster(out_radius)
        self.lib.call("preprocess", self._h, _stream_handle(stream), C.byref(params), rp, rpitch, ptrs, pitches,
                      mats.ctypes.data_as(C.c_void_p), op, opitch, np_, npitch, radp, radpitch)

    def dump_state(self, stream=None):
        """Returns (rows[25, n] float32 view of the SoA, surfels_size, merge_count)."""
        n = self.surfels_size()
        rows = np.zeros((_lib.ROW_COUNT, max(n, 1)), dtype=np.float32)
        size, merges = C.c_uint32(), C.c_uint32()
        self.lib.call("dump_state", self._h, _stream_handle(stream), rows.ctypes.data_as(C.c_void_p), rows.shape[1],
                      C.byref(size), C.byref(merges))
        assert size.value == n
        return rows[:, :n], n, merges.value

    def load_state(self, rows: np.ndarray, merge_count: int, stream=None):
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        assert rows.shape[0] == _lib.ROW_COUNT
        n = rows.shape[1]
        if n == 0:
            rows = np.zeros((_lib.ROW_COUNT, 1), dtype=np.float32)
        self.lib.call("load_state", self._h, _stream_handle(stream), rows.ctypes.data_as(C.c_void_p), rows.shape[1],
                      n, int(merge_count))

    def download_rasters(self, stream=None) -> dict:
        P = self.width * self.height
        out = {
            "supporting_surfels": np.empty(P, np.uint32), "supporting_surfel_counts": np.empty(P, np.uint32),
            "supporting_surfel_depth_sums": np.empty(P, np.float32), "conflicting_surfels": np.empty(P, np.uint32),
            "first_surfel_depth": np.empty(P, np.float32), "new_surfel_flag_vector": np.empty(P, np.uint8),
            "new_surfel_indices": np.empty(P, np.uint32),
        }
        self.lib.call("download_rasters", self._h, _stream_handle(stream),
                      *[v.ctypes.data_as(C.c_void_p) for v in out.values()])
        return {k: v.reshape(self.height, self.width) for k, v in out.items()}

    def stream_run(self, stream, depth, color, global_T_frame, frame_T_global, others_TR_reference,
                   pp: PreprocessParams, ip: IntegrateParams, first_frame: int, last_frame: int) -> StreamStats:
        """Frame loop of APP/main.cc:885-1223 over frames [first_frame, last_frame).

        depth [F,H,W] uint16 and color [F,H,W,3] uint8 are either CUDA tensors (device-resident
        stream) or pinned CPU tensors (uploaded frame by frame inside the call)."""
        on_host = not depth.is_cuda
        assert depth.is_contiguous() and color.is_contiguous() and color.is_cuda == depth.is_cuda
        if on_host:
            assert depth.is_pinned() and color.is_pinned(), "host frames must be in pinned memory"
        F = depth.shape[0]
        g = np.ascontiguousarray(np.asarray(global_T_frame, np.float32).reshape(F, 12))
        l = np.ascontiguousarray(np.asarray(frame_T_global, np.float32).reshape(F, 12))
        o = np.ascontiguousarray(np.asarray(others_TR_reference, np.float32).reshape(F, -1, 12))
        assert o.shape[1] == pp.outlier_filtering_frame_count
        desc = StreamDesc(self.width, self.height, F, 1 if on_host else 0, depth.data_ptr(), color.data_ptr(),
                          g.ctypes.data, l.ctypes.data, o.ctypes.data)
        stats = StreamStats()
        self.lib.call("stream_run", self._h, _stream_handle(stream), C.byref(desc), C.byref(pp), C.byref(ip),
                      int(first_frame), int(last_frame), C.byref(stats))
        return stats
