"""ctypes loader for oracle/_ref/libcpu_walk.so (oracle/cpu_walk.c) — TEST INFRASTRUCTURE.

CPU restatement of the depth filter chain (a1-a5) and of the min-depth + association loop
(a7/a8); see the header of cpu_walk.c for the reference lines each function follows.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
LIB_PATH = ORACLE_DIR / "_ref" / "libcpu_walk.so"

_lib = None


def build():
    subprocess.run(["make", "-C", str(ORACLE_DIR), "_ref/libcpu_walk.so"], check=True, capture_output=True)


def load():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        _lib = C.CDLL(str(LIB_PATH))
        _lib.cw_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def max_threads() -> int:
    return load().cw_max_threads()


def set_threads(n: int):
    load().cw_set_threads(C.c_int(n))


def bilateral(raw, sigma_xy, sigma_value_factor, radius_factor, max_depth, valid_radius, value_to_ignore=0):
    H, W = raw.shape
    out = np.empty_like(raw)
    load().cw_bilateral(C.c_float(sigma_xy), C.c_float(sigma_value_factor), C.c_uint16(value_to_ignore),
                        C.c_float(radius_factor), C.c_uint16(max_depth), C.c_float(valid_radius), W, H, _p(raw), _p(out))
    return out


def _other_ptrs(others):
    arr = [np.ascontiguousarray(o) for o in others]
    return arr, (C.c_void_p * len(arr))(*[a.ctypes.data for a in arr])


def outlier(depth, others, mats, tolerance, fx, fy, cx, cy, required_count=-1):
    H, W = depth.shape
    out = np.empty_like(depth)
    keep, ptrs = _other_ptrs(others)
    mats = np.ascontiguousarray(mats, dtype=np.float32)
    load().cw_outlier(len(keep), required_count, C.c_float(tolerance), C.c_float(fx), C.c_float(fy), C.c_float(cx),
                      C.c_float(cy), W, H, _p(depth), ptrs, _p(mats), _p(out))
    return out


def erode(depth, radius):
    H, W = depth.shape
    out = np.empty_like(depth)
    load().cw_erode(radius, W, H, _p(depth), _p(out))
    return out


def normals(depth, angle_deg, depth_scaling, fx, fy, cx, cy):
    H, W = depth.shape
    out = np.empty_like(depth)
    nrm = np.empty((H, W, 2), dtype=np.float32)
    load().cw_normals(C.c_float(angle_deg), C.c_float(depth_scaling), C.c_float(fx), C.c_float(fy), C.c_float(cx),
                      C.c_float(cy), W, H, _p(depth), _p(out), _p(nrm))
    return out, nrm


def radii(depth, ext, clamp, depth_scaling, fx, fy, cx, cy):
    H, W = depth.shape
    out = np.empty_like(depth)
    rad = np.zeros((H, W), dtype=np.float32)
    load().cw_radii(C.c_float(ext), C.c_float(clamp), C.c_float(depth_scaling), C.c_float(fx), C.c_float(fy),
                    C.c_float(cx), C.c_float(cy), W, H, _p(depth), _p(rad), _p(out))
    return out, rad


def preprocess(pp, fx, fy, cx, cy, raw, others, mats):
    """pp: surfelmeshing_b200._lib.PreprocessParams (same layout as cw_preprocess_params)."""
    H, W = raw.shape
    keep, ptrs = _other_ptrs(others)
    mats = np.ascontiguousarray(mats, dtype=np.float32)
    A, B, out = np.empty_like(raw), np.empty_like(raw), np.empty_like(raw)
    nrm = np.empty((H, W, 2), dtype=np.float32)
    rad = np.zeros((H, W), dtype=np.float32)
    load().cw_preprocess(C.byref(pp), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), W, H, _p(raw), ptrs,
                         _p(mats), _p(A), _p(B), _p(out), _p(nrm), _p(rad))
    return out, nrm, rad


def associate(rows, frame_index, fx, fy, cx, cy, local_T_global, depth, nrm, sensor_noise_factor=0.05,
              normal_threshold_deg=40.0, depth_scaling=5000.0, active_window=2**31 - 1):
    """rows: [25, n] float32 SoA. Returns the five association rasters."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    n = rows.shape[1]
    H, W = depth.shape
    P = H * W
    out = dict(supporting_surfels=np.empty(P, np.uint32), supporting_surfel_counts=np.empty(P, np.uint32),
               supporting_surfel_depth_sums=np.empty(P, np.float32), conflicting_surfels=np.empty(P, np.uint32),
               first_surfel_depth=np.empty(P, np.float32))
    T = np.ascontiguousarray(local_T_global, dtype=np.float32).reshape(-1)
    nrm = np.ascontiguousarray(nrm, dtype=np.float32)
    depth = np.ascontiguousarray(depth)
    load().cw_associate(_p(rows), C.c_size_t(rows.shape[1]), C.c_uint32(n), C.c_uint32(frame_index), C.c_int(active_window),
                        C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), _p(T), C.c_float(sensor_noise_factor),
                        C.c_float(normal_threshold_deg), C.c_float(depth_scaling), W, H, _p(depth), _p(nrm),
                        *[_p(v) for v in out.values()])
    return {k: v.reshape(H, W) for k, v in out.items()}


def associate_events(rows, frame_index, fx, fy, cx, cy, local_T_global, depth, nrm, sensor_noise_factor=0.05,
                     normal_threshold_deg=40.0, depth_scaling=5000.0, active_window=2**31 - 1):
    """Like `associate`, plus the supporter sets: returns (rasters, event_pixel, event_key) where every
    event is one (pixel, surfel) association that reaches the reference's atomicCAS
    (kernels.cu:1688); event_key = surfel index | 0x80000000 for a secondary-pixel association."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    n = rows.shape[1]
    H, W = depth.shape
    P = H * W
    out = dict(supporting_surfels=np.empty(P, np.uint32), supporting_surfel_counts=np.empty(P, np.uint32),
               supporting_surfel_depth_sums=np.empty(P, np.float32), conflicting_surfels=np.empty(P, np.uint32),
               first_surfel_depth=np.empty(P, np.float32))
    T = np.ascontiguousarray(local_T_global, dtype=np.float32).reshape(-1)
    nrm = np.ascontiguousarray(nrm, dtype=np.float32)
    depth = np.ascontiguousarray(depth)
    max_events = 2 * n + 16
    ev_pixel = np.empty(max_events, np.uint32)
    ev_key = np.empty(max_events, np.uint32)
    count = C.c_uint64(0)
    load().cw_associate_events(_p(rows), C.c_size_t(rows.shape[1]), C.c_uint32(n), C.c_uint32(frame_index),
                               C.c_int(active_window), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), _p(T),
                               C.c_float(sensor_noise_factor), C.c_float(normal_threshold_deg), C.c_float(depth_scaling),
                               W, H, _p(depth), _p(nrm), *[_p(v) for v in out.values()], _p(ev_pixel), _p(ev_key),
                               C.c_uint64(max_events), C.byref(count))
    m = min(int(count.value), max_events)
    return {k: v.reshape(H, W) for k, v in out.items()}, ev_pixel[:m], ev_key[:m]


def median_filter_and_densify(depth):
    """One pass of MedianFilterAndDensifyDepthMap (APP/main.cc:207-252) on a [H, W] uint16 array."""
    depth = np.ascontiguousarray(depth, dtype=np.uint16)
    H, W = depth.shape
    out = np.empty_like(depth)
    load().cw_median_filter_and_densify(W, H, _p(depth), _p(out))
    return out
