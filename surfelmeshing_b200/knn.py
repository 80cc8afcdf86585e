"""Host-side mirror of the meshing thread's neighbour search.

Reference: `CompressedOctree::FindNearestSurfelsWithinRadius<include_completed_surfels, include_free_surfels>`
(applications/surfel_meshing/src/surfel_meshing/octree.h:471, octree.cc:313-470), called once per surfel by
`SurfelMeshing::TriangulateSurfel` (surfel_meshing.cc:421) and `ResetSurfelsForRemeshing` (surfel_meshing.cc:821).
`SurfelKnnIndex` answers the same question for a batch of queries on the GPU through the `sm_knn_*` entry points of
libsurfel_b200.so; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

MESHING_STATE_FREE = 0        # Surfel::MeshingState, surfel.h:67-71
MESHING_STATE_FRONT = 1
MESHING_STATE_COMPLETED = 2
MESHING_STATE_ABSENT = 255    # the slot holds no surfel


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream_handle(stream):
    if stream is None:
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    return C.c_void_p(stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream))


class SurfelKnnIndex:
    """Hashed uniform grid over one snapshot of the surfel cloud + batched radius k-NN queries."""

    def __init__(self, max_points: int):
        self.lib = _lib.load_product()
        self._h = C.c_void_p()
        self.lib.call("knn_create", C.byref(self._h), int(max_points))
        self.max_points = int(max_points)
        self.point_count = 0

    def close(self):
        if self._h:
            self.lib.fn["knn_destroy"](self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build(self, x, y, z, cell_size: float, radius_squared=None, state=None, stream=None):
        """x, y, z: float32 CUDA tensors; points with radius_squared <= 0 or state == 255 are left out."""
        for t in (x, y, z, radius_squared):
            assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous())
        assert state is None or (state.is_cuda and state.dtype == torch.uint8 and state.is_contiguous())
        self.point_count = int(x.numel())
        self.lib.call("knn_build", self._h, _stream_handle(stream), self.point_count, _ptr(x), _ptr(y), _ptr(z),
                      _ptr(radius_squared), _ptr(state), float(cell_size))

    def build_from_reconstruction(self, reconstruction, cell_size: float, stream=None) -> int:
        """Indexes the handle's current surfels (smooth positions, radius_squared > 0); returns surfels_size()."""
        n = C.c_uint32()
        self.lib.call("knn_build_from_reconstruction", self._h, reconstruction._h, _stream_handle(stream),
                      float(cell_size), C.byref(n))
        self.point_count = n.value
        return n.value

    def batch_host(self, x, y, z, radius_squared, radius_factor_squared: float, max_result_count: int = 64,
                   cell_size: float = 0.0, stream=None, out=None):
        """sm_knn_batch_host: numpy arrays in (the CUDASurfelBuffersCPU arrays), numpy arrays out:
        (distances_squared [N, k] f32, indices [N, k] u32, counts [N] i32); `out` = such a triple to fill (a meshing
        thread reuses its arrays from iteration to iteration)."""
        import numpy as np
        arrays = [np.ascontiguousarray(a, np.float32) for a in (x, y, z, radius_squared)]
        n, k = len(arrays[0]), int(max_result_count)
        if out is not None:
            d2, idx, cnt = out
            assert d2.shape == (n, k) and idx.shape == (n, k) and cnt.shape == (n,)
            assert d2.dtype == np.float32 and idx.dtype == np.uint32 and cnt.dtype == np.int32
            assert d2.flags.c_contiguous and idx.flags.c_contiguous and cnt.flags.c_contiguous
        else:
            d2 = np.empty((n, k), np.float32)
            idx = np.empty((n, k), np.uint32)
            cnt = np.empty((n,), np.int32)
        self.lib.call("knn_batch_host", self._h, _stream_handle(stream), n, *[C.c_void_p(a.ctypes.data) for a in arrays],
                      float(radius_factor_squared), float(cell_size), k, C.c_void_p(d2.ctypes.data), C.c_void_p(idx.ctypes.data),
                      C.c_void_p(cnt.ctypes.data))
        self.point_count = n
        return d2, idx, cnt

    def FindNearestSurfelsWithinRadius(self, qx, qy, qz, radius_squared, max_result_count: int, state=None,
                                       include_completed_surfels: bool = True, include_free_surfels: bool = True,
                                       stream=None):
        """Batch form of octree.cc:433-470. Returns (distances_squared [Q, k], indices [Q, k] (int64 view of
        u32), counts [Q]) as CUDA tensors; entries past a query's count are +inf / 0xFFFFFFFF."""
        q = int(qx.numel())
        k = int(max_result_count)
        dev = qx.device
        d2 = torch.empty((q, k), dtype=torch.float32, device=dev)
        idx = torch.empty((q, k), dtype=torch.int32, device=dev)
        cnt = torch.empty((q,), dtype=torch.int32, device=dev)
        self.lib.call("knn_query", self._h, _stream_handle(stream), q, _ptr(qx), _ptr(qy), _ptr(qz),
                      _ptr(radius_squared), _ptr(state), int(bool(include_completed_surfels)),
                      int(bool(include_free_surfels)), k, _ptr(d2), _ptr(idx), _ptr(cnt))
        return d2, idx, cnt
