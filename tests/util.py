"""Helpers shared by the parity tests."""
import numpy as np

from surfelmeshing_b200._lib import IntegrateParams, PreprocessParams

# Rows of the surfel SoA that Integrate() determines from deterministic rasters only
# (everything except smooth positions, neighbour links and scratch rows).
INTEGRATE_ROWS = (0, 1, 2, 6, 7, 8, 9, 10, 17, 18, 24)
SMOOTH_ROWS = (3, 4, 5)
NEIGHBOR_ROWS = (19, 20, 21, 22)
INVALID = 0xFFFFFFFF


def bits(a):
    a = np.asarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def count_mismatch(a, b, mask=None):
    ne = bits(a) != bits(b)
    if mask is not None:
        ne &= mask
    return int(ne.sum())


def golden_params(golden):
    pp = PreprocessParams.defaults()
    pp.depth_valid_region_radius = float(golden["valid_region_radius"][0])
    return pp, IntegrateParams.defaults()


def golden_camera(golden):
    W, H, fx, fy, cx, cy = golden["camera"]
    return int(W), int(H), float(fx), float(fy), float(cx), float(cy)


def other_frames(frame, K):
    half = K // 2
    return [frame - (i + 1) for i in range(half)] + [frame + (i + 1) for i in range(half)]


def check_state_invariants(rows, n):
    """Size-independent properties of a surfel SoA (rows [25, n])."""
    r2 = rows[7]
    stamps = rows[18].view(np.uint32)
    merged = r2 < 0
    assert np.all(stamps[merged] == 0), "merged surfels carry last-update stamp 0"
    assert np.all((rows[24].view(np.uint32)[merged] >> 24) == 1), "merged surfels carry the detach flag"
    live = ~merged
    assert np.isfinite(rows[0:11][:, live]).all(), "no NaN/Inf in live surfel attributes"
    nn = np.sqrt((rows[8:11][:, live] ** 2).sum(axis=0))
    assert np.all(np.abs(nn - 1) < 1e-3), "normals are unit length"
    nbr = rows[19:23].view(np.uint32)
    assert np.all((nbr == INVALID) | (nbr < n)), "neighbour links point inside the cloud"
    conf = rows[6][live]
    assert np.all(conf > 0)
