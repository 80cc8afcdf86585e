"""CPU-side checks of the oracle pieces added in round 2 (oracle/cpu_walk.c):
  * cw_median_filter_and_densify against a line-by-line Python port of
    MedianFilterAndDensifyDepthMap (APP/main.cc:207-252) - the reference has no golden vector for it,
    so the C restatement is pinned by this independent port on small cases;
  * cw_associate_events: the supporter sets it lists are consistent with its own rasters, on the golden
    state produced by the reference's kernels."""
import numpy as np

from oracle import cpu_walk
from tests.util import INVALID, golden_camera, golden_params


def median_densify_python(depth):
    """APP/main.cc:207-252, literally (std::sort, float sum / size, fabs of the two middle candidates)."""
    H, W = depth.shape
    out = np.empty_like(depth)
    for y in range(H):
        for x in range(W):
            values = []
            for dy in range(max(0, y - 1), min(H - 1, y + 1) + 1):
                for dx in range(max(0, x - 1), min(W - 1, x + 1) + 1):
                    if depth[dy, dx] != 0:
                        values.append(int(depth[dy, dx]))
            if len(values) >= 2:
                values.sort()
                n = len(values)
                if n % 2 == 0:
                    total = np.float32(0)
                    for v in values:
                        total = np.float32(total + np.float32(v))
                    average = np.float32(total / np.float32(n))
                    prev_diff = abs(np.float32(np.float32(values[n // 2 - 1]) - average))
                    next_diff = abs(np.float32(np.float32(values[n // 2]) - average))
                    out[y, x] = values[n // 2 - 1] if prev_diff < next_diff else values[n // 2]
                else:
                    out[y, x] = values[n // 2]
            else:
                out[y, x] = depth[y, x]
    return out


def test_median_densify_matches_the_python_port():
    rng = np.random.RandomState(3)
    for shape, hole in (((17, 23), 0.4), ((9, 9), 0.8), ((12, 31), 0.1), ((3, 2), 0.5), ((1, 7), 0.3)):
        depth = rng.randint(1, 65535, size=shape).astype(np.uint16)
        depth[rng.rand(*shape) < hole] = 0
        depth[rng.rand(*shape) < 0.1] = 65535
        # close pairs make the even-count "closer to the average" rule bite
        depth[::2, ::3] = np.where(depth[::2, ::3] != 0, 1000 + (depth[::2, ::3] % 3), 0)
        expect = median_densify_python(depth)
        got = cpu_walk.median_filter_and_densify(depth)
        assert np.array_equal(got, expect), shape
    # the filter fills a hole that has two valid neighbours and keeps an isolated pixel
    d = np.zeros((5, 5), np.uint16)
    d[2, 1], d[2, 3] = 1000, 1004
    out = cpu_walk.median_filter_and_densify(d)
    assert out[2, 2] == 1004 and out[0, 0] == 0 and out[2, 1] == 1000  # tie -> upper middle element; a lone pixel stays


def test_association_events_are_consistent_with_the_rasters(golden):
    W, H, fx, fy, cx, cy = golden_camera(golden)
    _, ip = golden_params(golden)
    first, last = [int(v) for v in golden["frames"]]
    frame = last - 1
    rows = golden[f"f{frame - 1}_state"]
    depth, normals = golden[f"f{frame}_pre_depth"], golden[f"f{frame}_normals"]
    T = golden["frame_T_global"][frame]
    rasters, ev_pixel, ev_key = cpu_walk.associate_events(rows, frame, fx, fy, cx, cy, T, depth, normals,
                                                          ip.sensor_noise_factor, ip.normal_compatibility_threshold_deg,
                                                          ip.depth_scaling)
    plain = cpu_walk.associate(rows, frame, fx, fy, cx, cy, T, depth, normals, ip.sensor_noise_factor,
                               ip.normal_compatibility_threshold_deg, ip.depth_scaling)
    for k in ("supporting_surfels", "supporting_surfel_counts", "conflicting_surfels", "first_surfel_depth"):
        assert np.array_equal(rasters[k].view(np.uint32), plain[k].view(np.uint32)), k
    counts = rasters["supporting_surfel_counts"].reshape(-1)
    assert len(ev_pixel) == counts.sum() > 1000
    assert np.array_equal(np.bincount(ev_pixel, minlength=W * H), counts), "one event per counted association"
    idx = ev_key & 0x7FFFFFFF
    assert idx.max() < rows.shape[1]
    sup = rasters["supporting_surfels"].reshape(-1)
    assert np.array_equal(sup != INVALID, counts > 0)
    # the walk's canonical winner (primary before secondary, then lowest index) is a member of the set
    order = np.lexsort((ev_key, ev_pixel))
    first_of_pixel = np.concatenate([[True], np.diff(ev_pixel[order]) != 0])
    winners = dict(zip(ev_pixel[order][first_of_pixel].tolist(), (ev_key[order][first_of_pixel] & 0x7FFFFFFF).tolist()))
    check = np.flatnonzero(counts > 0)
    assert all(winners[int(p)] == int(sup[p]) for p in check)
