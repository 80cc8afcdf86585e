"""GPU parity of the batched radius k-NN against the reference's own CPU octree
(oracle/_ref/liboctree_ref.so = octree.cc compiled unmodified), its brute-force checker and the committed answers."""
import numpy as np
import pytest
import torch

from oracle import octree_ref
from tests import knn_cases
from tests.test_octree_oracle import FILTERS, GOLDEN_CASES, assert_same_neighbours

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_query(index, qx, qy, qz, r2, k, state=None, include_completed=True, include_free=True):
    d2, idx, cnt = index.FindNearestSurfelsWithinRadius(to_dev(qx), to_dev(qy), to_dev(qz), to_dev(r2), k,
                                                        state=None if state is None else to_dev(state),
                                                        include_completed_surfels=include_completed,
                                                        include_free_surfels=include_free)
    torch.cuda.synchronize()
    return d2.cpu().numpy(), idx.cpu().numpy().view(np.uint32), cnt.cpu().numpy()


def build(x, y, z, cell_size, state=None, radius_squared=None):
    from surfelmeshing_b200.knn import SurfelKnnIndex
    index = SurfelKnnIndex(len(x))
    index.build(to_dev(x), to_dev(y), to_dev(z), cell_size, state=None if state is None else to_dev(state),
                radius_squared=None if radius_squared is None else to_dev(radius_squared))
    return index


def assert_identical(got, want):
    for g, w in zip(got, want):
        assert np.array_equal(g.view(np.uint32) if g.dtype == np.float32 else g, w.view(np.uint32) if w.dtype == np.float32 else w)


@pytest.mark.parametrize("name,cloud,r2_value,k", GOLDEN_CASES, ids=[c[0] for c in GOLDEN_CASES])
def test_knn_matches_golden_and_brute_force(name, cloud, r2_value, k):
    """Counts and squared distances bit for bit against the committed octree answers; indices too, up to the
    order inside runs of equal distance (which the octree leaves open); everything exact against the brute-force
    restatement, which fixes that order the way the product does."""
    golden = np.load(octree_ref.LIB_PATH.parents[2] / "tests" / "golden" / "knn_golden.npz")
    x, y, z = cloud()
    state = knn_cases.states(len(x), 14)
    qi = golden[f"{name}_query_index"]
    r2 = np.full(len(qi), r2_value, np.float32)
    for cell_size in (float(np.sqrt(r2_value)), 0.37 * float(np.sqrt(r2_value)), 3.1 * float(np.sqrt(r2_value))):
        index = build(x, y, z, cell_size, state=state)
        for label, ic, fr in FILTERS:
            got = gpu_query(index, x[qi], y[qi], z[qi], r2, k, state, ic, fr)
            want = (golden[f"{name}_{label}_d2"], golden[f"{name}_{label}_idx"], golden[f"{name}_{label}_cnt"])
            assert_same_neighbours(got, want, k)
            if octree_ref.available():
                assert_identical(got, octree_ref.brute_force(x, y, z, state, x[qi], y[qi], z[qi], r2, k, ic, fr))
        index.close()


@pytest.mark.skipif(not octree_ref.available(), reason="oracle/_ref/liboctree_ref.so not built")
@pytest.mark.parametrize("k", [1, 5, 32, 33, 64])
def test_knn_live_octree(k):
    """Free query positions (not on a point), per-query radii from 0 to beyond the cell size, every cap."""
    x, y, z = knn_cases.random_cloud(20000, 21, extent=2.0)
    rng = np.random.default_rng(22)
    q = 1500
    qx, qy, qz = [(rng.random(q, dtype=np.float32) * 4.4 - 2.2).astype(np.float32) for _ in range(3)]
    r2 = (rng.random(q, dtype=np.float32) * 0.5).astype(np.float32) ** 2
    r2[:10] = 0
    r2[10:20] = -1
    tree = octree_ref.Octree(x, y, z)
    want = tree.query(qx, qy, qz, r2, k)[:3]
    tree.close()
    index = build(x, y, z, 0.25)
    got = gpu_query(index, qx, qy, qz, r2, k)
    assert_same_neighbours(got, want, k)
    assert_identical(got, octree_ref.brute_force(x, y, z, None, qx, qy, qz, r2, k))
    assert (got[2] == k).any() and (got[2] == 0).any()
    index.close()


@pytest.mark.skipif(not octree_ref.available(), reason="oracle/_ref/liboctree_ref.so not built")
def test_knn_radius_far_above_cell_size_and_duplicates():
    """A ball of thousands of cells takes the bounded all-records path; exact duplicates come out by index."""
    x, y, z = knn_cases.nasty_cloud(400, 31)
    n = len(x)
    qi = np.arange(0, n, 7)
    r2 = np.full(len(qi), 6.0 ** 2, np.float32)
    index = build(x, y, z, 0.05)
    got = gpu_query(index, x[qi], y[qi], z[qi], r2, 64)
    assert_identical(got, octree_ref.brute_force(x, y, z, None, x[qi], y[qi], z[qi], r2, 64))
    tree = octree_ref.Octree(x, y, z)
    assert_same_neighbours(got, tree.query(x[qi], y[qi], z[qi], r2, 64)[:3], 64)
    tree.close()
    # nearest two of a base point are itself and its exact duplicate, lower index first
    small = gpu_query(index, x[qi], y[qi], z[qi], np.full(len(qi), 1e-6, np.float32), 4)
    base = qi[qi % 8 == 0]
    sel = np.isin(qi, base)
    assert (small[2][sel] == 2).all() and (small[1][sel, 0] == base).all() and (small[1][sel, 1] == base + 1).all()
    index.close()


def test_knn_empty_and_rebuild():
    x, y, z = knn_cases.random_cloud(1000, 41)
    from surfelmeshing_b200.knn import SurfelKnnIndex
    index = SurfelKnnIndex(2000)
    absent = np.full(1000, 255, np.uint8)
    index.build(to_dev(x), to_dev(y), to_dev(z), 1.0, state=to_dev(absent))
    got = gpu_query(index, x[:50], y[:50], z[:50], np.full(50, 4.0, np.float32), 8)
    assert (got[2] == 0).all() and np.isinf(got[0]).all() and (got[1] == 0xFFFFFFFF).all()
    # the same index object serves the next snapshot; merged surfels (radius^2 <= 0) stay out
    r2_rows = np.ones(1000, np.float32)
    r2_rows[::2] = -1
    index.build(to_dev(x), to_dev(y), to_dev(z), 1.0, radius_squared=to_dev(r2_rows))
    got = gpu_query(index, x[:50], y[:50], z[:50], np.full(50, 9.0, np.float32), 64)
    valid = got[1] != 0xFFFFFFFF
    assert valid.any() and (got[1][valid] % 2 == 1).all()
    if octree_ref.available():
        state = np.where(r2_rows > 0, 0, 255).astype(np.uint8)
        assert_identical(got, octree_ref.brute_force(x, y, z, state, x[:50], y[:50], z[:50], np.full(50, 9.0, np.float32), 64))
    with pytest.raises(Exception):
        index.FindNearestSurfelsWithinRadius(to_dev(x[:5]), to_dev(y[:5]), to_dev(z[:5]), to_dev(r2_rows[:5]), 65)
    index.close()


@pytest.mark.skipif(not octree_ref.available(), reason="oracle/_ref/liboctree_ref.so not built")
def test_knn_over_a_reconstruction(product):
    """The snapshot the meshing thread gets (TransferAllToCPU arrays) fed to the reference octree vs the index built
    straight from the handle on the device: every active surfel asks for its neighbours within its own radius, the
    query of TriangulateSurfel (surfel_meshing.cc:323,421)."""
    from surfelmeshing_b200 import reconstruction as R
    from surfelmeshing_b200 import synthetic as S
    from surfelmeshing_b200._lib import IntegrateParams, PreprocessParams
    from surfelmeshing_b200.knn import SurfelKnnIndex
    cam = S.Camera.tum(320, 240)
    st = S.make_stream(cam, 16, stream_id=3, device="cpu")
    st.depth, st.color = st.depth.cuda(), st.color.cuda()
    pp = PreprocessParams.defaults()
    pp.depth_valid_region_radius = cam.valid_region_radius()
    first, last = st.integrated_range()
    rec = R.CUDASurfelReconstruction(400_000, cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, lib=product)
    rec.stream_run(None, st.depth, st.color, st.global_T_frame, st.frame_T_global, st.others_TR_reference, pp,
                   IntegrateParams.defaults(), first, last)
    buffers = rec.TransferAllToCPU(None, last - 1)
    n = buffers["surfel_count"]
    x, y, z, r2_rows = [np.ascontiguousarray(buffers[key][:n]) for key in
                        ("surfel_x_buffer", "surfel_y_buffer", "surfel_z_buffer", "surfel_radius_squared_buffer")]
    active = np.nonzero(r2_rows > 0)[0]
    assert len(active) > 1000
    state = np.where(r2_rows > 0, 0, 255).astype(np.uint8)
    qi = active[:: max(1, len(active) // 2000)]
    r2 = (r2_rows[qi] * 4).astype(np.float32)
    tree = octree_ref.Octree(x, y, z, state)
    want = tree.query(x[qi], y[qi], z[qi], r2, 64)[:3]
    tree.close()
    index = SurfelKnnIndex(rec.max_surfel_count)
    assert index.build_from_reconstruction(rec, 2 * float(np.sqrt(np.median(r2)))) == n
    got = gpu_query(index, x[qi], y[qi], z[qi], r2, 64)
    assert_same_neighbours(got, want, 64)
    assert (got[1][:, 0] == qi).mean() > 0.99      # neighbour 0 is the surfel itself (surfel_meshing.cc:433-437)
    index.close()
    rec.close()


@pytest.mark.parametrize("kind,n", [("cube", 1000), ("sheet", 10000), ("cube", 10000)])
def test_gpu_batch_feeds_the_reference_triangulation(kind, n):
    """BASELINE config 1 with the GPU in the loop: the reference's own CPU meshing (surfel_meshing.cc, unmodified,
    oracle/_ref/libmeshing_ref.so) triangulates the cloud once asking its octree and once answering the octree
    queries of TriangulateSurfel / RemeshTrianglesAt from ONE sm_knn_query batch (every surfel, all meshing states,
    radius = the largest TriangulateSurfel can ask for); the meshes must agree triangle for triangle."""
    from oracle import meshing_ref
    from tests.test_meshing_oracle import batch_radius, run_reference
    if not meshing_ref.available():
        pytest.skip("oracle/_ref/libmeshing_ref.so not built")
    cloud = knn_cases.meshing_cloud(n, 6, kind)
    want = run_reference(cloud)
    r2 = batch_radius(cloud)
    index = build(cloud["x"], cloud["y"], cloud["z"], 2.0 * float(np.sqrt(r2.max())), radius_squared=cloud["radius_squared"])
    d2, idx, cnt = gpu_query(index, cloud["x"], cloud["y"], cloud["z"], r2, 64)
    index.close()
    got = run_reference(cloud, batch=(d2, idx, cnt, r2))
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    served, fallback = got[3]
    print(f"{kind} {n}: {len(want[0])} triangles, {served} octree queries answered from the GPU batch, {fallback} by the octree")
    assert served > 0 and (kind == "cube" and n == 10000 or served > fallback)


def test_knn_batch_host_matches_the_device_api():
    """sm_knn_batch_host (host arrays in and out, what a CPU meshing thread holds) returns exactly what build + query on
    device buffers return, merged points (radius^2 < 0) are neither indexed nor answered, and the reference's
    triangulation fed with it produces the octree's mesh."""
    from surfelmeshing_b200.knn import SurfelKnnIndex
    cloud = knn_cases.meshing_cloud(10000, 8, "sheet")
    r2_rows = cloud["radius_squared"].copy()
    r2_rows[::17] = -1.0                                   # merged surfels
    factor = np.float32(4.0)
    index = SurfelKnnIndex(20000)
    d2, idx, cnt = index.batch_host(cloud["x"], cloud["y"], cloud["z"], r2_rows, float(factor), 64)
    cell = 2.0 * float(np.sqrt(r2_rows.max() * factor))
    other = build(cloud["x"], cloud["y"], cloud["z"], cell, radius_squared=r2_rows)
    want = gpu_query(other, cloud["x"], cloud["y"], cloud["z"], (r2_rows * factor).astype(np.float32), 64)
    other.close()
    assert_identical((d2, idx, cnt), want)
    assert (cnt[::17] == 0).all() and not np.isin(idx[idx != 0xFFFFFFFF], np.arange(0, 10000, 17)).any()
    # a second, smaller batch reuses the staging
    d2b, idxb, cntb = index.batch_host(cloud["x"][:3000], cloud["y"][:3000], cloud["z"][:3000], r2_rows[:3000], float(factor), 16)
    assert d2b.shape == (3000, 16) and (cntb <= 16).all() and cntb.max() > 1
    index.close()
    from oracle import meshing_ref
    if meshing_ref.available():
        from tests.test_meshing_oracle import batch_radius, run_reference
        cloud["radius_squared"] = np.abs(r2_rows)          # the meshing input itself has no merged surfels here
        index = SurfelKnnIndex(10000)
        batch = index.batch_host(cloud["x"], cloud["y"], cloud["z"], cloud["radius_squared"], float(factor), 64)
        index.close()
        want_mesh = run_reference(cloud)
        got_mesh = run_reference(cloud, batch=(*batch, batch_radius(cloud)))
        assert np.array_equal(got_mesh[0], want_mesh[0]) and np.array_equal(got_mesh[1], want_mesh[1]) and got_mesh[3][0] > 0
