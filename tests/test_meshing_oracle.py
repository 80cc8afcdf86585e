"""CPU: BASELINE config 1 — random surfels through the reference's own CPU meshing (oracle/_ref/libmeshing_ref.so =
surfel_meshing.cc + octree.cc compiled unmodified): IntegrateCUDABuffers -> CheckRemeshing -> Triangulate, the pattern
of test/test_triangulation.cc. Pins the build (Eigen / libvis stand-ins) against committed answers and checks the
rule by which a batch of neighbour-search answers may replace the octree queries (here the batch comes from the
brute-force restatement; tests/test_knn_gpu.py feeds the GPU's)."""
import hashlib

import numpy as np
import pytest

from oracle import meshing_ref, octree_ref
from tests import knn_cases

pytestmark = pytest.mark.skipif(not (meshing_ref.available() and octree_ref.available()), reason="oracle/_ref not built")

# (kind, surfels) -> (triangles, sha256 of the uint32 triangle array) after Triangulate and after the remesh round:
# answers of the reference code built here, written by `PYTHONPATH=. python tests/test_meshing_oracle.py`
GOLDEN = {
    ("cube", 1000): (1569, "af1a07122067219bd2f6843bfb8d91576db098a52de7a8b0a5cb648e0df9eb55", 1616, "a451b9310be404a153e2d9bb348c3b2a0598c424dc2bb2a155fe282ca817c521"),
    ("sheet", 10000): (18306, "242ecf8089c31d777376e4e8a728eb5294394dca0301f2c2e40582bee68b2bb7", 18341, "15cd493579e22d29a1960b9361798bc9df7835025a7ba637b983de6620a2eb7a"),
    ("cube", 10000): (21417, "34371d1be198e111c2b6c33afe55f2f037b686f1cc16d3756f2e84519ab2a275", 21576, "1c7d72c29887c4841febd8d6b4fc4a68b4f8446968868faa9497cb26d715ba0b"),
}


def run_reference(cloud, batch=None):
    m = meshing_ref.SurfelMeshing()
    m.integrate(1, **cloud)
    if batch is not None:
        m.set_knn_batch(*batch)
    m.check_remeshing()
    m.triangulate()
    tri, states, stats = m.triangles(), m.meshing_states(), m.query_stats()
    # second half of the reference's test: remesh around the first ten surfels, triangulate again
    for i in range(10):
        m.remesh_at(i, 4.0)
    m.triangulate()
    tri2 = m.triangles()
    m.close()
    return tri, tri2, states, stats


def digest(tri):
    return hashlib.sha256(np.ascontiguousarray(tri, np.uint32).tobytes()).hexdigest()


def batch_radius(cloud):
    f = meshing_ref.SurfelMeshing.MAX_NEIGHBOR_SEARCH_RANGE_INCREASE_FACTOR
    return (cloud["radius_squared"] * np.float32(f * f)).astype(np.float32)


@pytest.mark.parametrize("kind,n", list(GOLDEN))
def test_reference_meshing_config1(kind, n):
    cloud = knn_cases.meshing_cloud(n, 5, kind)
    tri, tri2, states, stats = run_reference(cloud)
    assert len(tri) > 0 and stats[0] == 0 and stats[1] >= n          # every surfel asked the octree at least once
    assert tri.max() < n and (tri[:, 0] != tri[:, 1]).all()
    again = run_reference(cloud)
    assert np.array_equal(tri, again[0]) and np.array_equal(tri2, again[1])       # the CPU code is deterministic
    want = GOLDEN[(kind, n)]
    assert want is not None and (len(tri), digest(tri), len(tri2), digest(tri2)) == want


@pytest.mark.parametrize("kind,n", [("cube", 1000), ("sheet", 3000)])
def test_batch_answers_give_the_same_mesh(kind, n):
    """The rule of oracle/meshing_driver.cc (filter the all-states batch row by the CURRENT state and octree membership,
    fall back to the octree when the row was cut at 64) reproduces the reference's mesh triangle for triangle."""
    cloud = knn_cases.meshing_cloud(n, 6, kind)
    want = run_reference(cloud)
    r2 = batch_radius(cloud)
    d2, idx, cnt = octree_ref.brute_force(cloud["x"], cloud["y"], cloud["z"], None, cloud["x"], cloud["y"], cloud["z"], r2, 64)
    got = run_reference(cloud, batch=(d2, idx, cnt, r2))
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    served, fallback = got[3]
    assert served > n // 2, (served, fallback)


if __name__ == "__main__":
    for (kind, n) in GOLDEN:
        tri, tri2, _, _ = run_reference(knn_cases.meshing_cloud(n, 5, kind))
        print(f'    ("{kind}", {n}): ({len(tri)}, "{digest(tri)}", {len(tri2)}, "{digest(tri2)}"),')
