"""CPU: pins the neighbour-search oracle. The reference's own octree (oracle/_ref/liboctree_ref.so) must agree with
the restatement of its test's brute-force checker (test/test_octree.cc:116-149) and with the committed answers in
tests/golden/knn_golden.npz (made by tests/golden/make_knn_golden.py from that octree)."""
import numpy as np
import pytest

from oracle import octree_ref
from tests import knn_cases

pytestmark = pytest.mark.skipif(not octree_ref.available(), reason="oracle/_ref/liboctree_ref.so not built")

GOLDEN_CASES = (("random", lambda: knn_cases.random_cloud(4000, 11), 1.5 ** 2, 16),
                ("nasty", lambda: knn_cases.nasty_cloud(300, 12), 0.35 ** 2, 6),
                ("surface", lambda: knn_cases.surface_cloud(6000, 13), 0.07 ** 2, 64))
FILTERS = (("all", 1, 1), ("triangulate", 0, 1), ("remesh", 1, 0))


def assert_same_neighbours(got, want, k):
    """Counts and squared distances bit for bit; indices exact once runs of equal distance are ordered by index,
    except inside a run that the cap k cuts (the reference keeps whichever member its traversal met last)."""
    d2_g, idx_g, cnt_g = got
    d2_w, idx_w, cnt_w = want
    assert np.array_equal(cnt_g, cnt_w)
    assert np.array_equal(d2_g.view(np.uint32), d2_w.view(np.uint32))
    d2_g, idx_g = octree_ref.canonical_ties(d2_g, idx_g, cnt_g)
    d2_w, idx_w = octree_ref.canonical_ties(d2_w, idx_w, cnt_w)
    for j in range(len(cnt_g)):
        c = int(cnt_g[j])
        settled = np.ones(c, bool) if c < k else d2_g[j, :c] < d2_g[j, c - 1]
        assert np.array_equal(idx_g[j, :c][settled], idx_w[j, :c][settled]), j


@pytest.mark.parametrize("name,cloud,r2_value,k", GOLDEN_CASES, ids=[c[0] for c in GOLDEN_CASES])
def test_reference_octree_matches_brute_force_and_golden(name, cloud, r2_value, k):
    golden = np.load(octree_ref.LIB_PATH.parents[2] / "tests" / "golden" / "knn_golden.npz")
    x, y, z = cloud()
    state = knn_cases.states(len(x), 14)
    qi = golden[f"{name}_query_index"]
    r2 = np.full(len(qi), r2_value, np.float32)
    for label, ic, fr in FILTERS:
        tree = octree_ref.Octree(x, y, z, state)
        d2, idx, cnt, _ = tree.query(x[qi], y[qi], z[qi], r2, k, ic, fr)
        tree.close()
        assert np.array_equal(cnt, golden[f"{name}_{label}_cnt"])
        assert np.array_equal(d2.view(np.uint32), golden[f"{name}_{label}_d2"].view(np.uint32))
        assert np.array_equal(idx, golden[f"{name}_{label}_idx"])
        assert_same_neighbours((d2, idx, cnt), octree_ref.brute_force(x, y, z, state, x[qi], y[qi], z[qi], r2, k, ic, fr), k)
        assert cnt.max() > 1 and ((cnt == k).any() or name == "random")


def test_reference_octree_edge_cases():
    x, y, z = knn_cases.random_cloud(500, 3)
    tree = octree_ref.Octree(x, y, z)
    far = np.array([100.0], np.float32)
    one = np.array([1.0], np.float32)
    d2, idx, cnt, _ = tree.query(far, far, far, one, 8)
    assert cnt[0] == 0 and np.isinf(d2).all() and (idx == 0xFFFFFFFF).all()
    d2, idx, cnt, _ = tree.query(x[:1], y[:1], z[:1], np.array([0.0], np.float32), 8)   # radius 0: the point itself
    assert cnt[0] == 1 and idx[0, 0] == 0 and d2[0, 0] == 0
    d2, idx, cnt, _ = tree.query(x[:1], y[:1], z[:1], np.array([-1.0], np.float32), 8)
    assert cnt[0] == 0
    tree.close()
