"""Point clouds and query sets for the neighbour-search tests (patterns of the reference's octree test)."""
import numpy as np


def random_cloud(n, seed, extent=10.0):
    """`extent * Vec3f::Random()` (test/test_octree.cc:450-460): uniform in [-extent, extent]^3."""
    rng = np.random.default_rng(seed)
    return [np.ascontiguousarray(a) for a in (rng.random((3, n), dtype=np.float32) * 2 * extent - extent).astype(np.float32)]


def nasty_cloud(base_count, seed):
    """Base points plus copies offset by +-0.1 along each axis (kOffsets, test/test_octree.cc:297-304): many equal
    distances and one exact duplicate per base point."""
    rng = np.random.default_rng(seed)
    base = (rng.random((base_count, 3), dtype=np.float32) * 20 - 10).astype(np.float32)
    offsets = np.array([[0, 0, 0], [0, 0, 0], [0.1, 0, 0], [-0.1, 0, 0], [0, 0.1, 0], [0, -0.1, 0], [0, 0, 0.1], [0, 0, -0.1]],
                       np.float32)
    pts = (base[:, None, :] + offsets[None, :, :]).reshape(-1, 3).astype(np.float32)
    return [np.ascontiguousarray(pts[:, a]) for a in range(3)]


def surface_cloud(n, seed, spacing=0.01):
    """Points near a wavy sheet with ~`spacing` between neighbours: what the meshing thread queries (a surfel
    cloud of a surface, query radius^2 = the surfel's radius^2, ~2 x spacing)."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(np.sqrt(n)))
    u, v = np.meshgrid(np.arange(side, dtype=np.float32), np.arange(side, dtype=np.float32))
    u, v = u.ravel()[:n], v.ravel()[:n]
    x = (u + rng.normal(0, 0.2, n)) * spacing
    y = (v + rng.normal(0, 0.2, n)) * spacing
    z = 1.0 + 0.05 * np.sin(3 * x) * np.cos(2 * y) + rng.normal(0, 0.1 * spacing, n)
    perm = rng.permutation(n)   # slot order is creation order, not spatial order
    return [np.ascontiguousarray(a[perm].astype(np.float32)) for a in (x, y, z)]


def states(n, seed):
    rng = np.random.default_rng(seed)
    s = rng.integers(0, 3, n).astype(np.uint8)
    s[rng.random(n) < 0.05] = 255
    return s


def meshing_cloud(n, seed, kind="sheet", thickness=0.02):
    """Input of one meshing iteration (the eight CUDASurfelBuffersCPU arrays). "cube": the reference's triangulation
    test (test/test_triangulation.cc:70-88: positions 0.5 * range * Random(), one radius, normal +x), the shape
    BASELINE config 1 scales to 10 k surfels; "sheet": surfels of a surface with radius 1.5 x their spacing, what a
    reconstruction hands over."""
    rng = np.random.default_rng(seed)
    if kind == "cube":
        x, y, z = [(rng.random(n, dtype=np.float32) - 0.5).astype(np.float32) for _ in range(3)]
        r2 = np.full(n, 0.1 * 0.1, np.float32)
    else:
        x = ((rng.random(n, dtype=np.float32) - np.float32(0.5)) * np.float32(thickness)).astype(np.float32)
        y = (rng.random(n, dtype=np.float32) - 0.5).astype(np.float32)
        z = (rng.random(n, dtype=np.float32) - 0.5).astype(np.float32)
        r2 = np.full(n, (1.5 / np.sqrt(n)) ** 2, np.float32)
    nx, ny, nz = np.ones(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    return dict(x=x, y=y, z=z, radius_squared=r2, nx=nx, ny=ny, nz=nz, stamp=np.ones(n, np.uint32))
