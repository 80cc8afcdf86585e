"""CPU-side tests: synthetic stream generator, pose helpers, replica sharding over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from surfelmeshing_b200 import distributed as D
from surfelmeshing_b200 import synthetic as S
from surfelmeshing_b200.reconstruction import invert_rigid


def test_invert_rigid_roundtrip():
    poses = S.trajectory(5, stream_id=1)
    for p in poses:
        inv = invert_rigid(p)
        m = np.vstack([p, [0, 0, 0, 1]]) @ np.vstack([inv.astype(np.float64), [0, 0, 0, 1]])
        assert np.allclose(m, np.eye(4), atol=1e-6)


def test_stream_is_deterministic_and_tum_shaped():
    cam = S.Camera.tum(160, 120)
    a = S.make_stream(cam, 10, stream_id=4)
    b = S.make_stream(cam, 10, stream_id=4)
    c = S.make_stream(cam, 10, stream_id=5)
    assert a.depth.dtype == torch.uint16 and tuple(a.depth.shape) == (10, 120, 160)
    assert a.color.dtype == torch.uint8 and tuple(a.color.shape) == (10, 120, 160, 3)
    assert torch.equal(a.depth.view(torch.int16), b.depth.view(torch.int16)) and torch.equal(a.color, b.color)
    assert not torch.equal(a.depth.view(torch.int16), c.depth.view(torch.int16))
    d = a.depth.numpy().astype(np.float64) / 5000.0
    assert 0.3 < d[d > 0].min() and d.max() < 13.2 and (d > 0).mean() > 0.95
    assert a.integrated_range() == (4, 6)
    assert a.global_T_frame.dtype == np.float32 and a.others_TR_reference.shape == (10, 8, 3, 4)


def test_others_TR_reference_follows_main_cc():
    """others_TR_reference[f][k] maps a point of frame f (in raw u16 units) into other frame k
    (APP/main.cc:1039-1058): a point seen in both frames reprojects onto the same world point."""
    poses = S.trajectory(10, stream_id=0)
    m = S.others_TR_reference(poses, 5000.0, 8)
    frame, k, other = 5, 0, 4  # k = 0 -> frame - 1
    p_ref = np.array([0.1, -0.2, 1.5]) * 5000.0
    q = m[frame, k].astype(np.float64) @ np.append(p_ref, 1.0)
    world = poses[frame] @ np.append(p_ref / 5000.0, 1.0)
    expect = S.invert_poses(poses[other:other + 1])[0] @ np.append(world, 1.0) * 5000.0
    assert np.allclose(q, expect, rtol=1e-5, atol=0.05)
    assert np.allclose(m[0, 0], np.eye(4)[:3]), "frames without K/2 neighbours carry identity"


def test_high_noise_stream_is_noisier():
    cam = S.Camera.tum(160, 120)
    clean = S.make_stream(cam, 2, stream_id=0, sigma_depth=0.0).depth.numpy().astype(np.int32)
    noisy = S.make_stream(cam, 2, stream_id=0, sigma_depth=0.05).depth.numpy().astype(np.int32)
    both = (clean > 0) & (noisy > 0)
    assert 0.03 * 5000 < np.std((noisy - clean)[both]) < 0.07 * 5000


def test_stream_assignment_covers_every_stream_once():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            seen += D.stream_ids_for_rank(D.RankInfo(r, r, world), 8)
        assert sorted(seen) == list(range(8))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    info = D.rank_info_from_env()
    D.init_process_group(info, backend="gloo")
    D.barrier(info)
    # rank r "processed" 100 + r frames in 10 * (r + 1) ms
    t, u = D.aggregate(info, 10.0 * (rank + 1), 100.0 + rank)
    out.put((rank, t, u, D.stream_ids_for_rank(info, 8)))
    torch.distributed.destroy_process_group()


def test_replica_aggregation_over_gloo_world_size_2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, t, u, streams in res:
        assert t == 20.0 and u == 201.0, "max time over ranks, total frames over ranks"
        assert streams == [s for s in range(8) if s % 2 == rank]
