#!/usr/bin/env python
"""Debug: find blended-depth mismatches product vs oracle for a given blend radius."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from surfelmeshing_b200 import _lib, synthetic as S, reconstruction as R
from surfelmeshing_b200._lib import IntegrateParams, PreprocessParams

radius = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ref = _lib.load_reference_oracle()
cam = S.Camera.tum(640, 480)
st = S.make_stream(cam, 13, stream_id=11, device="cuda")
pp, ip = PreprocessParams.defaults(), IntegrateParams.defaults()
ip.measurement_blending_radius = radius
W, H = 640, 480
rp = R.CUDASurfelReconstruction(600_000, W, H, cam.fx, cam.fy, cam.cx, cam.cy)
rr = R.CUDASurfelReconstruction(600_000, W, H, cam.fx, cam.fy, cam.cx, cam.cy, lib=ref)
rb = R.CUDASurfelReconstruction(600_000, W, H, cam.fx, cam.fy, cam.cx, cam.cy, lib=ref)
first, last = st.integrated_range()
for frame in range(first, last):
    others = [st.depth[frame - (i + 1)] for i in range(4)] + [st.depth[frame + (i + 1)] for i in range(4)]
    d0 = torch.zeros((H, W), dtype=torch.uint16, device="cuda"); n0 = torch.zeros((H, W, 2), device="cuda"); r0 = torch.zeros((H, W), device="cuda")
    rr.preprocess(None, pp, st.depth[frame], others, st.others_TR_reference[frame], d0, n0, r0)
    rows, nb, merges = rr.dump_state()
    rp.load_state(rows, merges); rb.load_state(rows, merges)
    ds = {}
    for k, rec in (("p", rp), ("r", rr), ("b", rb)):
        d = d0.clone()
        rec.integrate(None, frame, ip, d, n0, r0, st.color[frame], st.global_T_frame[frame], st.frame_T_global[frame])
        torch.cuda.synchronize()
        ds[k] = d.cpu().numpy().astype(np.int32)
    rasp, rasr = rp.download_rasters(), rr.download_rasters()
    orig = d0.cpu().numpy().astype(np.int32)
    for other in ("p", "b"):
        ys, xs = np.nonzero(ds[other] != ds["r"])
        print(f"frame {frame} {other} vs r: {len(ys)} mismatching px; changed px by blend: {(ds['r'] != orig).sum()}")
        for y, x in list(zip(ys, xs))[:3]:
            print("  at", y, x, "orig", orig[y, x], other, ds[other][y, x], "ref", ds["r"][y, x])
            sl = (slice(y - 2, y + 3), slice(x - 2, x + 3))
            print("  orig depth\n", orig[sl]); print("  mine\n", ds[other][sl]); print("  ref\n", ds["r"][sl])
            print("  counts\n", rasr["supporting_surfel_counts"][sl])
            print("  sums p\n", rasp["supporting_surfel_depth_sums"][sl].view(np.uint32)); print("  sums r\n", rasr["supporting_surfel_depth_sums"][sl].view(np.uint32))
            print("  sup\n", (rasr["supporting_surfels"][sl] != 0xFFFFFFFF).astype(int))
