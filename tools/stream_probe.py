#!/usr/bin/env python
"""Runs one pass of a synthetic stream through the product or the reference oracle and
prints frames/s and the per-stage GPU timings (used under ncu for launch lists)."""
import argparse
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from surfelmeshing_b200 import _lib, synthetic as S  # noqa: E402
from surfelmeshing_b200 import reconstruction as R  # noqa: E402
from surfelmeshing_b200._lib import IntegrateParams, PreprocessParams  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="product")
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--cap", type=int, default=5_000_000)
    ap.add_argument("--gen", default="cuda")
    ap.add_argument("--host", action="store_true", help="frames in pinned host memory (e2e path)")
    ap.add_argument("--timings", action="store_true")
    ap.add_argument("--lib", default=None, help="path of a product library variant (A/B builds)")
    args = ap.parse_args()
    if args.lib:
        lib = _lib.Library(Path(args.lib).resolve(), "sm_", product=True)
    else:
        lib = _lib.load_product() if args.impl == "product" else _lib.load_reference_oracle()
    cam = S.Camera.tum(args.width, args.height)
    t0 = time.time()
    st = S.make_stream(cam, args.frames, device=args.gen)
    depth, color = st.depth, st.color
    if args.host:
        depth, color = depth.cpu().pin_memory(), color.cpu().pin_memory()
    else:
        depth, color = depth.cuda(), color.cuda()
    print(f"generated {args.frames} frames in {time.time() - t0:.1f}s", flush=True)
    pp = PreprocessParams.defaults()
    pp.depth_valid_region_radius = cam.valid_region_radius()
    ip = IntegrateParams.defaults()
    f0, f1 = st.integrated_range()
    rec = R.CUDASurfelReconstruction(args.cap, cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, lib=lib)
    for rep in range(args.reps):
        rec.reset()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        stats = rec.stream_run(None, depth, color, st.global_T_frame, st.frame_T_global, st.others_TR_reference, pp, ip,
                               f0, f1)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.time() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        print(f"{args.impl} rep {rep}: {stats.frames_integrated} frames, {ms:.2f} ms gpu / {wall:.2f} ms wall -> "
              f"{stats.frames_integrated / ms * 1e3:.1f} fps; surfels {stats.surfels_size} count {stats.surfel_count} "
              f"launches {stats.kernel_launches} h2d {stats.h2d_bytes} host_enqueue_ms {stats.host_enqueue_ms:.2f}", flush=True)
    if args.timings:
        rec.enable_timings(True)
        # time the last frames individually through the stream runner (1 frame per call)
        import numpy as np
        acc = np.zeros(7)
        n = 0
        rec.reset()
        rec.stream_run(None, depth, color, st.global_T_frame, st.frame_T_global, st.others_TR_reference, pp, ip, f0, f1 - 8)
        for fr in range(f1 - 8, f1):
            rec.stream_run(None, depth, color, st.global_T_frame, st.frame_T_global, st.others_TR_reference, pp, ip, fr, fr + 1)
            acc += np.array(rec.GetTimings())
            n += 1
        names = ["association", "merging", "blending", "integration", "neighbors", "creation", "regularization"]
        print("stage ms (mean of last 8 frames): " + ", ".join(f"{k} {v / n:.4f}" for k, v in zip(names, acc)))


if __name__ == "__main__":
    main()
