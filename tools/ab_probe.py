#!/usr/bin/env python
"""Same-box A/B of pipeline / kernel variants (run under gpurun): one synthetic stream, one process,
every configuration = environment knobs (read per call / per handle) and optionally another build of
the product library (--lib name=path). Prints frames/s (best and median of --reps passes, CUDA events)
and the surfel counts; writes gpurun_out/ab_probe.json."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from surfelmeshing_b200 import _lib, synthetic as S  # noqa: E402
from surfelmeshing_b200 import reconstruction as R  # noqa: E402
from surfelmeshing_b200._lib import IntegrateParams, PreprocessParams  # noqa: E402

KNOBS = ("SM_B200_GRAPH", "SM_B200_GRAPH_PDL", "SM_B200_SPLIT_PROJECT", "SM_B200_TAIL_FILL", "SM_B200_PDL",
         "SM_B200_CARVEOUT", "SM_B200_GRAPH_PRIO", "SM_B200_OFFCHAIN_GRID_PERCENT", "SM_B200_GRID_PERCENT", "SM_B200_TIEBREAK")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=500)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--cap", type=int, default=5_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--host", action="store_true", help="pinned host frames (e2e path)")
    ap.add_argument("--lib", action="append", default=[], help="name=path of another product build")
    ap.add_argument("--config", action="append", default=[],
                    help="name:KEY=VAL+KEY=VAL[+lib=name]; 'default' is always run first and last")
    ap.add_argument("--out", default="gpurun_out/ab_probe.json")
    args = ap.parse_args()

    libs = {"product": _lib.load_product()}
    for item in args.lib:
        name, path = item.split("=", 1)
        libs[name] = _lib.Library(Path(path).resolve(), "sm_", product=True)
    configs = [("default", {}, "product")]
    for item in args.config:
        name, _, rest = item.partition(":")
        env, lib = {}, "product"
        for kv in filter(None, rest.split("+")):
            k, v = kv.split("=", 1)
            if k == "lib":
                lib = v
            else:
                env[k] = v
        configs.append((name, env, lib))
    configs.append(("default_again", {}, "product"))

    cam = S.Camera.tum(args.width, args.height)
    st = S.make_stream(cam, args.frames, device="cuda")
    depth, color = st.depth, st.color
    if args.host:
        depth, color = depth.cpu().pin_memory(), color.cpu().pin_memory()
    pp = PreprocessParams.defaults()
    pp.depth_valid_region_radius = cam.valid_region_radius()
    ip = IntegrateParams.defaults()
    f0, f1 = st.integrated_range()
    torch.cuda.synchronize()
    results = []
    for name, env, lib in configs:
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        rec = R.CUDASurfelReconstruction(args.cap, cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, lib=libs[lib])
        rates, host_ms = [], []
        stats = None
        for rep in range(args.reps + 1):  # first pass = warm-up (graph instantiation, buffers)
            rec.reset()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            stats = rec.stream_run(None, depth, color, st.global_T_frame, st.frame_T_global, st.others_TR_reference, pp,
                                   ip, f0, f1)
            e1.record()
            torch.cuda.synchronize()
            if rep:
                rates.append(stats.frames_integrated / e0.elapsed_time(e1) * 1e3)
                host_ms.append(stats.host_enqueue_ms)
        rec.close()
        entry = {"config": name, "env": env, "lib": lib, "fps_best": max(rates), "fps_median": float(np.median(rates)),
                 "host_enqueue_ms": float(np.median(host_ms)), "surfels_size": int(stats.surfels_size),
                 "surfel_count": int(stats.surfel_count), "launches": int(stats.kernel_launches)}
        results.append(entry)
        print(f"{name:28s} best {entry['fps_best']:9.1f}  median {entry['fps_median']:9.1f} fps   host enqueue "
              f"{entry['host_enqueue_ms']:7.2f} ms   surfels {entry['surfels_size']} / {entry['surfel_count']}  "
              f"launches {entry['launches']}", flush=True)
    Path(args.out).parent.mkdir(exist_ok=True)
    Path(args.out).write_text(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
