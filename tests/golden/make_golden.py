#!/usr/bin/env python
"""Generates tests/golden/golden_320x240.npz from the REFERENCE's own kernels.

Run on a B200 box (gpurun): the oracle library oracle/_ref/libsurfel_ref.so (the
reference's unmodified .cu files rebuilt for sm_100a + oracle/ref_driver.cu) processes a
small synthetic stream; inputs and the oracle's outputs of every stage are stored as the
project's known-answer set (the reference itself has no fixtures for this path).

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden_320x240.npz'
    cp gpurun_out/golden_320x240.npz tests/golden/
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from surfelmeshing_b200 import _lib, synthetic as S  # noqa: E402
from surfelmeshing_b200 import reconstruction as R  # noqa: E402
from surfelmeshing_b200._lib import IntegrateParams, PreprocessParams  # noqa: E402

W, H, FRAMES, CAP = 320, 240, 12, 400_000


def main(out_path):
    ref = _lib.load_reference_oracle()
    cam = S.Camera.tum(W, H)
    st = S.make_stream(cam, FRAMES, stream_id=7, device="cpu")
    depth, color = st.depth.cuda(), st.color.cuda()
    pp = PreprocessParams.defaults()
    pp.depth_valid_region_radius = cam.valid_region_radius()
    ip = IntegrateParams.defaults()
    K = pp.outlier_filtering_frame_count
    first, last = st.integrated_range()
    out = {
        "camera": np.array([W, H, cam.fx, cam.fy, cam.cx, cam.cy], dtype=np.float64),
        "depth": st.depth.numpy(), "color": st.color.numpy(),
        "global_T_frame": st.global_T_frame, "frame_T_global": st.frame_T_global,
        "others_TR_reference": st.others_TR_reference,
        "frames": np.array([first, last]), "cap": np.array([CAP]),
        "valid_region_radius": np.array([pp.depth_valid_region_radius], dtype=np.float32),
    }

    def u16():
        return torch.zeros((H, W), dtype=torch.uint16, device="cuda")

    rec = R.CUDASurfelReconstruction(CAP, W, H, cam.fx, cam.fy, cam.cx, cam.cy, lib=ref)
    for frame in range(first, last):
        others = [depth[frame - (i + 1)] for i in range(K // 2)] + [depth[frame + (i + 1)] for i in range(K // 2)]
        mats = st.others_TR_reference[frame]
        # the five stages one by one (reference host wrappers)
        A, B, A2, B2, A3 = u16(), u16(), u16(), u16(), u16()
        normals = torch.zeros((H, W, 2), dtype=torch.float32, device="cuda")
        radius = torch.zeros((H, W), dtype=torch.float32, device="cuda")
        R.BilateralFilteringAndDepthCutoffCUDA(None, pp.bilateral_filter_sigma_xy, pp.bilateral_filter_sigma_depth_factor,
                                               0, pp.bilateral_filter_radius_factor, int(pp.depth_scaling * pp.max_depth),
                                               pp.depth_valid_region_radius, depth[frame], A, lib=ref)
        R.OutlierDepthMapFusionCUDA(None, pp.outlier_filtering_depth_tolerance_factor, A, cam.fx, cam.fy, cam.cx, cam.cy,
                                    others, mats, B, lib=ref)
        R.ErodeDepthMapCUDA(None, pp.depth_erosion_radius, B, A2, lib=ref)
        R.ComputeNormalsAndDropBadPixelsCUDA(None, pp.observation_angle_threshold_deg, pp.depth_scaling, cam.fx, cam.fy,
                                             cam.cx, cam.cy, A2, B2, normals, lib=ref)
        R.ComputePointRadiiAndRemoveIsolatedPixelsCUDA(None, pp.point_radius_extension_factor,
                                                       pp.point_radius_clamp_factor, pp.depth_scaling, cam.fx, cam.fy,
                                                       cam.cx, cam.cy, B2, radius, A3, lib=ref)
        torch.cuda.synchronize()
        pre = f"f{frame}_"
        for k, v in (("bilateral", A), ("outlier", B), ("erode", A2), ("normals_depth", B2), ("normals", normals),
                     ("radius", radius), ("pre_depth", A3)):
            out[pre + k] = v.cpu().numpy()
        d = A3.clone()
        rec.integrate(None, frame, ip, d, normals, radius, color[frame], st.global_T_frame[frame],
                      st.frame_T_global[frame])
        torch.cuda.synchronize()
        out[pre + "blended_depth"] = d.cpu().numpy()
        for k, v in rec.download_rasters().items():
            out[pre + k] = v
        rows, n, merges = rec.dump_state()
        out[pre + "state"] = rows.copy()
        out[pre + "counts"] = np.array([n, merges])
        print(f"frame {frame}: surfels {n} merges {merges} valid px {int((A3 != 0).sum())}")
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, Path(out_path).stat().st_size, "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden_320x240.npz")
