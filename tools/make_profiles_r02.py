#!/usr/bin/env python
"""Assembles profiles/r02_bench.md (+ ncu summaries, launch lists) from what the round-2 GPU calls left in
gpurun_out/. Later calls override earlier ones (c5 > c4 > c3 > c2) wherever both hold the same item."""
import io
import json
import shutil
import subprocess
import sys
from contextlib import redirect_stdout
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
SRC, OUT = ROOT / "gpurun_out", ROOT / "profiles"
sys.path.insert(0, str(ROOT / "tools"))
import launch_list_summary  # noqa: E402


def load_line(path):
    try:
        return json.loads([ln for ln in path.read_text().splitlines() if ln.startswith("{")][-1])
    except (OSError, IndexError, ValueError):
        return None


def latest(pattern):
    for call in ("c20", "c19", "c18", "c17", "c16", "c15", "c14", "c13", "c12", "c11", "c10", "c9", "c8", "c7", "c6", "c5", "c4", "c3", "c2", "c1"):
        p = SRC / pattern.format(call=call)
        if p.exists() and p.stat().st_size > 0:
            return p
    return None


def launch_table(path):
    buf = io.StringIO()
    old = sys.argv
    sys.argv = ["launch_list_summary.py", str(path)]
    try:
        with redirect_stdout(buf):
            launch_list_summary.main()
    finally:
        sys.argv = old
    return buf.getvalue()


def bench_rows(tag, title, lines):
    p, r = latest("{call}_bench_product" + tag + ".json"), latest("{call}_bench_reference" + tag + ".json")
    pj, rj = (load_line(p) if p else None), (load_line(r) if r else None)
    if not pj or not rj:
        return None, None
    fp = pj["config"]["frames_per_step"]
    lines.append(f"| {title} | {pj['value']:.0f} | {pj['e2e']['value']:.0f} | {rj['value']:.0f} | {rj['e2e']['value']:.0f} | "
                 f"{pj['value'] / rj['value']:.2f}x | {pj['e2e']['value'] / rj['e2e']['value']:.2f}x | "
                 f"{pj['config']['surfels_after_step']} / {rj['config']['surfels_after_step']} | "
                 f"{pj['gpu_launches'] / pj['steps'] / fp:.1f} / {rj['gpu_launches'] / rj['steps'] / fp:.1f} |")
    return pj, rj


def kernel_table(pj, lines):
    rf = pj.get("roofline")
    if not rf:
        return
    c = rf["counters"]
    lines.append(f"Counters of the last frame: P = {c['P']}, N = {c['N']}, V = {c['V']}, S = {c['S']}, M = {c['M']}, A = {c['A']}, "
                 f"D = {c.get('D')}. Roofline region: {rf.get('region', 'whole step')}; frame period inside the frame graph "
                 f"{rf.get('pipelined_frame_period_us', float('nan')):.1f} us.")
    lines.append(f"`roofline`: kernel `{rf['kernel']}` (longest on the binding dependency cycle `{rf.get('binding_cycle')}`), "
                 f"{rf['achieved']:.0f} GB/s on {rf['algorithmic_bytes_per_launch'] / 1e6:.1f} MB algorithmic bytes = "
                 f"{100 * rf['frac']:.1f} % of the measured HBM peak ({rf['peak']:.0f} GB/s); ncu DRAM traffic per launch: {rf['traffic']}.\n")
    lines.append("| kernel | launches (region) | events us (serial pass) | share | in-graph us (device timeline) | GB/s on algorithmic bytes (events) |")
    lines.append("|---|---:|---:|---:|---:|---:|")
    for k, v in pj["kernels"].items():
        if "mean_us" not in v:
            continue
        gbs = f"{v['achieved_gbs']:.0f} ({100 * v['frac_of_hbm_peak']:.1f} %)" if "achieved_gbs" in v else ""
        lines.append(f"| {k} | {v['launches']} | {v['mean_us']:.2f} | {100 * v['share']:.1f} % | {v.get('pipelined_us', float('nan')):.2f} | {gbs} |")
    lines.append("")


def main():
    OUT.mkdir(exist_ok=True)
    L = ["# r02: benchmark lines, A/B runs and ncu evidence (B200)\n",
         "All numbers from `gpurun` calls of this round (fresh B200 box each, SM clock 1965 MHz, no throttle reason in any line). "
         "`value` = frames resident in HBM, `e2e` = pinned host frames in (upload stream) + `TransferAllToCPU` out. The reference arm "
         "is the reference's own kernels rebuilt for sm_100a (`oracle/_ref`).\n",
         "## Bench lines\n",
         "| configuration | product value | product e2e | reference value | reference e2e | ratio | e2e ratio | surfels after a step (product / reference) | launches per frame |",
         "|---|---:|---:|---:|---:|---:|---:|---:|---:|"]
    c2p, _ = bench_rows("", "C2: VGA, 500 frames, 5 M cap (the headline)", L)
    c3p, _ = bench_rows("_C3", "C3: 1280x960, 1000 frames, 20 M cap", L)
    bench_rows("_C5", "C5: VGA, sigma 0.05 m, 2000 frames (the 2 % outlier test removes every pixel: empty cloud in both arms)", L)
    bench_rows("_C5_req5", "C5, required inliers 5 of 8 (still empty: erosion needs a full 5x5 window)", L)
    bench_rows("_C5b_500", "C5b: sigma 0.05 m, required inliers 1, erosion 0, 40 M cap, 500 frames: populated cloud under heavy noise", L)
    bench_rows("_C5b_2000", "C5b, 2000 frames (with the default 5 M cap both arms overflow: the noisy stream creates ~3 300 surfels per frame)", L)
    n2p, n2r = load_line(SRC / "c11_bench_product_n2.json"), load_line(SRC / "c11_bench_reference_n2.json")
    if n2p and n2r:
        L.append(f"\nReplicas (`bench.py --gpus 2` under torchrun as the driver launches it, one independent C2 stream per GPU, "
                 f"max-over-ranks timing): product {n2p['value']:.0f} frames/s (e2e {n2p['e2e']['value']:.0f}), reference "
                 f"{n2r['value']:.0f} (e2e {n2r['e2e']['value']:.0f}); round 1's driver run measured 0.98 scaling efficiency up to 8 GPUs "
                 f"with the same plumbing (no collective on the data path).")
    n4p, n4r = load_line(SRC / "c23_bench_product_n4.json"), load_line(SRC / "c23_bench_reference_n4.json")
    if n4p and n4r:
        L.append(f"\nThe same at 4 GPUs (GPU call 23, a box whose host was busy): product {n4p['value']:.0f} frames/s ({n4p['ms_per_step']:.1f} ms per step "
                 f"on the slowest rank against 39.2 ms alone), reference {n4r['value']:.0f} ({n4r['ms_per_step']:.0f} ms per step against 180 ms alone: it blocks "
                 f"the host twice per frame).")
    diag = [(n, load_line(SRC / f"c24_{n}.json")) for n in ("graph", "graph_noclocks", "streams", "graph_again")]
    if all(d for _, d in diag):
        L.append("\nGPU call 24 (4 GPUs, 128 host cores, load average 17 from other tenants), per-rank step times from `ms_per_step_per_rank`:\n")
        L.append("| run | frames/s | ms per step, rank 0..3 | replica efficiency vs 12 546 alone |")
        L.append("|---|---:|---|---:|")
        names = {"graph": "frame graph (default)", "graph_noclocks": "frame graph, nvidia-smi sampler off (diagnosis)",
                 "streams": "round 1's stream pipeline (`SM_B200_GRAPH=0`)", "graph_again": "frame graph again"}
        for n, d in diag:
            L.append(f"| {names[n]} | {d['value']:.0f} | {', '.join(f'{v:.1f}' for v in d['ms_per_step_per_rank'])} | {d['value'] / (4 * 12546):.2f} |")
        L.append("\nEvery rank is ~7 % slower than alone (no straggler, the sampler is not the cause); the stream pipeline, whose host side "
                 "is four times as expensive per frame, loses a third on the same box. Round 1's driver run measured 0.98 at 8 GPUs on an idle box.")
    L.append("")
    if c2p:
        L.append("### C2 kernel table\n")
        kernel_table(c2p, L)
        cb = c2p.get("cpu_baseline")
        if cb:
            L.append(f"`cpu_baseline`: {cb['value']:.1f} frames/s on {cb['cores']} host cores ({cb['sample']}); runs {cb.get('runs')}, "
                     f"spread {cb.get('spread')}.\n")
    if c3p:
        L.append("### C3 kernel table (the HBM-bound regime)\n")
        kernel_table(c3p, L)
    for call, title in (("c1", "call 1 (graph with programmatic edges + split projection as default)"),
                        ("c2", "call 2 (graph, plain edges, one projection launch as default)"),
                        ("c4", "call 4 (scheduling hooks)"), ("c7", "call 7 (final kernels: light-first-batch gathers, multiply-based key decode)"),
                        ("c8", "call 8 (k_update_neighbors with block-level survivor compaction, -DSM_UPDATE_COMPACT=1: slower at VGA, kept off)"),
                        ("c8_hd", "call 8, 1280x960 / 400 frames / 20 M cap (3 passes)"),
                        ("c8c", "call 8c (ordering-only graph edges, SM_B200_GRAPH_ORDER: 1 create before update_neighbors, 2 associate(f+1) "
                                "before the regularisation of f, 4 merge(f+1) before it, 8 project(f+1) before update_neighbors(f): all slower, kept off)"),
                        ("c8c_hd", "call 8c, 1280x960 / 400 frames / 20 M cap (3 passes)"),
                        ("c22", "call 22 (k_update_neighbors: the 'some candidate is new' gate before the other gates' gathers, "
                                "-DSM_UPDATE_ANYNEW_FIRST=1; `head` = the default build as a variant library; later configurations of one "
                                "process run up to 5 % slower than the first ones, so each pair is listed twice)"),
                        ("c22_hd", "call 22, 1280x960 / 400 frames / 20 M cap (3 passes)")):
        p = SRC / (f"{call}_ab.json" if not call.endswith("_hd") else f"{call[:-3]}_ab_hd.json")
        if p.exists():
            L.append(f"## Same-box A/B, {title}\n")
            L.append("`tools/ab_probe.py`: one process, one stream, 5 timed passes per configuration (+ warm-up), CUDA events.\n")
            L.append("| configuration | best frames/s | median | host enqueue ms / pass | surfels | launches / pass |")
            L.append("|---|---:|---:|---:|---:|---:|")
            for e in json.loads(p.read_text()):
                knobs = ", ".join(f"{k}={v}" for k, v in e["env"].items()) + (f", lib={e['lib']}" if e["lib"] != "product" else "")
                L.append(f"| {e['config']} ({knobs or 'defaults'}) | {e['fps_best']:.0f} | {e['fps_median']:.0f} | {e['host_enqueue_ms']:.2f} | "
                         f"{e['surfels_size']} / {e['surfel_count']} | {e['launches']} |")
            L.append("")
    L.append("## f1: delta `TransferAllToCPU` (`tools/transfer_probe.py`: stream_run in chunks, a hand-off to pageable arrays after each)\n")
    L.append("| stream | transfer every | mode | frames/s end to end | ms inside the transfer calls | D2H bytes | transfers |")
    L.append("|---|---:|---|---:|---:|---:|---:|")
    for pattern, stream, every in (("{call}_transfer_probe.json", "C2 (VGA, 0.54 M surfels)", 30),
                                   ("{call}_transfer_probe_every5.json", "C2 (VGA, 0.54 M surfels)", 5),
                                   ("{call}_transfer_probe_hd.json", "1280x960, 400 frames (2.05 M surfels)", 10)):
        tp = latest(pattern)
        if not tp:
            continue
        j = json.loads(tp.read_text())
        for mode in ("full", "delta"):
            v = j[mode]
            L.append(f"| {stream} | {every} | {mode} | {v['frames_per_s']:.0f} | {v['transfer_ms']:.1f} | {v['d2h_bytes']} | {v['transfers']} |")
    L.append("\nThe delta moves 40 - 50 % fewer bytes (a third of the slots are inside the 30-frame regularisation window at any "
             "time, so they did change) but is not faster: its host-side scatter of the records into the CUDASurfelBuffersCPU "
             "arrays (3.3 - 4.0 ms for 0.68 M records on four threads) costs more than the 16 MB it saves on a 25 GB/s link. It "
             "pays when the consumer uses the changed-slot list instead of its O(N) comparison "
             "(`surfel_meshing.cc:199-250`), or over a slower link.\n")
    L.append("## f4: radius k-NN for the meshing thread (`tools/knn_probe.py`)\n")
    L.append("Every point queries its neighbours, k <= 64; the "
             "reference arm is the reference's own octree (`oracle/_ref/liboctree_ref.so`) on one host thread, the way the "
             "meshing thread calls it; every sampled query is compared with the GPU answer before a number is printed.\n")
    L.append("| cloud | points = queries | query radius, cell size | neighbours found (mean) | build ms | query ms | queries/s resident | "
             "queries/s end to end (H2D of points + queries, D2H of results) | reference octree queries/s (sample) | end-to-end ratio |")
    L.append("|---|---:|---|---:|---:|---:|---:|---:|---:|---:|")
    for pattern, label in (("c9_knn_probe.json", "surfel sheet, 5 mm spacing"), ("c8b_knn_probe_cell10.json", "same, cells of one radius"),
                           ("c8b_knn_probe_r5.json", "same, radius 25 mm (the 64-cap binds for 99 % of the queries)"),
                           ("c9_knn_probe_random1m.json", "uniform in a cube"),
                           ("c9_knn_probe_c1.json", "uniform in a cube, BASELINE config 1 size")):
        kp = SRC / pattern
        if not kp.exists():
            continue
        j = json.loads(kp.read_text())
        ref = j.get("reference_octree", {})
        w = j["workload"]
        pts = w.split(" ")[0]
        geometry = w[w.index("radius"):]
        L.append(f"| {label} | {pts} | {geometry} | {j['mean_neighbours_found']:.1f} | {j['build_ms']:.3f} | {j['query_ms']:.3f} | "
                 f"{j['queries_per_s_resident'] / 1e6:.0f} M | {j['queries_per_s_e2e'] / 1e6:.1f} M | "
                 f"{ref.get('queries_per_s', 0) / 1e3:.0f} k ({ref.get('sample', '')}) | {j.get('speedup_e2e_vs_reference_octree', 0):.0f}x |")
    L.append("\nThe first version of the query kernel inserted every candidate into the sorted list one by one (call 6: 1 M queries "
             "in 1.99 ms with cells of 2 r, 3.33 ms with cells of r); staging the candidates in shared memory and sorting once "
             "(call 8b / 9) brought that to 1.21 ms / 2.68 ms. `profiles/r02_knn_ncu_full_summary.csv` is the `ncu --set full` capture "
             "of the final kernels.\n")
    for pattern, title in (("{call}_meshing_probe_100k.json", "100 000"), ("{call}_meshing_probe_1m.json", "1 000 000")):
        mp = latest(pattern)
        if not mp:
            continue
        j = json.loads(mp.read_text())
        if "### One meshing iteration" not in "\n".join(L):
            L.append("### One meshing iteration of the reference's CPU code (BASELINE config 1 at scale, `tools/meshing_probe.py`)\n")
            L.append("`IntegrateCUDABuffers -> CheckRemeshing -> Triangulate` of `oracle/_ref/libmeshing_ref.so` (the reference's `surfel_meshing.cc` + "
                     "`octree.cc`, unmodified) over N fresh surfels, its two octree queries answered by the octree or by one `sm_knn_query` batch.\n")
            L.append("| surfels | triangles | identical mesh | octree: integrate + check + triangulate (s) | GPU batch: same (s) | batch end to end on the GPU, "
                     "pinned arrays (s) | batch through `sm_knn_batch_host`, pageable arrays (s) | queries from the batch / left to the octree | "
                     "iteration speed-up (pinned / host entry point) |")
            L.append("|---:|---:|---|---:|---:|---:|---:|---:|---:|")
        o, g = j["octree"], j["gpu_batch"]
        L.append(f"| {title} | {j['triangles']} | {j['identical_mesh']} | {o['integrate_s']:.3f} + {o['check_remeshing_s']:.3f} + {o['triangulate_s']:.3f} = {o['total_s']:.3f} | "
                 f"{g['integrate_s']:.3f} + {g['check_remeshing_s']:.3f} + {g['triangulate_s']:.3f} = {g['total_s']:.3f} | {j['gpu_batch_end_to_end_s']:.3f} | "
                 f"{j.get('gpu_batch_host_api_s', float('nan')):.3f} | "
                 f"{j['queries_answered_from_the_batch']} / {j['queries_left_to_the_octree']} | {j['iteration_speedup']:.2f}x / "
                 f"{j.get('iteration_speedup_host_api', float('nan')):.2f}x |")
    L.append("")
    for arm, note in (("product", "`-k regex:k_`, frames ~450-480 of one pass of `tools/stream_probe.py --frames 500`"),
                      ("reference", "`-k regex:Kernel`, the same frames of `--impl reference`")):
        f = latest("{call}_launches_" + arm + ".csv")
        if f:
            shutil.copy(f, OUT / f"r02_launches_{arm}.csv")
            L.append(f"## ncu launch list, {arm} ({note}; `--metrics gpu__time_duration.sum --clock-control none`: serialised, cold "
                     f"caches - the SHARE of a kernel is what carries over, not the absolute)\n\nRaw list: `profiles/r02_launches_{arm}.csv`.\n")
            L.append(launch_table(f))
    for rep, name, title in (("frame460_C2", "r02_frame460_C2_ncu_full", "C2 stream, frame 460 (N ~ 0.5 M: L2-resident working set)"),
                             ("frame960_C3", "r02_frame960_C3_ncu_full", "C3 stream, frame 960 (N ~ 3.5 M: HBM-bound regime)")):
        f = latest("{call}_" + rep + ".ncu-rep")
        if f:
            res = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(f), str(OUT / name)],
                                 capture_output=True, text=True)
            L.append(f"## `ncu --set full`, one launch per kernel: {title}\n\n`profiles/{name}_summary.csv`, `..._traffic.json`.\n\n"
                     f"```\n{res.stdout}```\n")
    (OUT / "r02_bench.md").write_text("\n".join(L) + "\n")
    print(OUT / "r02_bench.md")


if __name__ == "__main__":
    main()
