#!/usr/bin/env python
"""Assembles profiles/rNN_bench.md from the files a GPU run left in gpurun_out/:
bench JSON lines of both arms (N = 1 and, if present, N = 2), the two ncu launch lists and the
`ncu --set full` capture (through tools/ncu_summary.py). Usage: tools/make_profiles.py r01"""
import io
import json
import shutil
import subprocess
import sys
from contextlib import redirect_stdout
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
import launch_list_summary  # noqa: E402


def table(path):
    buf = io.StringIO()
    old = sys.argv
    sys.argv = ["launch_list_summary.py", str(path)]
    try:
        with redirect_stdout(buf):
            launch_list_summary.main()
    finally:
        sys.argv = old
    return buf.getvalue()


def load_line(path):
    """The JSON line of a bench run (libraries may print banners around it)."""
    return json.loads([ln for ln in path.read_text().splitlines() if ln.startswith("{")][-1])


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    out, src = ROOT / "profiles", ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    lines = [f"# {tag}: benchmark lines and ncu evidence (B200)\n"]
    p1 = load_line(src / f"bench_product_{tag[0]}{int(tag[1:])}.json")
    r1 = load_line(src / f"bench_reference_{tag[0]}{int(tag[1:])}.json")
    lines.append(f"`python bench.py --steps {p1['steps']} --warmup {p1['warmup']}` / `python bench.py --impl reference "
                 f"--steps {r1['steps']} --warmup {r1['warmup']}`, same box, back to back, SM clock "
                 f"{p1['clocks']['sm_mhz']:.0f} MHz, throttle reasons {p1['clocks']['reasons']}.\n")
    lines.append("| arm | value (frames/s, frames in HBM) | e2e (frames/s, pinned host frames in, cloud out) | launches / frame |")
    lines.append("|---|---:|---:|---:|")
    fp = p1["config"]["frames_per_step"]
    lines.append(f"| product | {p1['value']:.0f} | {p1['e2e']['value']:.0f} | {p1['gpu_launches'] / p1['steps'] / fp:.1f} |")
    lines.append(f"| reference kernels (sm_100a rebuild) | {r1['value']:.0f} | {r1['e2e']['value']:.0f} | {r1['gpu_launches'] / r1['steps'] / fp:.1f} |")
    lines.append(f"| ratio | {p1['value'] / r1['value']:.2f}x | {p1['e2e']['value'] / r1['e2e']['value']:.2f}x | |\n")
    n2p, n2r = src / "bench_product_n2.json", src / "bench_reference_n2.json"
    if n2p.exists() and n2r.exists():
        a, b = load_line(n2p), load_line(n2r)
        lines.append(f"Two GPUs (torchrun, one independent stream per rank, max-over-ranks time): product {a['value']:.0f} "
                     f"frames/s ({a['value'] / p1['value']:.2f}x of N = 1), reference {b['value']:.0f} frames/s.")
        n4 = src / "bench_product_n4.json"
        if n4.exists():
            c = load_line(n4)
            lines.append(f"Four GPUs: product {c['value']:.0f} frames/s ({c['value'] / p1['value']:.2f}x of N = 1; "
                         f"{c['ms_per_step']:.1f} ms per step on the slowest rank).")
        lines.append("")
    rf = p1.get("roofline")
    if rf:
        lines.append(f"Roofline object of the product line: dominant kernel `{rf['kernel']}`, {rf['achieved']:.0f} GB/s on its "
                     f"algorithmic bytes = {100 * rf['frac']:.1f} % of the measured HBM peak ({rf['peak']:.0f} GB/s); "
                     f"DRAM traffic per launch (ncu) {rf['traffic']}. The kernel is FP32/SFU-issue bound, see DESIGN.md §5.\n")
        lines.append("| kernel | launches | events µs (serial, host-launch bound) | share | pipelined µs (device timeline) |")
        lines.append("|---|---:|---:|---:|---:|")
        for k, v in p1["kernels"].items():
            lines.append(f"| {k} | {v['launches']} | {v['mean_us']:.2f} | {100 * v['share']:.1f} % | {v.get('pipelined_us', float('nan')):.2f} |")
        if "pipelined_frame_period_us" in rf:
            lines.append(f"\nFrame period inside the pipeline (project start to project start, median): {rf['pipelined_frame_period_us']:.1f} µs.")
        lines.append("")
    lines.append("## JSON lines\n")
    lines.append("```\n" + json.dumps(p1) + "\n```\n")
    lines.append("```\n" + json.dumps(r1) + "\n```\n")
    for arm, note in (("product", "frames ~416-450 of the first pass of `bench.py --steps 1 --warmup 3`, `-k regex:k_`"),
                      ("reference", "the same frames of `bench.py --impl reference`, `-k regex:Kernel`")):
        f = src / f"launches_{arm}_{tag[0]}{int(tag[1:])}.csv"
        if f.exists():
            shutil.copy(f, out / f"{tag}_launches_{arm}.csv")
            lines.append(f"## ncu launch list, {arm} ({note}; `--metrics gpu__time_duration.sum --clock-control none`, "
                         f"serialised, cold caches)\n\nRaw list: `profiles/{tag}_launches_{arm}.csv`.\n")
            lines.append(table(f))
    rep = sorted(src.glob(f"frame*_{tag[0]}{int(tag[1:])}.ncu-rep"))
    if rep:
        name = rep[-1].stem.split("_")[0]
        res = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(rep[-1]),
                              str(out / f"{tag}_{name}_ncu_full")], capture_output=True, text=True)
        lines.append(f"## `ncu --set full` capture (one launch per kernel, frame {name[5:]})\n\n"
                     f"`profiles/{tag}_{name}_ncu_full_summary.csv`, `..._traffic.json`.\n\n```\n{res.stdout}```\n")
    (out / f"{tag}_bench.md").write_text("\n".join(lines))
    print(out / f"{tag}_bench.md")


if __name__ == "__main__":
    main()
