#!/usr/bin/env python
"""profiles/r02_race_stats.md from the JSON tools/race_stats.py leaves in gpurun_out/."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
W = 296 * 1024


def main():
    src = Path(sys.argv[1] if len(sys.argv) > 1 else ROOT / "gpurun_out" / "c2_race_stats.json")
    j = json.loads(src.read_text())
    r, samples, free = j["race"], j["samples"], j["free"]
    out = ["# r02: how the reference's supporting-surfel race resolves, and the rule that reproduces it\n",
           f"Source: `tools/race_stats.py` on a B200 (`{src.name}`); 500-frame VGA benchmark stream, {len(samples)} teacher-forced "
           "sample frames (the state before the frame is copied from oracle A into oracle B and into the product, once per "
           "candidate rule), supporter sets from `oracle/cpu_walk.c`, winners from oracle A's raster.\n",
           "## Who wins the reference's `atomicCAS` (kernels.cu:1688)\n",
           f"* contested pixels analysed: {r['multi_pixels']} (winner outside the CPU-computed set: {r['winner_not_in_set']}; "
           f"set size differs from the GPU's count, skipped: {r['set_size_mismatch']})",
           f"* two supporters of the same kind in DIFFERENT launch waves ({W} slots = 296 blocks of 1024 threads): the earlier wave "
           f"won {r['pair_two_waves_lower_wave_wins']} of {r['pair_two_waves']}",
           f"* same kind, same wave: the lower slot won {100 * r['pair_one_wave_lower_index_wins'] / r['pair_one_wave']:.1f} % of "
           f"{r['pair_one_wave']} pairs ({100 * r['pair_one_wave_near_lower_wins'] / r['pair_one_wave_near']:.1f} % when less than "
           f"32 blocks apart, {100 * r['pair_one_wave_far_lower_wins'] / r['pair_one_wave_far']:.1f} % when further)",
           f"* primary and secondary associations on one pixel: a secondary won {100 * r['mixed_secondary_wins'] / r['mixed']:.2f} % "
           f"of {r['mixed']} contests overall, but only {100 * r['mixed_samewave_secondary_wins'] / r['mixed_samewave']:.2f} % of the "
           f"{r['mixed_samewave']} whose supporters share a wave: a secondary wins when it sits in an earlier wave\n",
           "So the race resolves by launch wave first, primary before secondary second, and inside a wave mostly - not always - by "
           "slot. Round 1's rule (secondary bit above everything, then lowest slot) hands every contested pixel of the newest "
           "surfels (second wave, created in the current view) to their own primary association instead of an older surfel's "
           "secondary one; that removes merge candidates, hence the systematic deficit of merges and of new surfels.\n",
           "## Merge counts of the candidate rules on the teacher-forced frames\n",
           "`wave_qQ_bB` = waves of 303 104 slots, fraction Q of the secondaries competes like primaries, fraction B of the "
           "pixels orders a wave by slot (the rest by a per-frame random permutation); `plain` = round 1; oracle B = a second run "
           "of the reference on the same state (its own envelope).\n",
           "| rule | merges (all samples) | vs oracle A | N < 303 k | N >= 303 k |", "|---|---:|---:|---:|---:|"]
    names = [k for k in samples[0] if k not in ("frame", "n_before")]
    tot_a = sum(e["oracle_a"] for e in samples)
    lo = [e for e in samples if e["n_before"] < W]
    hi = [e for e in samples if e["n_before"] >= W]
    for k in sorted(names, key=lambda k: sum(e[k] for e in samples)):
        t = sum(e[k] for e in samples)
        dl = sum(e[k] - e["oracle_a"] for e in lo)
        dh = sum(e[k] - e["oracle_a"] for e in hi)
        out.append(f"| {k} | {t} | {100 * (t - tot_a) / tot_a:+.2f} % | {dl:+d} | {dh:+d} |")
    out += ["", "## Free-running totals after the 492 integrated frames\n",
            "| run | surfels_size (slots) | surfel_count (live) | merged |", "|---|---:|---:|---:|"]
    for k, v in free.items():
        out.append(f"| {k} | {v[0]} | {v[1]} | {v[0] - v[1]} |")
    out += ["", "Chosen default (`csrc/sm_handle.cuh`): waves of 296 x 1024 slots, 1 % early secondaries, 44 % of the pixels in slot "
            "order - between `wave_q0.0_b0.44` and `wave_q0.02_b0.44` above, which bracket the oracle on slots, live surfels and "
            "merges. `tests/test_round2_gpu.py::test_free_running_stream_inside_the_reference_envelope` asserts the result "
            "against three oracle runs."]
    (ROOT / "profiles" / "r02_race_stats.md").write_text("\n".join(out) + "\n")
    print(ROOT / "profiles" / "r02_race_stats.md")


if __name__ == "__main__":
    main()
