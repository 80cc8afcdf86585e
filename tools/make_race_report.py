#!/usr/bin/env python
"""profiles/r02_race_stats.md from the JSON tools/race_stats.py leaves in gpurun_out/."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
W = 296 * 1024


def pair_table(sets_path):
    """Who wins as a function of where the two threads sit in the reference's launch (tools/race_sets_fit.py)."""
    import numpy as np
    d = np.load(sets_path)
    n, winner, slots, sec = d["n"], d["winner"], d["slots"], d["secondary"]
    two = n == 2
    a, b = slots[two, 0].astype(np.int64), slots[two, 1].astype(np.int64)
    low = winner[two] == 0
    sa, sb = (sec[two] & 1).astype(bool), ((sec[two] >> 1) & 1).astype(bool)
    same_kind = sa == sb
    wave_a, wave_b = a // W, b // W
    ra, rb = a % W, b % W
    same_wave = same_kind & (wave_a == wave_b)
    same_block = same_wave & (ra // 1024 == rb // 1024)
    same_warp = same_block & (ra // 32 == rb // 32)
    db = rb // 1024 - ra // 1024
    rows = [("different launch waves", same_kind & (wave_a != wave_b)),
            ("same wave, same warp (32 consecutive slots)", same_warp),
            ("same wave, same block, different warps", same_block & ~same_warp),
            ("same wave, adjacent blocks", same_wave & (db == 1)),
            ("same wave, 2 - 7 blocks apart", same_wave & (db >= 2) & (db <= 7)),
            ("same wave, 8 - 31 blocks apart", same_wave & (db >= 8) & (db <= 31)),
            ("same wave, 32 - 295 blocks apart", same_wave & (db >= 32))]
    out = ["| the two supporters (same kind) sit in | pairs | lower slot wins |", "|---|---:|---:|"]
    for name, m in rows:
        out.append(f"| {name} | {int(m.sum())} | {100 * low[m].mean():.1f} % |")
    mixed = (sa != sb) & (wave_a == wave_b)
    prim = np.where(sa, ~low, low)
    out.append(f"\nOne primary and one secondary association in one wave ({int(mixed.sum())} pairs): the primary wins "
               f"{100 * prim[mixed].mean():.1f} % ({100 * prim[mixed & (ra // 1024 == rb // 1024)].mean():.1f} % when both sit in one block).")
    out += ["\nThe first wave (its 296 blocks start together) against the second (its blocks start one by one as blocks of the "
            "first retire):\n", "| wave | same-kind pairs | lower slot wins | mixed pairs | secondary wins | ... when its block is > 32 blocks "
            "ahead of the primary's | ... 1 - 32 blocks ahead | ... in the same or a later block |", "|---|---:|---:|---:|---:|---:|---:|---:|"]
    sec_block = np.where(sa, ra, rb) // 1024
    prim_block = np.where(sa, rb, ra) // 1024
    lead = prim_block - sec_block
    for w in (0, 1):
        sk, mx = same_wave & (wave_a == w), mixed & (wave_a == w)
        if sk.sum() == 0 or mx.sum() == 0:
            continue
        cell = lambda m: f"{100 * (1 - prim[m].mean()):.1f} %" if m.sum() else "-"
        out.append(f"| {w} | {int(sk.sum())} | {100 * low[sk].mean():.1f} % | {int(mx.sum())} | {cell(mx)} | {cell(mx & (lead > 32))} | "
                   f"{cell(mx & (lead >= 1) & (lead <= 32))} | {cell(mx & (lead <= 0))} |")
    return out, int(two.sum())


def free_table(free, out):
    out += ["| run | surfels_size (slots) | surfel_count (live) | merged |", "|---|---:|---:|---:|"]
    for k, v in free.items():
        out.append(f"| {k} | {v[0]} | {v[1]} | {v[0] - v[1]} |")
    out.append("")


def main():
    src = Path(sys.argv[1] if len(sys.argv) > 1 else ROOT / "gpurun_out" / "c14_race_stats.json")
    earlier = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "gpurun_out" / "c6_race_stats.json"
    j = json.loads(src.read_text())
    r, samples, free = j["race"], j["samples"], j["free"]
    out = ["# r02: how the reference's supporting-surfel race resolves, and the rule that reproduces it\n",
           f"Source: `tools/race_stats.py` on a B200 (`{src.name}`); 500-frame VGA benchmark stream, {len(samples)} teacher-forced "
           "sample frames (the state before the frame is copied from oracle A into oracle B and into the product, once per "
           "candidate rule), supporter sets from `oracle/cpu_walk.c`, winners from oracle A's raster.\n",
           "## Who wins the reference's `atomicCAS` (kernels.cu:1688)\n",
           f"* contested pixels analysed: {r['multi_pixels']} (winner outside the CPU-computed set: {r['winner_not_in_set']}; "
           f"set size differs from the GPU's count, skipped: {r['set_size_mismatch']})",
           f"* two runs of the reference on the same state pick the same winner on {100 * r['contested_same_winner_ab'] / r['contested_pixels_ab']:.1f} % "
           f"of the {r['contested_pixels_ab']} contested pixels: that is the reference's own reproducibility",
           f"* primary and secondary associations on one pixel: a secondary won {100 * r['mixed_secondary_wins'] / r['mixed']:.2f} % "
           f"of {r['mixed']} contests overall, but only {100 * r['mixed_samewave_secondary_wins'] / r['mixed_samewave']:.2f} % of the "
           f"{r['mixed_samewave']} whose supporters share a wave: a secondary wins when it sits in an earlier wave\n"]
    sets = src.with_name(src.stem + "_sets.npz")
    if sets.exists():
        table, pairs = pair_table(sets)
        out.append(f"Pixels with exactly two supporters ({pairs} of them), by the position of the two threads in the reference's launch "
                   f"(1024-thread blocks, a wave = {W} slots = 296 resident blocks on a B200):\n")
        out += table
    out += ["\nSo the race resolves by launch wave first, primary before secondary second; two lanes of one warp issue their "
            "compare-and-swap in lane order; the warps of a block arrive in no particular order; across the blocks of a wave the "
            "lower block is ahead more often the further apart they are (launch stagger against memory-latency jitter). "
            "Round 1's rule (secondary bit above everything, then lowest slot) hands every contested pixel of the newest "
            "surfels (second wave, created in the current view) to their own primary association instead of an older surfel's "
            "secondary one; that removes merge candidates, hence its systematic deficit of merges and of new surfels.\n",
            "## The product's rule and its parameters\n",
            "Arrival key, most significant first: launch wave of the slot; a late bit for secondary associations (all but a "
            "pseudo-random fraction q of them); inside a wave, for a fraction b of the pixels (per-frame hash) plain slot order, "
            "for the rest a per-frame pseudo-random order of the wave's warps (32 consecutive slots) with the lanes of a warp in "
            "order. `wave_qQ_bB`: one (q, b) for all waves; `..._later_qQ1_bB1`: (Q, B) for the first wave, (Q1, B1) for the later "
            "ones.\n",
            "### Per-frame agreement with oracle A (what the envelope tests bound)\n",
            "| rule | same winner as oracle A on the contested pixels | differing merge flags over the sampled frames | vs oracle B |",
            "|---|---:|---:|---:|"]
    agree, flags = j.get("same_winner_as_oracle_a", {}), j.get("differing_merge_flags_vs_oracle_a", {})
    for k in agree:
        out.append(f"| {k} | {100 * agree[k] / r['contested_pixels_ab']:.2f} % | {flags[k]} | {flags[k] / max(flags['oracle_b'], 1):.2f}x |")
    out += ["", "### Merge counts on the teacher-forced frames\n",
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
    out += ["", "### Free-running totals after the 492 integrated frames\n"]
    free_table(free, out)
    out += ["Chosen default (`csrc/sm_handle.cuh`): waves of 296 x 1024 slots, warps kept intact, (q, b) = (1 %, 25 %) in the first wave, "
            "(1.5 %, 45 %) in the second and (3 %, 45 %) in the later ones (the tables above have `wave_q0.01_b0.25_later_q0.03_b0.45`, i.e. "
            "3 % in the second wave too; the free-running tables below the final choice). The totals of the ORACLE differ between "
            "processes by more than within one (merges of the 500-frame stream: 118 755 - 119 721 over the runs of this round), which is "
            "why the second sweep ran the 500-frame stream in two processes. `tests/test_round2_gpu.py::test_free_running_"
            "stream_inside_the_reference_envelope` asserts |product - mean(oracle)| <= 3 x the oracle's spread (+ 0.1 %) against three fresh "
            "oracle runs; `...::test_race_bound_rows_inside_the_reference_envelope` and `tests/test_parity_gpu.py` assert the per-frame rows "
            "at 2 x oracle B (+ 4 sigma of the count + a floor).\n"]
    for tag, title in (("vga500", "640x480, 500 frames"), ("vga1000", "640x480, 1000 frames"), ("hd1000", "1280x960, 1000 frames, 20 M cap")):
        rows = []
        for call, label in (("c14", "sweep 1 (one fraction for all later waves)"), ("c17", "sweep 2 (second wave separately)"),
                            ("c17", "sweep 2, another process"), ("c18", "final defaults")):
            suffix = "" if not (call == "c17" and tag == "vga500") else ("_2" if "another" in label else "_1")
            if "another" in label and tag != "vga500":
                continue
            f = ROOT / "gpurun_out" / f"{call}_free_{tag}{suffix}.json"
            if f.exists():
                rows.append((label, json.loads(f.read_text())))
        if not rows:
            continue
        out += [f"### Free-running totals, {title} (`tools/free_running_check.py`)\n",
                "| run | rule (wave, q, b, lanes, phase, q later, b later, q second wave) | slots | live | merged | deviation from the oracle mean in oracle spreads | relative |",
                "|---|---|---:|---:|---:|---|---|"]
        for label, j in rows:
            sp = j["oracle_spread"]
            out.append(f"| {label}: three oracle runs | - | " + " / ".join(str(r[0]) for r in j["oracle_runs"]) + " | " +
                       " / ".join(str(r[1]) for r in j["oracle_runs"]) + " | " + " / ".join(str(r[2]) for r in j["oracle_runs"]) +
                       f" | spread {sp} | |")
            for name, v in j["rules"].items():
                t, d, rel = v["totals"], v["deviation_in_oracle_spreads"], v["relative"]
                out.append(f"| {label} | {name} | {t[0]} | {t[1]} | {t[2]} | {d[0]:+.1f} / {d[1]:+.1f} / {d[2]:+.1f} | "
                           f"{100 * rel[0]:+.3f} % / {100 * rel[1]:+.3f} % / {100 * rel[2]:+.3f} % |")
        out.append("")
    phase = ROOT / "gpurun_out" / "c13_race_stats.json"
    if phase.exists():
        pj = json.loads(phase.read_text())
        ag, fl, cp = pj["same_winner_as_oracle_a"], pj["differing_merge_flags_vs_oracle_a"], pj["race"]["contested_pixels_ab"]
        out += ["### Tried and dropped: wave boundaries at a per-pixel random phase (`c13_race_stats.json`)\n",
                "If the ordering probability grew smoothly with the slot distance, cutting the slot axis into waves at a random phase "
                "per pixel would model it. It does not: the boundaries at multiples of 303 104 slots are real.\n",
                "| rule | same winner as oracle A | differing merge flags vs oracle B |", "|---|---:|---:|"]
        for k in ag:
            out.append(f"| {k} | {100 * ag[k] / cp:.2f} % | {fl[k] / max(fl['oracle_b'], 1):.2f}x |")
        out.append("")
    if earlier.exists():
        e = json.loads(earlier.read_text())
        out += [f"### Earlier sweep with every slot shuffled (l = 1; `{earlier.name}`)\n",
                "`wave_qQ_bB` there: the same rule without the warp structure (every slot of a wave shuffled). It matched the free-running totals equally well (the "
                "totals depend on the marginal win rates, which both versions reproduce) but differed from oracle A on 1.5 x as "
                "many merge flags as oracle B does (2.3 - 8 x on single frames), because a quarter of the same-wave contests are "
                "between lanes of one warp, which the reference resolves the same way every time.\n"]
        free_table(e["free"], out)
    (ROOT / "profiles" / "r02_race_stats.md").write_text("\n".join(out) + "\n")
    print(ROOT / "profiles" / "r02_race_stats.md")


if __name__ == "__main__":
    main()
