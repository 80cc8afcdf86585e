#!/bin/bash
# round 2, GPU call 14: wave-dependent early / slot-order fractions (first wave vs later waves)
mkdir -p gpurun_out
timeout 1500 python tools/race_stats.py --out gpurun_out/c14_race_stats.json > gpurun_out/c14_race.log 2>&1; echo "race rc=$?" >> gpurun_out/c14_race.log
grep -A60 "per-frame agreement" gpurun_out/c14_race.log | cut -c1-170
B="303104,0.01,0.25,32,0"
RULES="--rule default --rule $B,0.03 --rule $B,0.04 --rule $B,0.06 --rule $B,0.03,0.45 --rule $B,0.04,0.45 --rule $B,0.06,0.45 --rule $B,0.04,0.7 --rule $B,0.08,0.7"
timeout 900 python tools/free_running_check.py --frames 1000 $RULES --out gpurun_out/c14_free_vga1000.json > gpurun_out/c14_free_vga1000.log 2>&1; cat gpurun_out/c14_free_vga1000.log | cut -c1-220
timeout 1500 python tools/free_running_check.py --width 1280 --height 960 --frames 1000 --cap 20000000 $RULES --out gpurun_out/c14_free_hd1000.json > gpurun_out/c14_free_hd1000.log 2>&1; cat gpurun_out/c14_free_hd1000.log | cut -c1-220
