#!/bin/bash
# round 2, GPU call 13: wave boundaries at a per-pixel random phase: per-frame agreement + free-running totals (VGA 500 /
# 1000 frames, 1280x960 1000 frames)
mkdir -p gpurun_out
timeout 1500 python tools/race_stats.py --out gpurun_out/c13_race_stats.json > gpurun_out/c13_race.log 2>&1; echo "race rc=$?" >> gpurun_out/c13_race.log
grep -A60 "per-frame agreement" gpurun_out/c13_race.log | cut -c1-150
RULES="--rule default --rule 303104,0.01,0.0,32,1 --rule 303104,0.01,0.1,32,1 --rule 303104,0.01,0.25,32,1 --rule 303104,0.02,0.0,32,1 --rule 303104,0.02,0.1,32,1 --rule 303104,0.02,0.25,32,1 --rule 303104,0.03,0.1,32,1 --rule 303104,0.02,0.25,32,0"
timeout 900 python tools/free_running_check.py --frames 1000 $RULES --out gpurun_out/c13_free_vga1000.json > gpurun_out/c13_free_vga1000.log 2>&1; cat gpurun_out/c13_free_vga1000.log | cut -c1-220
timeout 1500 python tools/free_running_check.py --width 1280 --height 960 --frames 1000 --cap 20000000 $RULES --out gpurun_out/c13_free_hd1000.json > gpurun_out/c13_free_hd1000.log 2>&1; cat gpurun_out/c13_free_hd1000.log | cut -c1-220
