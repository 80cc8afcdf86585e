#!/bin/bash
# round 2, GPU call 12: free-running totals at 1000 VGA frames and at 1280x960 under candidate rules; final
# compute-sanitizer passes over the parity tests
mkdir -p gpurun_out
RULES="--rule default --rule 303104,0.0,0.25,32 --rule 303104,0.02,0.25,32 --rule 303104,0.03,0.25,32 --rule 303104,0.01,0.0,32 --rule 303104,0.01,0.5,32 --rule 0,0,0,1"
timeout 900 python tools/free_running_check.py --frames 1000 $RULES --out gpurun_out/c12_free_vga1000.json > gpurun_out/c12_free_vga1000.log 2>&1; cat gpurun_out/c12_free_vga1000.log | cut -c1-220
timeout 1500 python tools/free_running_check.py --width 1280 --height 960 --frames 1000 --cap 20000000 $RULES --out gpurun_out/c12_free_hd1000.json > gpurun_out/c12_free_hd1000.log 2>&1; cat gpurun_out/c12_free_hd1000.log | cut -c1-220
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -m gpu -q \
   -k "golden or ragged or empty_cloud or timeline or median_densify_bit_exact or delta_transfer or visualization or vis_depth" > gpurun_out/c12_memcheck.log 2>&1
tail -4 gpurun_out/c12_memcheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_parity_gpu.py -m gpu -q \
   -k "teacher_forced_against_golden or fused_preprocess" > gpurun_out/c12_racecheck.log 2>&1
tail -4 gpurun_out/c12_racecheck.log | cut -c1-200
