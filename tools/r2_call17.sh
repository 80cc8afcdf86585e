#!/bin/bash
# round 2, GPU call 17: early fraction of the second wave separately from the later ones (free-running totals at the three
# stream sizes, two processes each for the 500-frame stream: the oracle's totals differ between processes), the new
# host-array k-NN entry point
mkdir -p gpurun_out
B="303104,0.01,0.25,32,0"
RULES="--rule default --rule $B,0.03,0.45,0.015 --rule $B,0.03,0.45,0.02 --rule $B,0.035,0.45,0.02 --rule $B,0.04,0.45,0.02 --rule $B,0.04,0.45,0.015 --rule $B,0.02,0.45"
for i in 1 2; do
timeout 900 python tools/free_running_check.py --frames 500 --oracle-runs 4 $RULES --out gpurun_out/c17_free_vga500_$i.json > gpurun_out/c17_free_vga500_$i.log 2>&1; cat gpurun_out/c17_free_vga500_$i.log | cut -c1-200
done
timeout 900 python tools/free_running_check.py --frames 1000 $RULES --out gpurun_out/c17_free_vga1000.json > gpurun_out/c17_free_vga1000.log 2>&1; cat gpurun_out/c17_free_vga1000.log | cut -c1-200
timeout 1500 python tools/free_running_check.py --width 1280 --height 960 --frames 1000 --cap 20000000 $RULES --out gpurun_out/c17_free_hd1000.json > gpurun_out/c17_free_hd1000.log 2>&1; cat gpurun_out/c17_free_hd1000.log | cut -c1-200
timeout 600 python -m pytest tests/test_knn_gpu.py -m gpu -q -k "batch_host or feeds" > gpurun_out/c17_pytest_knn.log 2>&1; tail -3 gpurun_out/c17_pytest_knn.log | cut -c1-300
timeout 1200 python tools/meshing_probe.py --points 1000000 --out gpurun_out/c17_meshing_probe_1m.json > gpurun_out/c17_meshing_probe_1m.log 2>&1; tail -1 gpurun_out/c17_meshing_probe_1m.log | cut -c1-1400
