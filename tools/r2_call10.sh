#!/bin/bash
# round 2, GPU call 10: config 1 with the GPU in the loop (parity test + probe at scale)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_knn_gpu.py -m gpu -q -s > gpurun_out/c10_pytest_knn.log 2>&1; echo "rc=$?" >> gpurun_out/c10_pytest_knn.log
grep -n "triangles,\|passed\|failed\|^E  \|rc=" gpurun_out/c10_pytest_knn.log | cut -c1-300 | tail -12
timeout 900 python tools/meshing_probe.py --points 100000 --out gpurun_out/c10_meshing_probe_100k.json > gpurun_out/c10_meshing_probe_100k.log 2>&1; tail -1 gpurun_out/c10_meshing_probe_100k.log | cut -c1-900
timeout 1200 python tools/meshing_probe.py --points 1000000 --out gpurun_out/c10_meshing_probe_1m.json > gpurun_out/c10_meshing_probe_1m.log 2>&1; tail -1 gpurun_out/c10_meshing_probe_1m.log | cut -c1-900
