#!/bin/bash
# k-NN query with the append-then-sort path: parity, memcheck, throughput
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_knn_gpu.py -m gpu -q > gpurun_out/c8b_pytest_knn.log 2>&1; echo "rc=$?" >> gpurun_out/c8b_pytest_knn.log; tail -5 gpurun_out/c8b_pytest_knn.log | cut -c1-300
timeout 600 python tools/knn_probe.py --out gpurun_out/c8b_knn_probe.json > gpurun_out/c8b_knn_probe.log 2>&1; tail -1 gpurun_out/c8b_knn_probe.log | cut -c1-700
timeout 600 python tools/knn_probe.py --cell-factor 1.0 --cpu-sample 20000 --out gpurun_out/c8b_knn_probe_cell10.json > gpurun_out/c8b_knn_probe_cell10.log 2>&1; tail -1 gpurun_out/c8b_knn_probe_cell10.log | cut -c1-400
timeout 600 python tools/knn_probe.py --radius-factor 5 --cpu-sample 20000 --out gpurun_out/c8b_knn_probe_r5.json > gpurun_out/c8b_knn_probe_r5.log 2>&1; tail -1 gpurun_out/c8b_knn_probe_r5.log | cut -c1-400
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_knn_gpu.py -m gpu -q -k "golden or empty or duplicates" > gpurun_out/c8b_memcheck.log 2>&1; tail -3 gpurun_out/c8b_memcheck.log | cut -c1-200
