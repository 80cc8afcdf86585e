#!/bin/bash
# round 2, GPU call 19: host-array k-NN entry point with the pinned bounce download (parity, memcheck, timing)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_knn_gpu.py -m gpu -q > gpurun_out/c19_pytest_knn.log 2>&1; echo "rc=$?" >> gpurun_out/c19_pytest_knn.log; tail -3 gpurun_out/c19_pytest_knn.log | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_knn_gpu.py -m gpu -q -k "batch_host" > gpurun_out/c19_memcheck.log 2>&1; tail -3 gpurun_out/c19_memcheck.log | cut -c1-200
timeout 900 python tools/meshing_probe.py --points 100000 --out gpurun_out/c19_meshing_probe_100k.json > gpurun_out/c19_meshing_probe_100k.log 2>&1; tail -1 gpurun_out/c19_meshing_probe_100k.log | cut -c1-1400
timeout 1200 python tools/meshing_probe.py --points 1000000 --out gpurun_out/c19_meshing_probe_1m.json > gpurun_out/c19_meshing_probe_1m.log 2>&1; tail -1 gpurun_out/c19_meshing_probe_1m.log | cut -c1-1400
