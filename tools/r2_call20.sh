#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/meshing_probe.py --points 100000 --out gpurun_out/c20_meshing_probe_100k.json > gpurun_out/c20_meshing_probe_100k.log 2>&1; tail -1 gpurun_out/c20_meshing_probe_100k.log | cut -c1-200
timeout 1200 python tools/meshing_probe.py --points 1000000 --out gpurun_out/c20_meshing_probe_1m.json > gpurun_out/c20_meshing_probe_1m.log 2>&1; tail -1 gpurun_out/c20_meshing_probe_1m.log | cut -c1-200
