#!/bin/bash
# round 2, GPU call 1: tests after the restructure, race statistics, graph vs stream pipeline A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_round2_gpu.py > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -15 gpurun_out/c1_pytest.log
timeout 900 python -m pytest tests/test_round2_gpu.py -m gpu -q > gpurun_out/c1_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest2.log
tail -40 gpurun_out/c1_pytest2.log
for mode in "SM_B200_GRAPH=1" "SM_B200_SPLIT_PROJECT=0" "SM_B200_GRAPH=0" "SM_B200_GRAPH_PDL=0" "SM_B200_GRAPH_PDL=2"; do
  echo "== $mode" >> gpurun_out/c1_bench.log
  env $mode timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline >> gpurun_out/c1_bench.log 2>&1
done
python - <<'PY'
import json
for line in open('gpurun_out/c1_bench.log'):
    if line.startswith('=='): print(line.strip())
    elif line.startswith('{'):
        j=json.loads(line); print('  value', round(j['value']), 'e2e', round(j['e2e']['value']), 'surfels', j['config']['surfels_after_step'], 'launches', j['gpu_launches'])
    else: print('  ', line.strip()[:200])
PY
timeout 300 python tools/timeline_probe.py --out gpurun_out/c1_timeline.csv > gpurun_out/c1_timeline.log 2>&1
tail -60 gpurun_out/c1_timeline.log
timeout 900 python tools/race_stats.py > gpurun_out/c1_race.log 2>&1; echo "race rc=$?" >> gpurun_out/c1_race.log
tail -60 gpurun_out/c1_race.log
