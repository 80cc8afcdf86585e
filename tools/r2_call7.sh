#!/bin/bash
# round 2, GPU call 7: test suite with the warp-granular arrival order, race statistics (per-frame agreement with
# oracle A + free-running totals) around the new default, VGA scheduling / kernel A/B, C5b reference at 2000 frames,
# bench lines of both arms (C2)
mkdir -p gpurun_out
PY=python
timeout 900 $PY -m pytest tests -m gpu -q -s > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c7_pytest.log
grep -n "envelope:\|product-vs-A\|passed\|failed\|^FAILED\|^E  " gpurun_out/c7_pytest.log | cut -c1-300 | tail -60
timeout 1500 $PY tools/race_stats.py --out gpurun_out/c7_race_stats.json > gpurun_out/c7_race.log 2>&1; echo "race rc=$?" >> gpurun_out/c7_race.log
grep -A80 "per-frame agreement" gpurun_out/c7_race.log | tail -60
timeout 900 $PY tools/ab_probe.py --lib old4=variants/lib_old4.so --config old4:lib=old4 \
  --config prio1:SM_B200_GRAPH_PRIO=1 --config prio2:SM_B200_GRAPH_PRIO=2 \
  --config offchain50:SM_B200_OFFCHAIN_GRID_PERCENT=50 --config offchain75:SM_B200_OFFCHAIN_GRID_PERCENT=75 \
  --config prio1_offchain50:SM_B200_GRAPH_PRIO=1+SM_B200_OFFCHAIN_GRID_PERCENT=50 \
  --config plain_tiebreak:SM_B200_TIEBREAK=0,0,0 --config streams_r1:SM_B200_GRAPH=0 \
  --out gpurun_out/c7_ab.json > gpurun_out/c7_ab.log 2>&1
cat gpurun_out/c7_ab.log
timeout 600 $PY bench.py --steps 5 --warmup 3 > gpurun_out/c7_bench_product.json 2> gpurun_out/c7_bench_product.err
timeout 600 $PY bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/c7_bench_reference.json 2> gpurun_out/c7_bench_reference.err
timeout 900 $PY bench.py --impl reference --sigma-depth 0.05 --frames 2000 --cap 40000000 --required-inliers 1 --erosion-radius 0 --steps 2 --warmup 3 \
   > gpurun_out/c7_bench_reference_C5b_2000.json 2> gpurun_out/c7_bench_reference_C5b_2000.err
for f in gpurun_out/c7_bench_*.json; do echo $f; $PY tools/show_bench.py $f 2>/dev/null | head -1; done
