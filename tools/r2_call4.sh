#!/bin/bash
# round 2, GPU call 4: full test suite (old-gather-order build first, then the default build with the light-first-batch
# kernels), smoke, scheduling / kernel A/B at VGA and at 1280x960, race statistics around the tie-break default,
# bench lines (C2) of both arms
mkdir -p gpurun_out
SM_B200_LIB=variants/lib_old4.so timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c4_pytest_old4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest_old4.log
tail -30 gpurun_out/c4_pytest_old4.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
tail -30 gpurun_out/c4_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c4_smoke.log 2>&1; tail -2 gpurun_out/c4_smoke.log
timeout 900 python tools/ab_probe.py --lib old4=variants/lib_old4.so --config old4:lib=old4 \
  --config prio1:SM_B200_GRAPH_PRIO=1 --config prio2:SM_B200_GRAPH_PRIO=2 \
  --config offchain50:SM_B200_OFFCHAIN_GRID_PERCENT=50 --config offchain75:SM_B200_OFFCHAIN_GRID_PERCENT=75 \
  --config prio1_offchain50:SM_B200_GRAPH_PRIO=1,SM_B200_OFFCHAIN_GRID_PERCENT=50 \
  --config plain_tiebreak:SM_B200_TIEBREAK=0,0,0 --config streams_r1:SM_B200_GRAPH=0 \
  --out gpurun_out/c4_ab.json > gpurun_out/c4_ab.log 2>&1
cat gpurun_out/c4_ab.log
timeout 900 python tools/ab_probe.py --width 1280 --height 960 --frames 400 --cap 20000000 --reps 3 --lib old4=variants/lib_old4.so \
  --config old4:lib=old4 --config prio1:SM_B200_GRAPH_PRIO=1 --config streams_r1:SM_B200_GRAPH=0 \
  --out gpurun_out/c4_ab_hd.json > gpurun_out/c4_ab_hd.log 2>&1
cat gpurun_out/c4_ab_hd.log
timeout 900 python tools/race_stats.py --out gpurun_out/c4_race_stats.json > gpurun_out/c4_race.log 2>&1; echo "race rc=$?" >> gpurun_out/c4_race.log
tail -45 gpurun_out/c4_race.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c4_bench_product.json 2> gpurun_out/c4_bench_product.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/c4_bench_reference.json 2> gpurun_out/c4_bench_reference.err
python tools/show_bench.py gpurun_out/c4_bench_product.json | head -3
