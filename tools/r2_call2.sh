#!/bin/bash
# round 2, GPU call 2: tests, A/B combos, timeline, race statistics (fixed wave-major key), bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_round2_gpu.py > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
tail -12 gpurun_out/c2_pytest.log
timeout 1500 python -m pytest tests/test_round2_gpu.py -m gpu -q > gpurun_out/c2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest2.log
tail -60 gpurun_out/c2_pytest2.log
timeout 900 python tools/ab_probe.py --lib reg5=variants/lib_reg5.so --lib reg6=variants/lib_reg6.so \
  --config split:SM_B200_SPLIT_PROJECT=1 --config pdl1:SM_B200_GRAPH_PDL=1 --config tail_vector:SM_B200_TAIL_FILL=vector \
  --config carve_default:SM_B200_CARVEOUT=-1 --config carve100:SM_B200_CARVEOUT=100 --config carve64:SM_B200_CARVEOUT=64 \
  --config reg5:lib=reg5 --config reg6:lib=reg6 --config streams_r1:SM_B200_GRAPH=0 \
  --out gpurun_out/c2_ab.json > gpurun_out/c2_ab.log 2>&1
cat gpurun_out/c2_ab.log
timeout 300 python tools/timeline_probe.py --out gpurun_out/c2_timeline.csv > gpurun_out/c2_timeline.log 2>&1
tail -48 gpurun_out/c2_timeline.log
timeout 900 python tools/race_stats.py --out gpurun_out/c2_race_stats.json > gpurun_out/c2_race.log 2>&1; echo "race rc=$?" >> gpurun_out/c2_race.log
tail -70 gpurun_out/c2_race.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c2_bench_product.json 2> gpurun_out/c2_bench_product.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/c2_bench_reference.json 2> gpurun_out/c2_bench_reference.err
python tools/show_bench.py gpurun_out/c2_bench_product.json | head -30
