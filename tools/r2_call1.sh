#!/bin/bash
# round 2, GPU call 1: tests, pipeline / kernel A/B, timeline, race statistics
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_round2_gpu.py > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -15 gpurun_out/c1_pytest.log
timeout 1200 python -m pytest tests/test_round2_gpu.py -m gpu -q > gpurun_out/c1_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest2.log
tail -40 gpurun_out/c1_pytest2.log
timeout 900 python tools/ab_probe.py --lib base=variants/lib_r2base.so \
  --config r2base_lib:lib=base --config tail_vector:SM_B200_TAIL_FILL=vector --config no_split:SM_B200_SPLIT_PROJECT=0 \
  --config graph_pdl0:SM_B200_GRAPH_PDL=0 --config graph_pdl2:SM_B200_GRAPH_PDL=2 --config streams_r1:SM_B200_GRAPH=0 \
  --config streams_r1_base:SM_B200_GRAPH=0,lib=base --out gpurun_out/c1_ab.json > gpurun_out/c1_ab.log 2>&1
cat gpurun_out/c1_ab.log
timeout 300 python tools/timeline_probe.py --out gpurun_out/c1_timeline.csv > gpurun_out/c1_timeline.log 2>&1
tail -60 gpurun_out/c1_timeline.log
timeout 900 python tools/race_stats.py > gpurun_out/c1_race.log 2>&1; echo "race rc=$?" >> gpurun_out/c1_race.log
tail -60 gpurun_out/c1_race.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c1_bench_product.json 2> gpurun_out/c1_bench_product.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/c1_bench_reference.json 2> gpurun_out/c1_bench_reference.err
python tools/show_bench.py gpurun_out/c1_bench_product.json gpurun_out/c1_bench_reference.json 2>/dev/null | head -40
