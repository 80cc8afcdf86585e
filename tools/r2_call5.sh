#!/bin/bash
# round 2, GPU call 5: full test suite + smoke, VGA scheduling / kernel A/B, race statistics with the supporter-set
# dump, bench lines of both arms (C2, C3, C5b), delta-transfer probes, ncu launch lists + full captures,
# compute-sanitizer passes
mkdir -p gpurun_out
PY=python
timeout 900 $PY -m pytest tests -m gpu -q -s > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
tail -30 gpurun_out/c5_pytest.log
timeout 300 $PY -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c5_smoke.log 2>&1; tail -2 gpurun_out/c5_smoke.log
timeout 900 $PY tools/ab_probe.py --lib old4=variants/lib_old4.so --config old4:lib=old4 \
  --config prio1:SM_B200_GRAPH_PRIO=1 --config prio2:SM_B200_GRAPH_PRIO=2 \
  --config offchain50:SM_B200_OFFCHAIN_GRID_PERCENT=50 --config offchain75:SM_B200_OFFCHAIN_GRID_PERCENT=75 \
  --config prio1_offchain50:SM_B200_GRAPH_PRIO=1+SM_B200_OFFCHAIN_GRID_PERCENT=50 \
  --config plain_tiebreak:SM_B200_TIEBREAK=0,0,0 --config streams_r1:SM_B200_GRAPH=0 \
  --out gpurun_out/c5_ab.json > gpurun_out/c5_ab.log 2>&1
cat gpurun_out/c5_ab.log
timeout 900 $PY tools/race_stats.py --out gpurun_out/c5_race_stats.json > gpurun_out/c5_race.log 2>&1; echo "race rc=$?" >> gpurun_out/c5_race.log
tail -45 gpurun_out/c5_race.log
timeout 600 $PY bench.py --steps 5 --warmup 3 > gpurun_out/c5_bench_product.json 2> gpurun_out/c5_bench_product.err
timeout 600 $PY bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/c5_bench_reference.json 2> gpurun_out/c5_bench_reference.err
timeout 1500 $PY bench.py --width 1280 --height 960 --frames 1000 --cap 20000000 --steps 3 --warmup 3 --no-cpu-baseline \
   > gpurun_out/c5_bench_product_C3.json 2> gpurun_out/c5_bench_product_C3.err
timeout 1500 $PY bench.py --impl reference --width 1280 --height 960 --frames 1000 --cap 20000000 --steps 2 --warmup 3 \
   > gpurun_out/c5_bench_reference_C3.json 2> gpurun_out/c5_bench_reference_C3.err
timeout 900 $PY bench.py --sigma-depth 0.05 --frames 2000 --required-inliers 1 --erosion-radius 0 --steps 3 --warmup 3 --no-cpu-baseline \
   > gpurun_out/c5_bench_product_C5b.json 2> gpurun_out/c5_bench_product_C5b.err
timeout 900 $PY bench.py --impl reference --sigma-depth 0.05 --frames 2000 --required-inliers 1 --erosion-radius 0 --steps 2 --warmup 3 \
   > gpurun_out/c5_bench_reference_C5b.json 2> gpurun_out/c5_bench_reference_C5b.err
for f in gpurun_out/c5_bench_*.json; do echo $f; $PY tools/show_bench.py $f 2>/dev/null | head -2; done
timeout 600 $PY tools/transfer_probe.py --every 5 --out gpurun_out/c5_transfer_probe_every5.json > gpurun_out/c5_transfer_every5.log 2>&1
timeout 900 $PY tools/transfer_probe.py --width 1280 --height 960 --frames 400 --cap 20000000 --every 10 --out gpurun_out/c5_transfer_probe_hd.json > gpurun_out/c5_transfer_hd.log 2>&1
cat gpurun_out/c5_transfer_every5.log gpurun_out/c5_transfer_hd.log
# launch lists of the bench command itself (late frames of its first pass); a number printed under ncu is not a bench value
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --launch-skip 5280 -c 400 --csv \
   --log-file gpurun_out/c5_launches_product.csv $PY bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/c5_ncu_product.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:Kernel --launch-skip 17160 -c 1330 --csv \
   --log-file gpurun_out/c5_launches_reference.csv $PY bench.py --impl reference --steps 1 --warmup 3 > gpurun_out/c5_ncu_reference.log 2>&1
SM_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 5520 -c 13 -f \
   -o gpurun_out/c5_frame460_C2 $PY tools/stream_probe.py --frames 500 --reps 1 > gpurun_out/c5_ncu_full_C2.log 2>&1
SM_B200_GRAPH=0 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 11520 -c 13 -f \
   -o gpurun_out/c5_frame960_C3 $PY tools/stream_probe.py --width 1280 --height 960 --frames 1000 --cap 20000000 --reps 1 > gpurun_out/c5_ncu_full_C3.log 2>&1
tail -2 gpurun_out/c5_ncu_full_C3.log
timeout 900 compute-sanitizer --tool memcheck $PY -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -m gpu -q \
   -k "golden or ragged or empty_cloud or timeline or median_densify_bit_exact or delta_transfer or visualization or vis_depth" > gpurun_out/c5_memcheck.log 2>&1
tail -5 gpurun_out/c5_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all $PY -m pytest tests/test_parity_gpu.py -m gpu -q \
   -k "teacher_forced_against_golden or fused_preprocess" > gpurun_out/c5_racecheck.log 2>&1
tail -5 gpurun_out/c5_racecheck.log
