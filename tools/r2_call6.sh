#!/bin/bash
# round 2, GPU call 6: full test suite (incl. the k-NN parity tests) + smoke, VGA scheduling / kernel A/B, race
# statistics with the supporter-set dump, bench lines of both arms (C2), k-NN probe + ncu capture of its query kernel
mkdir -p gpurun_out
PY=python
timeout 900 $PY -m pytest tests -m gpu -q -s > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c6_pytest.log
tail -40 gpurun_out/c6_pytest.log
timeout 300 $PY -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c6_smoke.log 2>&1; tail -2 gpurun_out/c6_smoke.log
timeout 600 $PY tools/knn_probe.py --out gpurun_out/c6_knn_probe.json > gpurun_out/c6_knn_probe.log 2>&1; tail -3 gpurun_out/c6_knn_probe.log
timeout 600 $PY tools/knn_probe.py --points 4000000 --queries 4000000 --cpu-sample 100000 --out gpurun_out/c6_knn_probe_4m.json > gpurun_out/c6_knn_probe_4m.log 2>&1; tail -1 gpurun_out/c6_knn_probe_4m.log
timeout 600 $PY tools/knn_probe.py --cell-factor 0.5 --cpu-sample 1000 --out gpurun_out/c6_knn_probe_cell05.json > gpurun_out/c6_knn_probe_cell05.log 2>&1; tail -1 gpurun_out/c6_knn_probe_cell05.log
timeout 600 $PY tools/knn_probe.py --cell-factor 2.0 --cpu-sample 1000 --out gpurun_out/c6_knn_probe_cell20.json > gpurun_out/c6_knn_probe_cell20.log 2>&1; tail -1 gpurun_out/c6_knn_probe_cell20.log
timeout 900 $PY tools/ab_probe.py --lib old4=variants/lib_old4.so --config old4:lib=old4 \
  --config prio1:SM_B200_GRAPH_PRIO=1 --config prio2:SM_B200_GRAPH_PRIO=2 \
  --config offchain50:SM_B200_OFFCHAIN_GRID_PERCENT=50 --config offchain75:SM_B200_OFFCHAIN_GRID_PERCENT=75 \
  --config prio1_offchain50:SM_B200_GRAPH_PRIO=1+SM_B200_OFFCHAIN_GRID_PERCENT=50 \
  --config plain_tiebreak:SM_B200_TIEBREAK=0,0,0 --config streams_r1:SM_B200_GRAPH=0 \
  --out gpurun_out/c6_ab.json > gpurun_out/c6_ab.log 2>&1
cat gpurun_out/c6_ab.log
timeout 900 $PY tools/race_stats.py --out gpurun_out/c6_race_stats.json > gpurun_out/c6_race.log 2>&1; echo "race rc=$?" >> gpurun_out/c6_race.log
tail -45 gpurun_out/c6_race.log
timeout 600 $PY bench.py --steps 5 --warmup 3 > gpurun_out/c6_bench_product.json 2> gpurun_out/c6_bench_product.err
timeout 600 $PY bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/c6_bench_reference.json 2> gpurun_out/c6_bench_reference.err
for f in gpurun_out/c6_bench_*.json; do echo $f; $PY tools/show_bench.py $f 2>/dev/null | head -2; done
SM_B200_DELTA_TIMING=1 timeout 900 $PY tools/transfer_probe.py --width 1280 --height 960 --frames 400 --cap 20000000 --every 10 --out gpurun_out/c6_transfer_probe_hd.json > gpurun_out/c6_transfer_hd.log 2>&1
grep -v "delta transfer:" gpurun_out/c6_transfer_hd.log; grep "delta transfer:" gpurun_out/c6_transfer_hd.log | tail -3
timeout 600 $PY tools/transfer_probe.py --every 5 --out gpurun_out/c6_transfer_probe_every5.json > gpurun_out/c6_transfer_every5.log 2>&1; cat gpurun_out/c6_transfer_every5.log
for frames in 2000 500; do
timeout 900 $PY bench.py --sigma-depth 0.05 --frames $frames --cap 40000000 --required-inliers 1 --erosion-radius 0 --steps 2 --warmup 3 --no-cpu-baseline \
   > gpurun_out/c6_bench_product_C5b_$frames.json 2> gpurun_out/c6_bench_product_C5b_$frames.err
tail -1 gpurun_out/c6_bench_product_C5b_$frames.err; $PY tools/show_bench.py gpurun_out/c6_bench_product_C5b_$frames.json 2>/dev/null | head -2
done
timeout 900 $PY bench.py --impl reference --sigma-depth 0.05 --frames 500 --cap 40000000 --required-inliers 1 --erosion-radius 0 --steps 2 --warmup 3 \
   > gpurun_out/c6_bench_reference_C5b_500.json 2> gpurun_out/c6_bench_reference_C5b_500.err
tail -1 gpurun_out/c6_bench_reference_C5b_500.err; $PY tools/show_bench.py gpurun_out/c6_bench_reference_C5b_500.json 2>/dev/null | head -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn -c 12 -f \
   -o gpurun_out/c6_knn $PY tools/knn_probe.py --reps 1 --cpu-sample 1000 > gpurun_out/c6_ncu_knn.log 2>&1
tail -2 gpurun_out/c6_ncu_knn.log
