#!/bin/bash
# round 2, GPU call 9 (final state of the code): test suite, smoke, bench lines C2 / C3 of both arms, ncu launch
# lists + full captures (C2, C3, k-NN), k-NN probes (sheet, random cloud = BASELINE config 1 pattern),
# compute-sanitizer on the new kernels
mkdir -p gpurun_out
PY=python
timeout 900 $PY -m pytest tests -m gpu -q > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
tail -4 gpurun_out/c9_pytest.log | cut -c1-300
timeout 300 $PY -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c9_smoke.log 2>&1; tail -1 gpurun_out/c9_smoke.log
timeout 600 $PY bench.py --steps 5 --warmup 3 > gpurun_out/c9_bench_product.json 2> gpurun_out/c9_bench_product.err
timeout 600 $PY bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/c9_bench_reference.json 2> gpurun_out/c9_bench_reference.err
timeout 1500 $PY bench.py --width 1280 --height 960 --frames 1000 --cap 20000000 --steps 3 --warmup 3 --no-cpu-baseline \
   > gpurun_out/c9_bench_product_C3.json 2> gpurun_out/c9_bench_product_C3.err
timeout 1500 $PY bench.py --impl reference --width 1280 --height 960 --frames 1000 --cap 20000000 --steps 2 --warmup 3 \
   > gpurun_out/c9_bench_reference_C3.json 2> gpurun_out/c9_bench_reference_C3.err
for f in gpurun_out/c9_bench_*.json; do echo $f; $PY tools/show_bench.py $f 2>/dev/null | head -1; done
timeout 600 $PY tools/knn_probe.py --out gpurun_out/c9_knn_probe.json > gpurun_out/c9_knn_probe.log 2>&1; tail -1 gpurun_out/c9_knn_probe.log | cut -c1-400
timeout 600 $PY tools/knn_probe.py --cloud random --points 10000 --queries 10000 --spacing 0.05 --cpu-sample 10000 --out gpurun_out/c9_knn_probe_c1.json > gpurun_out/c9_knn_probe_c1.log 2>&1; tail -1 gpurun_out/c9_knn_probe_c1.log | cut -c1-400
timeout 600 $PY tools/knn_probe.py --cloud random --points 1000000 --queries 1000000 --spacing 0.01 --cpu-sample 100000 --out gpurun_out/c9_knn_probe_random1m.json > gpurun_out/c9_knn_probe_random1m.log 2>&1; tail -1 gpurun_out/c9_knn_probe_random1m.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --launch-skip 5280 -c 400 --csv \
   --log-file gpurun_out/c9_launches_product.csv $PY bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/c9_ncu_product.log 2>&1
SM_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 5520 -c 13 -f \
   -o gpurun_out/c9_frame460_C2 $PY tools/stream_probe.py --frames 500 --reps 1 > gpurun_out/c9_ncu_full_C2.log 2>&1
SM_B200_GRAPH=0 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 11520 -c 13 -f \
   -o gpurun_out/c9_frame960_C3 $PY tools/stream_probe.py --width 1280 --height 960 --frames 1000 --cap 20000000 --reps 1 > gpurun_out/c9_ncu_full_C3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn -c 12 -f \
   -o gpurun_out/c9_knn $PY tools/knn_probe.py --reps 1 --cpu-sample 1000 > gpurun_out/c9_ncu_knn.log 2>&1
timeout 900 compute-sanitizer --tool memcheck $PY -m pytest tests/test_knn_gpu.py tests/test_round2_gpu.py -m gpu -q \
   -k "knn or delta_transfer or median" > gpurun_out/c9_memcheck.log 2>&1
tail -4 gpurun_out/c9_memcheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all $PY -m pytest tests/test_knn_gpu.py -m gpu -q -k "golden" > gpurun_out/c9_racecheck.log 2>&1
tail -4 gpurun_out/c9_racecheck.log | cut -c1-200
