#!/bin/bash
# round 2, GPU call 3: BASELINE configs C3 / C5 (both arms), ncu launch lists + full captures (C2 and C3 late frame),
# delta-transfer probe, sanitizer pass
mkdir -p gpurun_out
PY=python
# ---- C3: 1280x960, 1000 frames, 20 M cap ----
timeout 1500 $PY bench.py --width 1280 --height 960 --frames 1000 --cap 20000000 --steps 3 --warmup 3 --no-cpu-baseline \
   > gpurun_out/c3_bench_product_C3.json 2> gpurun_out/c3_bench_product_C3.err
timeout 1500 $PY bench.py --impl reference --width 1280 --height 960 --frames 1000 --cap 20000000 --steps 2 --warmup 3 \
   > gpurun_out/c3_bench_reference_C3.json 2> gpurun_out/c3_bench_reference_C3.err
# ---- C5: sigma 0.05 m, 2000 frames; and the variant with required inliers 5 ----
for v in "" "--required-inliers 5"; do
  tag=C5$(echo $v | tr -d ' -' | sed 's/requiredinliers/_req/')
  timeout 900 $PY bench.py --sigma-depth 0.05 --frames 2000 $v --steps 3 --warmup 3 --no-cpu-baseline \
     > gpurun_out/c3_bench_product_$tag.json 2> gpurun_out/c3_bench_product_$tag.err
  timeout 900 $PY bench.py --impl reference --sigma-depth 0.05 --frames 2000 $v --steps 2 --warmup 3 \
     > gpurun_out/c3_bench_reference_$tag.json 2> gpurun_out/c3_bench_reference_$tag.err
done
for f in gpurun_out/c3_bench_*.json; do echo $f; $PY tools/show_bench.py $f 2>/dev/null | head -3; done
# ---- delta transfer ----
timeout 600 $PY tools/transfer_probe.py --out gpurun_out/c3_transfer_probe.json > gpurun_out/c3_transfer.log 2>&1
cat gpurun_out/c3_transfer.log
# ---- ncu: launch lists (durations), product (frame graph) and reference, C2 stream, late frames ----
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --launch-skip 5400 -c 400 --csv \
   --log-file gpurun_out/c3_launches_product.csv $PY tools/stream_probe.py --frames 500 --reps 1 > gpurun_out/c3_ncu_product.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:Kernel --launch-skip 17500 -c 1330 --csv \
   --log-file gpurun_out/c3_launches_reference.csv $PY tools/stream_probe.py --impl reference --frames 500 --reps 1 > gpurun_out/c3_ncu_reference.log 2>&1
# ---- ncu --set full: one late frame of C2 (N ~ 0.5 M) and of C3 (N large) ----
SM_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 5520 -c 13 -f \
   -o gpurun_out/c3_frame460_C2 $PY tools/stream_probe.py --frames 500 --reps 1 > gpurun_out/c3_ncu_full_C2.log 2>&1
SM_B200_GRAPH=0 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 11520 -c 13 -f \
   -o gpurun_out/c3_frame960_C3 $PY tools/stream_probe.py --width 1280 --height 960 --frames 1000 --cap 20000000 --reps 1 > gpurun_out/c3_ncu_full_C3.log 2>&1
tail -3 gpurun_out/c3_ncu_full_C3.log
ls -la gpurun_out/*.ncu-rep
