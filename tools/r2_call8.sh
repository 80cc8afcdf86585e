#!/bin/bash
# round 2, GPU call 8: test suite with the 2x envelope, the compacting k_update_neighbors variants (parity + A/B)
mkdir -p gpurun_out
PY=python
timeout 900 $PY -m pytest tests -m gpu -q -s > gpurun_out/c8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c8_pytest.log
grep -n "passed\|failed\|^FAILED\|^E  " gpurun_out/c8_pytest.log | cut -c1-300 | tail -30
SM_B200_LIB=variants/lib_compact3.so timeout 900 $PY -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -m gpu -q > gpurun_out/c8_pytest_compact3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c8_pytest_compact3.log
tail -15 gpurun_out/c8_pytest_compact3.log | cut -c1-300
timeout 900 $PY tools/ab_probe.py --lib compact3=variants/lib_compact3.so --lib compact2=variants/lib_compact2.so \
  --config compact3:lib=compact3 --config compact2:lib=compact2 \
  --out gpurun_out/c8_ab.json > gpurun_out/c8_ab.log 2>&1
cat gpurun_out/c8_ab.log
timeout 900 $PY tools/ab_probe.py --width 1280 --height 960 --frames 400 --cap 20000000 --reps 3 \
  --lib compact3=variants/lib_compact3.so --lib compact2=variants/lib_compact2.so \
  --config compact3:lib=compact3 --config compact2:lib=compact2 \
  --out gpurun_out/c8_ab_hd.json > gpurun_out/c8_ab_hd.log 2>&1
cat gpurun_out/c8_ab_hd.log
SM_B200_LIB=variants/lib_compact3.so timeout 600 $PY bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c8_bench_compact3.json 2> gpurun_out/c8_bench_compact3.err
$PY tools/show_bench.py gpurun_out/c8_bench_compact3.json | head -16
