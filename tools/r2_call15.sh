#!/bin/bash
# round 2, GPU call 15: the final defaults (separate early / slot-order fractions for the later waves): test suite, smoke,
# bench lines of both arms (C2, C3), free-running totals at the three stream sizes
mkdir -p gpurun_out
PY=python
timeout 900 $PY -m pytest tests -m gpu -q -s > gpurun_out/c15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c15_pytest.log
grep -n "passed\|failed\|^FAILED\|^E  \|product-vs-A" gpurun_out/c15_pytest.log | cut -c1-400 | tail -12
timeout 300 $PY -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c15_smoke.log 2>&1; tail -1 gpurun_out/c15_smoke.log
timeout 600 $PY bench.py --steps 5 --warmup 3 > gpurun_out/c15_bench_product.json 2> gpurun_out/c15_bench_product.err
timeout 600 $PY bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/c15_bench_reference.json 2> gpurun_out/c15_bench_reference.err
timeout 1500 $PY bench.py --width 1280 --height 960 --frames 1000 --cap 20000000 --steps 3 --warmup 3 --no-cpu-baseline \
   > gpurun_out/c15_bench_product_C3.json 2> gpurun_out/c15_bench_product_C3.err
timeout 1500 $PY bench.py --impl reference --width 1280 --height 960 --frames 1000 --cap 20000000 --steps 2 --warmup 3 \
   > gpurun_out/c15_bench_reference_C3.json 2> gpurun_out/c15_bench_reference_C3.err
for f in gpurun_out/c15_bench_*.json; do echo $f; $PY tools/show_bench.py $f 2>/dev/null | head -1; done
timeout 600 $PY tools/free_running_check.py --frames 500 --rule default --out gpurun_out/c15_free_vga500.json > gpurun_out/c15_free_vga500.log 2>&1; cat gpurun_out/c15_free_vga500.log | cut -c1-200
timeout 600 $PY tools/free_running_check.py --frames 1000 --rule default --out gpurun_out/c15_free_vga1000.json > gpurun_out/c15_free_vga1000.log 2>&1; cat gpurun_out/c15_free_vga1000.log | cut -c1-200
timeout 900 $PY tools/free_running_check.py --width 1280 --height 960 --frames 1000 --cap 20000000 --rule default --out gpurun_out/c15_free_hd1000.json > gpurun_out/c15_free_hd1000.log 2>&1; cat gpurun_out/c15_free_hd1000.log | cut -c1-200
