#!/bin/bash
# round 2, GPU call 18: final defaults (1 % / 1.5 % / 3 % early secondaries): the test suite three times (fresh oracle
# draws each), bench lines of both arms, free-running totals
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c18_pytest_$i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c18_pytest_$i.log
tail -3 gpurun_out/c18_pytest_$i.log | cut -c1-300
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c18_smoke.log 2>&1; tail -1 gpurun_out/c18_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c18_bench_product.json 2> gpurun_out/c18_bench_product.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/c18_bench_reference.json 2> gpurun_out/c18_bench_reference.err
timeout 1500 python bench.py --width 1280 --height 960 --frames 1000 --cap 20000000 --steps 3 --warmup 3 --no-cpu-baseline \
   > gpurun_out/c18_bench_product_C3.json 2> gpurun_out/c18_bench_product_C3.err
timeout 1500 python bench.py --impl reference --width 1280 --height 960 --frames 1000 --cap 20000000 --steps 2 --warmup 3 \
   > gpurun_out/c18_bench_reference_C3.json 2> gpurun_out/c18_bench_reference_C3.err
for f in gpurun_out/c18_bench_*.json; do echo $f; python tools/show_bench.py $f 2>/dev/null | head -1; done
timeout 600 python tools/free_running_check.py --frames 500 --oracle-runs 4 --rule default --out gpurun_out/c18_free_vga500.json > gpurun_out/c18_free_vga500.log 2>&1; cat gpurun_out/c18_free_vga500.log | cut -c1-200
timeout 600 python tools/free_running_check.py --frames 1000 --rule default --out gpurun_out/c18_free_vga1000.json > gpurun_out/c18_free_vga1000.log 2>&1; cat gpurun_out/c18_free_vga1000.log | cut -c1-200
timeout 900 python tools/free_running_check.py --width 1280 --height 960 --frames 1000 --cap 20000000 --rule default --out gpurun_out/c18_free_hd1000.json > gpurun_out/c18_free_hd1000.log 2>&1; cat gpurun_out/c18_free_hd1000.log | cut -c1-200
