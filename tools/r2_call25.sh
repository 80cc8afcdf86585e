#!/bin/bash
# round 2, GPU call 21: HEAD as committed: test suite, smoke, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c25_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c25_pytest.log; tail -3 gpurun_out/c25_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c25_smoke.log 2>&1; tail -1 gpurun_out/c25_smoke.log
timeout 600 python bench.py > gpurun_out/c25_bench_product.json 2> gpurun_out/c25_bench_product.err; python tools/show_bench.py gpurun_out/c25_bench_product.json | head -1
timeout 600 python bench.py --impl reference > gpurun_out/c25_bench_reference.json 2> gpurun_out/c25_bench_reference.err; python tools/show_bench.py gpurun_out/c25_bench_reference.json | head -1
