#!/bin/bash
# round 2, GPU call 16: the test suite as committed (free-running envelope against six oracle runs), three times over
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c16_pytest_$i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c16_pytest_$i.log
tail -3 gpurun_out/c16_pytest_$i.log | cut -c1-300
done
