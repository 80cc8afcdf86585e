#!/bin/bash
# round 2, GPU call 22: k_update_neighbors with the "some candidate is new" gate ahead of the other gates' gathers
mkdir -p gpurun_out
SM_B200_LIB=variants/lib_anynew.so timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -m gpu -q > gpurun_out/c22_pytest_anynew.log 2>&1; echo "rc=$?" >> gpurun_out/c22_pytest_anynew.log; tail -3 gpurun_out/c22_pytest_anynew.log | cut -c1-300
timeout 900 python tools/ab_probe.py --lib anynew=variants/lib_anynew.so --lib head=variants/lib_head.so \
  --config head:lib=head --config anynew:lib=anynew --config head2:lib=head --config anynew2:lib=anynew \
  --out gpurun_out/c22_ab.json > gpurun_out/c22_ab.log 2>&1
cat gpurun_out/c22_ab.log
timeout 900 python tools/ab_probe.py --width 1280 --height 960 --frames 400 --cap 20000000 --reps 3 --lib anynew=variants/lib_anynew.so --lib head=variants/lib_head.so \
  --config head:lib=head --config anynew:lib=anynew \
  --out gpurun_out/c22_ab_hd.json > gpurun_out/c22_ab_hd.log 2>&1
cat gpurun_out/c22_ab_hd.log
