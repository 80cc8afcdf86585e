#!/bin/bash
# frame graph with ordering-only edges (SM_B200_GRAPH_ORDER bit mask): same-box A/B at VGA and 1280x960
mkdir -p gpurun_out
PY=python
timeout 900 $PY tools/ab_probe.py --lib order=variants/lib_order.so --config base:lib=order \
  --config o1:lib=order+SM_B200_GRAPH_ORDER=1 --config o2:lib=order+SM_B200_GRAPH_ORDER=2 --config o3:lib=order+SM_B200_GRAPH_ORDER=3 \
  --config o7:lib=order+SM_B200_GRAPH_ORDER=7 --config o8:lib=order+SM_B200_GRAPH_ORDER=8 --config o10:lib=order+SM_B200_GRAPH_ORDER=10 \
  --config o11:lib=order+SM_B200_GRAPH_ORDER=11 --config o15:lib=order+SM_B200_GRAPH_ORDER=15 \
  --config o3_prio1:lib=order+SM_B200_GRAPH_ORDER=3+SM_B200_GRAPH_PRIO=1 \
  --out gpurun_out/c8c_ab.json > gpurun_out/c8c_ab.log 2>&1
cat gpurun_out/c8c_ab.log
timeout 900 $PY tools/ab_probe.py --width 1280 --height 960 --frames 400 --cap 20000000 --reps 3 --lib order=variants/lib_order.so \
  --config base:lib=order --config o3:lib=order+SM_B200_GRAPH_ORDER=3 --config o7:lib=order+SM_B200_GRAPH_ORDER=7 \
  --config o11:lib=order+SM_B200_GRAPH_ORDER=11 \
  --out gpurun_out/c8c_ab_hd.json > gpurun_out/c8c_ab_hd.log 2>&1
cat gpurun_out/c8c_ab_hd.log
