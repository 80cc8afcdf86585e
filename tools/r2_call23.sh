#!/bin/bash
# round 2, GPU call 23 (4 GPUs): replicas as the driver launches them
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 5 --warmup 3 \
   > gpurun_out/c23_bench_product_n4.json 2> gpurun_out/c23_bench_product_n4.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 4 --steps 3 --warmup 3 \
   > gpurun_out/c23_bench_reference_n4.json 2> gpurun_out/c23_bench_reference_n4.err
python - <<'PY'
import json
for f in ("gpurun_out/c23_bench_product_n4.json", "gpurun_out/c23_bench_reference_n4.json"):
    ln = [l for l in open(f) if l.startswith("{")]
    j = json.loads(ln[-1]); print(f, round(j["value"]), j["n_gpus"], round(j["e2e"]["value"]))
PY
