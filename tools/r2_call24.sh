#!/bin/bash
# round 2, GPU call 24 (4 GPUs): where does the replica spread come from? graph vs stream pipeline, with / without the
# nvidia-smi sampler; per-rank step times
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
     bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/c24_$name.json 2> gpurun_out/c24_$name.err
  python - "$name" <<'PY'
import json, sys
ln = [l for l in open(f"gpurun_out/c24_{sys.argv[1]}.json") if l.startswith("{")]
j = json.loads(ln[-1]); print(sys.argv[1], round(j["value"]), "per rank ms", [round(v, 2) for v in j.get("ms_per_step_per_rank", [])])
PY
}
run graph A=1
run graph_noclocks SM_BENCH_NO_CLOCKS=1
run streams SM_B200_GRAPH=0
run graph_again A=1
nproc; uptime
