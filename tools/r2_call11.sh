#!/bin/bash
# round 2, GPU call 11 (2 GPUs): the replica path of bench.py as the driver launches it
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 \
   > gpurun_out/c11_bench_product_n2.json 2> gpurun_out/c11_bench_product_n2.err
python tools/show_bench.py gpurun_out/c11_bench_product_n2.json | head -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 \
   > gpurun_out/c11_bench_reference_n2.json 2> gpurun_out/c11_bench_reference_n2.err
python tools/show_bench.py gpurun_out/c11_bench_reference_n2.json | head -1
