#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m pytest tests/test_models_gpu.py -m gpu -q -x -k "1024" --timeout 600 --timeout-method=thread 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 1200 gpurun_out/bench_2gpu.json; tail -2 gpurun_out/bench_2gpu.err
