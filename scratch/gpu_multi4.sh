#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/pytest_multi4.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi4.log
tail -3 gpurun_out/pytest_multi4.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
python scratch/short.py n4 < gpurun_out/bench_n4.json
tail -3 gpurun_out/bench_n4.err
