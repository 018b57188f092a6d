#!/bin/bash
# 2-GPU call: native halo exchange (CUDA IPC peer stores) -- rank-count independence, then weak scaling N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "2" 2>&1 | tail -25 > gpurun_out/r2_pytest_multi2.log
tail -5 gpurun_out/r2_pytest_multi2.log
for ov in 1 0; do
CHOLLA_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e --single-mode 2>gpurun_out/r2_bench_n2_ov$ov.err | tee gpurun_out/r2_bench_n2_ov$ov.json | python tools/short.py n2_ov$ov | tee -a gpurun_out/r2_run2.txt
done
CHOLLA_TRANSPORT=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e --single-mode 2>gpurun_out/r2_bench_n2_nccl.err | tee gpurun_out/r2_bench_n2_nccl.json | python tools/short.py n2_nccl | tee -a gpurun_out/r2_run2.txt
CHOLLA_OVERLAP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --single-mode --size 256 --workload kh 2>gpurun_out/r2_bench_n2_kh256.err | tee gpurun_out/r2_bench_n2_kh256.json | python tools/short.py n2_kh256 | tee -a gpurun_out/r2_run2.txt
tail -3 gpurun_out/r2_bench_n2_ov1.err
