#!/bin/bash
# 8-GPU call: rank-count independence at 8 ranks (hydro overlapped, MHD, dual energy + scalar + gravity), then the bench
# line at N = 8 with end-to-end copies (ranks bound to their GPU's NUMA node)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_multi.py -q -k "test_rank_count_independence[8] or test_rank_count_independence_mhd[8] or test_rank_count_independence_dual_energy_scalar_static_gravity[8]" 2>&1 | tail -10 ) 2>&1 | tee gpurun_out/r21_pytest_multi8.log
tr() { port=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 "$@"; }
( time tr 29840 --steps 10 --warmup 3 --no-configs --no-parity --single-mode ) > gpurun_out/r21_bench_n8.json 2> gpurun_out/r21_bench_n8.err
tail -3 gpurun_out/r21_bench_n8.err
python tools/short.py n8 < gpurun_out/r21_bench_n8.json | tee -a gpurun_out/r21_run21.txt
