#!/bin/bash
# 2-GPU call: overlap on / off after the thin-box chunk fix (512^3 and 256^3 local), then the full bench line at N = 2
mkdir -p gpurun_out
tr() { port=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 "$@"; }
q="--steps 8 --warmup 3 --no-e2e --single-mode --no-configs --no-parity"
for ov in 1 0; do
  CHOLLA_OVERLAP=$ov tr 2961$ov $q 2>gpurun_out/r5_n2_ov$ov.err | tee gpurun_out/r5_n2_ov$ov.json | python tools/short.py n2_512_ov$ov | tee -a gpurun_out/r5_run5.txt
  CHOLLA_OVERLAP=$ov tr 2962$ov $q --steps 20 --size 256 --workload kh 2>gpurun_out/r5_n2_kh256_ov$ov.err | tee gpurun_out/r5_n2_kh256_ov$ov.json | python tools/short.py n2_kh256_ov$ov | tee -a gpurun_out/r5_run5.txt
done
python bench.py --steps 20 --warmup 3 --no-e2e --single-mode --no-configs --no-parity --no-cpu-baseline --size 256 --workload kh 2>/dev/null | python tools/short.py n1_kh256 | tee -a gpurun_out/r5_run5.txt
( time tr 29630 --steps 10 --warmup 3 ) > gpurun_out/r5_bench_full_n2.json 2> gpurun_out/r5_bench_full_n2.err
tail -4 gpurun_out/r5_bench_full_n2.err
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "2" 2>&1 | tail -3
