#!/bin/bash
# 2-GPU call: MHD with the branch-free quotients (bits + speed), div B kernel, single-launch boundary shell (bits + overlap cost)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_mhd.py tests/test_gpu_divb.py tests/test_gpu_fullsize.py tests/test_gpu_decks.py -x -q 2>&1 | tail -6 | tee gpurun_out/r6_pytest.log
python bench.py --workload ot --global-size 256 --steps 10 --warmup 3 --no-e2e --no-configs --no-parity --no-cpu-baseline --single-mode 2>gpurun_out/r6_ot.err | python tools/short.py ot256_strict | tee -a gpurun_out/r6_run6.txt
python bench.py --workload ot --global-size 256 --steps 10 --warmup 3 --no-e2e --no-configs --no-parity --no-cpu-baseline --single-mode --mode fast 2>gpurun_out/r6_otf.err | python tools/short.py ot256_fast | tee -a gpurun_out/r6_run6.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "2" 2>&1 | tail -3 | tee -a gpurun_out/r6_pytest.log
tr() { port=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 "$@"; }
q="--steps 8 --warmup 3 --no-e2e --single-mode --no-configs --no-parity"
for ov in 1 0; do
  CHOLLA_OVERLAP=$ov tr 2971$ov $q 2>gpurun_out/r6_n2_ov$ov.err | python tools/short.py n2_512_ov$ov | tee -a gpurun_out/r6_run6.txt
  CHOLLA_OVERLAP=$ov tr 2972$ov $q --steps 20 --size 256 --workload kh 2>gpurun_out/r6_n2_kh256_ov$ov.err | python tools/short.py n2_kh256_ov$ov | tee -a gpurun_out/r6_run6.txt
done
