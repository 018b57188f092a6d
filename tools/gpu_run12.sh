#!/bin/bash
# CTA x-extent chosen per launch (hydro predictor, MHD y/z sweeps) + MHD predictor register budgets
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mhd.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4 | tee gpurun_out/r12_pytest.log
q="--steps 6 --warmup 3 --no-e2e --single-mode --no-configs --no-parity --no-cpu-baseline"
python bench.py $q 2>gpurun_out/r12_bench_strict.err | tee gpurun_out/r12_bench_strict.json | python tools/short.py r12_strict | tee -a gpurun_out/r12_run12.txt
python bench.py $q --size 256 --workload kh --steps 20 2>gpurun_out/r12_kh256.err | python tools/short.py r12_kh256 | tee -a gpurun_out/r12_run12.txt
m="--workload ot --global-size 256 --steps 10 --warmup 3 --no-e2e --no-configs --no-parity --no-cpu-baseline --single-mode"
python bench.py $m 2>gpurun_out/r12_ot.err | python tools/short.py ot256_default | tee -a gpurun_out/r12_run12.txt
for v in mhdA mhdB mhdC; do
  CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_$v.so python bench.py $m 2>gpurun_out/r12_ot_$v.err | python tools/short.py ot256_$v | tee -a gpurun_out/r12_run12.txt
done
