#!/bin/bash
# 8-GPU call: rank-count independence at 8 ranks (hydro with / without overlap, MHD), overlap on / off timings at
# 512^3 and 256^3 per GPU, then the full bench line with the faster setting
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r8_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "test_rank_count_independence[8] or test_rank_count_independence_mhd[8] or test_rank_count_independence_without_overlap[8]" 2>&1 | tail -8 | tee gpurun_out/r8_pytest_multi8.log
tr() { port=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 "$@"; }
q="--steps 6 --warmup 3 --no-e2e --single-mode --no-configs --no-parity"
for ov in 1 0; do
  CHOLLA_OVERLAP=$ov tr 2981$ov $q 2>gpurun_out/r8_n8_ov$ov.err | tee gpurun_out/r8_n8_ov$ov.json | python tools/short.py n8_512_ov$ov | tee -a gpurun_out/r8_run8.txt
  CHOLLA_OVERLAP=$ov tr 2982$ov $q --steps 20 --size 256 --workload kh 2>gpurun_out/r8_n8_kh256_ov$ov.err | tee gpurun_out/r8_n8_kh256_ov$ov.json | python tools/short.py n8_kh256_ov$ov | tee -a gpurun_out/r8_run8.txt
done
best=$(python - <<'PY'
import json
def ms(f):
    try:
        return json.loads([l for l in open(f) if l.startswith("{")][-1])["ms_per_step"]
    except Exception:
        return 1e9
print(1 if ms("gpurun_out/r8_n8_ov1.json") <= ms("gpurun_out/r8_n8_ov0.json") else 0)
PY
)
echo "overlap for the full line: $best" | tee -a gpurun_out/r8_run8.txt
( time CHOLLA_OVERLAP=$best tr 29830 --steps 10 --warmup 3 ) > gpurun_out/r8_bench_full_n8.json 2> gpurun_out/r8_bench_full_n8.err
tail -4 gpurun_out/r8_bench_full_n8.err
