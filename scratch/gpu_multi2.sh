#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/pytest_multi2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi2.log
tail -3 gpurun_out/pytest_multi2.log
for m in strict fast; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --mode $m > gpurun_out/bench_n2_$m.json 2> gpurun_out/bench_n2_$m.err
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --mode $m --no-cpu-baseline > gpurun_out/bench_n1_$m.json 2> gpurun_out/bench_n1_$m.err
done
python - <<'PY'
import json
for n in (1,2):
    for m in ("strict","fast"):
        try:
            d=json.loads(open("gpurun_out/bench_n%d_%s.json"%(n,m)).read().strip().splitlines()[-1])
            print(n,m,"Gcell/s %.3f ms/step %.2f e2e %.3f"%(d["value"]/1e9,d["ms_per_step"],(d["e2e"] or {}).get("value",0)/1e9))
        except Exception as e: print(n,m,"failed",e)
PY
tail -5 gpurun_out/bench_n2_strict.err
