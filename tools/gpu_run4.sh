#!/bin/bash
# 1-GPU call: rolled-loop PPM variants (bits + speed), then the full default bench line (parity + configs), drop-in and
# nosync-reference numbers
mkdir -p gpurun_out
run() { tag=$1; shift
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --single-mode --no-configs --no-parity "$@" 2>gpurun_out/r4_bench_$tag.err | tee gpurun_out/r4_bench_$tag.json | python tools/short.py $tag | tee -a gpurun_out/r4_run4.txt; }
for v in loop7 loop6; do
  export CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_$v.so
  timeout 600 python -m pytest "tests/test_gpu_parity.py::test_strict_mode_is_bit_exact_3d[2-0]" tests/test_gpu_reference.py::test_reference_kernels_blast_mixed_boundaries "tests/test_gpu_parity.py::test_strict_ieee_redo_path_extreme_operands[2-0]" -x -q 2>&1 | tail -3 | tee -a gpurun_out/r4_run4.txt
  run $v --mode strict
done
export CHOLLA_B200_LIB=$PWD/cholla_b200/libcholla_b200_loop8f.so
run loop8f_fast --mode fast
unset CHOLLA_B200_LIB
run default --mode strict
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/r4_bench_full.json 2> gpurun_out/r4_bench_full.err
tail -3 gpurun_out/r4_bench_full.err | tee -a gpurun_out/r4_run4.txt
python bench.py --impl dropin --steps 10 --warmup 3 > gpurun_out/r4_bench_dropin.json 2> gpurun_out/r4_bench_dropin.err
python bench.py --impl reference --ref-variant nosync --steps 10 --warmup 3 > gpurun_out/r4_bench_ref_nosync.json 2> gpurun_out/r4_bench_ref_nosync.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r4_bench_ref.json 2> gpurun_out/r4_bench_ref.err
python - <<'PY' | tee -a gpurun_out/r4_run4.txt
import json
for f in ("dropin", "ref_nosync", "ref"):
    try:
        d = json.loads(open("gpurun_out/r4_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), d.get("checksum"), d.get("unavailable"))
    except Exception as e:
        print(f, "ERR", e)
PY
