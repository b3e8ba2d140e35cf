#!/bin/bash
# round-2 GPU call 22: patch form: two epilogue groups, patch-resident wgrad: parity, l1.c2 timing, A/B bench (same box)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_conv_full_size.py tests/test_gpu_conv_variants.py tests/test_gpu_resnet.py tests/test_gpu_train_loop.py -q 2>&1 | tail -12 ) > gpurun_out/r2c22_pytest.log 2>&1
( DIRB200_TAG=patch4 timeout 200 python tests/cta2_check.py time "l1.c2" ) > gpurun_out/r2c22_time.log 2>&1
for v in on off on2; do
  case $v in on|on2) E="X=1";; off|off2) E="DIRB200_PATCH=0";; esac
  ( env $E timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c22_bench_$v.json 2> gpurun_out/r2c22_bench_$v.err
done
cat gpurun_out/r2c22_pytest.log; tail -2 gpurun_out/r2c22_time.log
for v in on off on2; do python - gpurun_out/r2c22_bench_$v.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), {k:v["ms_per_step"] for k,v in d["kernel_breakdown_ms"].items()})
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
exit 0
