#!/bin/bash
# round-2 GPU call 25: CUDA-graph replay of the forward body / backward stages (DIRB200_GRAPH=1): parity + A/B bench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( DIRB200_GRAPH=1 timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_train_loop.py tests/test_gpu_train_script.py -q 2>&1 | tail -12 ) > gpurun_out/r2c25_pytest.log 2>&1
for v in off on off2 on2; do
  case $v in on|on2) E="DIRB200_GRAPH=1";; off|off2) E="X=1";; esac
  ( env $E timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c25_bench_$v.json 2> gpurun_out/r2c25_bench_$v.err
done
cat gpurun_out/r2c25_pytest.log
for v in off on off2 on2; do python - gpurun_out/r2c25_bench_$v.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["gpu_launches_per_step"], round(d["wall_ms_per_step"],3))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -3 gpurun_out/r2c25_bench_on.err
exit 0
