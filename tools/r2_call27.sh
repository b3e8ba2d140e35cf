#!/bin/bash
# round-2 GPU call 27: wgrad side-stream overlap A/B (parity first), test durations
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( DIRB200_WGRAD_OVERLAP=1 timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_train_loop.py -q 2>&1 | tail -8 ) > gpurun_out/r2c27_pytest_ovl.log 2>&1
for v in off on on_nograph; do
  case $v in on) E="DIRB200_WGRAD_OVERLAP=1";; off) E="X=1";; on_nograph) E="DIRB200_WGRAD_OVERLAP=1 DIRB200_GRAPH=0";; esac
  ( env $E timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c27_bench_$v.json 2> gpurun_out/r2c27_bench_$v.err
done
( timeout 900 python -m pytest tests/ -q -m gpu --durations=12 2>&1 | tail -22 ) > gpurun_out/r2c27_durations.log 2>&1
cat gpurun_out/r2c27_pytest_ovl.log
for v in off on on_nograph; do python - gpurun_out/r2c27_bench_$v.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["gpu_launches_per_step"], round(d["wall_ms_per_step"],3))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -3 gpurun_out/r2c27_bench_on.err
cat gpurun_out/r2c27_durations.log
exit 0
