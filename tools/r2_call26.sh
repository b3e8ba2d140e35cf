#!/bin/bash
# round-2 GPU call 26: CUDA graphs on by default: whole GPU suite, smoke, bench (with and without graphs)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2c26_pytest.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2c26_smoke.log 2>&1
for v in on off; do
  case $v in on) E="X=1";; off) E="DIRB200_GRAPH=0";; esac
  ( env $E timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c26_bench_$v.json 2> gpurun_out/r2c26_bench_$v.err
done
cat gpurun_out/r2c26_pytest.log; tail -2 gpurun_out/r2c26_smoke.log
for v in on off; do python - gpurun_out/r2c26_bench_$v.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["gpu_launches_per_step"], round(d["wall_ms_per_step"],3), d["roofline"]["frac"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
exit 0
