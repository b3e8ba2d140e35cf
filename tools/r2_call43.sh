#!/bin/bash
# round-2 GPU call 43: packed (value, position) keys in the stem's BN+ReLU+max-pool; whole GPU suite + smoke + bench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r2c43_pytest.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2c43_smoke.log 2>&1
( timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c43_bench.json 2> gpurun_out/r2c43_bench.err
cat gpurun_out/r2c43_pytest.log; tail -1 gpurun_out/r2c43_smoke.log
python - gpurun_out/r2c43_bench.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), {k:v["ms_per_step"] for k,v in d["kernel_breakdown_ms"].items()}, d["clocks"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
exit 0
