#!/bin/bash
# round-2 GPU call 17: whole GPU suite (dz by the reduce pass, dgrad-fused BN moments, folded-BN eval), bench, and the
# shifted-descriptor micro-experiment (tools/exp_shift.cu)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 120 tools/exp_shift.bin ) > gpurun_out/r2c17_exp_shift.log 2>&1
( time timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 ) > gpurun_out/r2c17_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c17_pytest.log
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err
cat gpurun_out/r2c17_exp_shift.log
tail -8 gpurun_out/r2c17_pytest.log; for f in gpurun_out/r2c17_bench.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], d["ms_per_step"], d["e2e"]["ms_per_step"], {k:v["ms_per_step"] for k,v in d["kernel_breakdown_ms"].items()})
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -3 gpurun_out/r2c17_bench.err
exit 0
