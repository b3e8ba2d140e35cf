#!/bin/bash
# round-2 GPU call 40: vectorised split-K reduce: parity, bench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_resnet.py tests/test_gpu_train_loop.py tests/test_gpu_dense_ops.py -q 2>&1 | tail -8 ) > gpurun_out/r2c40_pytest.log 2>&1
( timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c40_bench.json 2> gpurun_out/r2c40_bench.err
cat gpurun_out/r2c40_pytest.log
python - gpurun_out/r2c40_bench.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), {k:v["ms_per_step"] for k,v in d["kernel_breakdown_ms"].items()}, d["clocks"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
exit 0
