#!/bin/bash
# round-2 GPU call 24: batched input pipeline (augment kernel) parity, stem BN+ReLU+max-pool per 2x2 output block, bench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 900 python -m pytest tests/test_gpu_input_pipeline.py tests/test_gpu_resnet.py tests/test_gpu_train_loop.py tests/test_gpu_train_script.py -q 2>&1 | tail -12 ) > gpurun_out/r2c24_pytest.log 2>&1
( timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c24_bench.json 2> gpurun_out/r2c24_bench.err
cat gpurun_out/r2c24_pytest.log
python - gpurun_out/r2c24_bench.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), {k:v["ms_per_step"] for k,v in d["kernel_breakdown_ms"].items()})
except Exception as e: print(sys.argv[1], "ERR", e)
PY
tail -3 gpurun_out/r2c24_bench.err
exit 0
