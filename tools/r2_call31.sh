#!/bin/bash
# round-2 GPU call 31: 5x5 convolutions, dense-prediction ops (upsample / concat / autograd wrappers), config-4 figures
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 900 python -m pytest tests/test_gpu_dense_ops.py tests/test_gpu_conv.py tests/test_gpu_resnet.py tests/test_gpu_train_loop.py -q 2>&1 | tail -15 ) > gpurun_out/r2c31_pytest.log 2>&1

( timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c31_bench.json 2> gpurun_out/r2c31_bench.err
cat gpurun_out/r2c31_pytest.log; 
python - gpurun_out/r2c31_bench.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["gpu_launches_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
exit 0
