#!/bin/bash
# round-2 GPU call 38: im2col-TMA-fed parity classes of the stride-2 dgrads: parity, per-layer timing, bench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_conv_full_size.py tests/test_gpu_conv_variants.py tests/test_gpu_resnet.py -q 2>&1 | tail -8 ) > gpurun_out/r2c38_pytest.log 2>&1
( DIRB200_TAG=s2tma timeout 200 python tests/cta2_check.py time "/2" ) > gpurun_out/r2c38_time.log 2>&1
( timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c38_bench.json 2> gpurun_out/r2c38_bench.err
cat gpurun_out/r2c38_pytest.log; tail -8 gpurun_out/r2c38_time.log
python - gpurun_out/r2c38_bench.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), {k:v["ms_per_step"] for k,v in d["kernel_breakdown_ms"].items()})
except Exception as e: print(sys.argv[1], "ERR", e)
PY
exit 0
