#!/bin/bash
# round-2 GPU call 18: patch-resident 3x3 form (layer1 conv2 fprop / dgrad): parity, per-layer timing, ResNet tests, bench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 600 python -m pytest tests/test_gpu_conv.py -q -x 2>&1 | tail -15 ) > gpurun_out/r2c18_conv.log 2>&1
( DIRB200_TAG=patch timeout 200 python tests/cta2_check.py time "l1.c2" ) > gpurun_out/r2c18_time_patch.log 2>&1
( DIRB200_PATCH=0 DIRB200_TAG=nopatch timeout 200 python tests/cta2_check.py time "l1.c2" ) > gpurun_out/r2c18_time_nopatch.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_conv_full_size.py tests/test_gpu_conv_variants.py tests/test_gpu_resnet.py tests/test_gpu_train_loop.py -q 2>&1 | tail -15 ) > gpurun_out/r2c18_pytest.log 2>&1
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c18_bench.json 2> gpurun_out/r2c18_bench.err
cat gpurun_out/r2c18_conv.log; tail -3 gpurun_out/r2c18_time_patch.log; tail -3 gpurun_out/r2c18_time_nopatch.log; cat gpurun_out/r2c18_pytest.log
python - gpurun_out/r2c18_bench.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], d["ms_per_step"], d["e2e"]["ms_per_step"], {k:v["ms_per_step"] for k,v in d["kernel_breakdown_ms"].items()})
except Exception as e: print(sys.argv[1], "ERR", e)
PY
tail -3 gpurun_out/r2c18_bench.err
exit 0
