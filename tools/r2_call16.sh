#!/bin/bash
# round-2 GPU call 16: dz stored by the block-output reduce, BN-backward moments in the dgrad epilogue, folded-BN eval
# forward: ResNet / train-loop / conv tests, bench with the A/B switches
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1200 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_train_loop.py tests/test_gpu_train_script.py tests/test_gpu_conv.py -q 2>&1 ) > gpurun_out/r2c16_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c16_pytest.log
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c16_bench.json 2> gpurun_out/r2c16_bench.err
tail -15 gpurun_out/r2c16_pytest.log; for f in gpurun_out/r2c16_bench.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], d["ms_per_step"], d["e2e"]["ms_per_step"], {k:v["ms_per_step"] for k,v in d["kernel_breakdown_ms"].items()})
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -3 gpurun_out/r2c16_bench.err
exit 0
