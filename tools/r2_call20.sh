#!/bin/bash
# round-2 GPU call 20: TMA-fed kernels with 192 threads (255-register cap), unrolled patch MMA issue, patch form: parity, bench, ncu --set full
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_resnet.py -q 2>&1 | tail -5 ) > gpurun_out/r2c20_pytest.log 2>&1
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c20_bench.json 2> gpurun_out/r2c20_bench.err
( DIRB200_TAG=patch200 timeout 200 python tests/cta2_check.py time "l1.c" ) > gpurun_out/r2c20_time.log 2>&1
prof() {  # name, layer substring, form
  ( timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 2 -c 1 -f \
      -o gpurun_out/r2q_$1 python tests/cta2_check.py one "$2" $3 ) > gpurun_out/r2c20_ncu_$1.log 2>&1
}
prof l1c2_fprop_patch "l1.c2 " fprop
cat gpurun_out/r2c20_pytest.log; tail -5 gpurun_out/r2c20_time.log
python - gpurun_out/r2c20_bench.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], d["ms_per_step"], d["e2e"]["ms_per_step"], {k:v["ms_per_step"] for k,v in d["kernel_breakdown_ms"].items()})
except Exception as e: print(sys.argv[1], "ERR", e)
PY
tail -3 gpurun_out/r2c20_bench.err
exit 0
