#!/bin/bash
# round-2 GPU call 23 (2 GPUs): multi-GPU check + 2-rank bench with the patch-form / fused-moment kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/mgpu_check.py ) > gpurun_out/r2c23_mgpu.log 2>&1
echo "mgpu rc=$?" >> gpurun_out/r2c23_mgpu.log
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r2c23_bench2.json 2> gpurun_out/r2c23_bench2.err
tail -6 gpurun_out/r2c23_mgpu.log; cut -c1-400 gpurun_out/r2c23_bench2.json; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2c23_bench2.json")); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["replica_check"])
except Exception as e: print("ERR", e)
PY
tail -3 gpurun_out/r2c23_bench2.err
exit 0
