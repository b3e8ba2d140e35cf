#!/bin/bash
# round-2 GPU call 42 (2 GPUs): SMs reserved for the NCCL kernel beside the first block of each backward stage
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/mgpu_check.py ) > gpurun_out/r2c42_mgpu.log 2>&1
( CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c42_bench1.json 2> gpurun_out/r2c42_bench1.err
for v in default r8 r16 r8all; do
  case $v in default) E="X=1";; r8) E="DIRB200_SM_RESERVE=8 NCCL_MAX_CTAS=8";; r16) E="DIRB200_SM_RESERVE=16 NCCL_MAX_CTAS=16";; r8all) E="DIRB200_SM_RESERVE=8 NCCL_MAX_CTAS=8 NCCL_MIN_CTAS=8";; esac
  ( env $E timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 30 --warmup 5 ) > gpurun_out/r2c42_bench2_$v.json 2> gpurun_out/r2c42_bench2_$v.err
done
tail -2 gpurun_out/r2c42_mgpu.log
for f in gpurun_out/r2c42_bench1.json gpurun_out/r2c42_bench2_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d.get("replica_check",{}) and d["replica_check"]["max_abs_param_diff_vs_rank0"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
exit 0
