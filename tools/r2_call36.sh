#!/bin/bash
# round-2 GPU call 36 (2 GPUs): where does the N=2 overhead come from -- overlapped vs blocking all-reduce, conv grids
# that leave SMs to NCCL
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c36_bench1.json 2> gpurun_out/r2c36_bench1.err
for v in default blocking sms140 sms132; do
  case $v in default) E="X=1";; blocking) E="DIRB200_OVERLAP_ALLREDUCE=0";; sms140) E="DIRB200_SMS=140";; sms132) E="DIRB200_SMS=132";; esac
  ( env $E timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 30 --warmup 5 ) > gpurun_out/r2c36_bench2_$v.json 2> gpurun_out/r2c36_bench2_$v.err
done
for f in gpurun_out/r2c36_bench1.json gpurun_out/r2c36_bench2_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
exit 0
