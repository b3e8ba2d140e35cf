#!/bin/bash
# round-2 GPU call 37 (2 GPUs): faster all-reduce -- NCCL channel counts with the blocking / overlapped gradient reduction
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for v in blocking blocking_ch32 blocking_ch16 overlap_ch32; do
  case $v in blocking) E="DIRB200_OVERLAP_ALLREDUCE=0";; blocking_ch32) E="DIRB200_OVERLAP_ALLREDUCE=0 NCCL_MIN_NCHANNELS=32";; blocking_ch16) E="DIRB200_OVERLAP_ALLREDUCE=0 NCCL_MIN_NCHANNELS=16";; overlap_ch32) E="NCCL_MIN_NCHANNELS=32";; esac
  ( env $E timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 30 --warmup 5 ) > gpurun_out/r2c37_bench2_$v.json 2> gpurun_out/r2c37_bench2_$v.err
done
( CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2c37_bench1.json 2> gpurun_out/r2c37_bench1.err
for f in gpurun_out/r2c37_bench1.json gpurun_out/r2c37_bench2_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
exit 0
