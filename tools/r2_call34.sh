#!/bin/bash
# round-2 GPU call 34: evidence of the final build -- whole GPU suite, bench (with the CPU arm), reference arm, ncu launch
# list + --set full of the new kernels, compute-sanitizer, config-4 figures
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r2c34_pytest.log 2>&1
( timeout 600 python bench.py --steps 30 --warmup 5 ) > gpurun_out/r2c34_bench.json 2> gpurun_out/r2c34_bench.err
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2c34_bench_ref.json 2> gpurun_out/r2c34_bench_ref.err
( timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_r2g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2c34_ncu.log 2>&1
prof() {  # name, layer substring, form
  ( timeout 300 ncu --set full --clock-control none --import-source on -k regex:"igemm_kernel|wgrad_patch_kernel" -s 2 -c 1 -f \
      -o gpurun_out/r2g_$1 python tests/cta2_check.py one "$2" $3 ) > gpurun_out/r2c34_ncu_$1.log 2>&1
}
prof l1c2_fprop "l1.c2 " fprop
prof l1c2_wgrad "l1.c2 " wgrad
prof l3c2_dgrad "l3.c2 " dgrad
( timeout 300 python tools/nyud2_bench.py ) > gpurun_out/r2c34_nyud2.json 2> gpurun_out/r2c34_nyud2.err
bash tools/r2_sanitizer.sh > gpurun_out/r2c34_sanitizer_summary.log 2>&1
cat gpurun_out/r2c34_pytest.log; cut -c1-300 gpurun_out/r2c34_bench.json; cut -c1-400 gpurun_out/r2c34_bench_ref.json; tail -2 gpurun_out/r2c34_ncu.log | cut -c1-200
cat gpurun_out/r2c34_nyud2.json; tail -2 gpurun_out/r2c34_nyud2.err; cat gpurun_out/r2c34_sanitizer_summary.log | tail -14
ls -la gpurun_out/*.ncu-rep | tail -4
exit 0
