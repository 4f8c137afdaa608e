#!/bin/bash
# Round-2 final evidence, second part (after the kernel fixes): launch list, K1 capture, the bench line.  Outputs: gpurun_out/.
mkdir -p gpurun_out
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/r2_launches_final.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --check-reads 0 --bases-per-step 4194304 > gpurun_out/r2_ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rtk_k1_inexact --launch-skip 2 -c 1 -f -o gpurun_out/r2_k1_inexact \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --check-reads 0 --bases-per-step 8388608 > gpurun_out/r2_ncu_k1.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -c 300 gpurun_out/r2_bench_n1.json; tail -2 gpurun_out/r2_bench_n1.err
