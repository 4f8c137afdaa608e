#!/bin/bash
# Round-2 final evidence run on the GPU box (one call).  Outputs: gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_gputest_full.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r2_launches_final.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --check-reads 0 --bases-per-step 8388608 > gpurun_out/r2_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rtk_k1_inexact --launch-skip 2 -c 1 -f -o gpurun_out/r2_k1_inexact \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --check-reads 0 --bases-per-step 8388608 > gpurun_out/r2_ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rtk_myers_fused_kernel|rtk_myers_fill_fused" --launch-skip 200 -c 4 -f -o gpurun_out/r2_myers \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --check-reads 0 --bases-per-step 8388608 > gpurun_out/r2_ncu_myers.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_racecheck.log 2>&1
tail -3 gpurun_out/r2_gputest_full.log; tail -c 400 gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/r2_sanitizer_memcheck.log; tail -3 gpurun_out/r2_sanitizer_racecheck.log
