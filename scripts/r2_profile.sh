#!/bin/bash
# Round-2 evidence run on the GPU box (one call): GPU test-suite, stage breakdown of a bench step, ncu launch list, one
# `ncu --set full` capture of the region-engine kernel, compute-sanitizer memcheck of the smoke path.  Outputs: gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_gputest.log
RTK_BROKER_PROFILE=1 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --check-reads 0 --bases-per-step 8388608 > gpurun_out/r2_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rtk_region_kernel --launch-skip 3 -c 2 -f -o gpurun_out/r2_region \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --check-reads 0 --bases-per-step 8388608 > gpurun_out/r2_ncu_region.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck.log 2>&1
tail -3 gpurun_out/r2_gputest.log; tail -c 300 gpurun_out/r2_bench.json; tail -5 gpurun_out/r2_sanitizer_memcheck.log
