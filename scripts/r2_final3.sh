#!/bin/bash
# Round-2 final evidence, third part (index-build rows): the whole GPU suite on the final library, the smoke path, colouring timing,
# one `ncu --set full` capture of the annotation kernels on the chr20-scale graph, compute-sanitizer memcheck of the annotation +
# colouring tests.  Outputs: gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_gputest_final.log 2>&1
tail -3 gpurun_out/r2_gputest_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python scripts/color_quick.py > gpurun_out/r2_color_quick.json 2> gpurun_out/r2_color_quick.err; cat gpurun_out/r2_color_quick.json
timeout 100 ncu --set full --clock-control none --import-source on -k "regex:rtk_(snp|cycles|edge_flags)_kernel" -c 3 -f -o gpurun_out/r2_annotate \
    python scripts/annotate_quick.py bench_data/F4 31 > gpurun_out/r2_ncu_annotate.log 2>&1
tail -2 gpurun_out/r2_ncu_annotate.log | cut -c 1-400
timeout 100 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_annotate.py tests/test_color.py -m gpu -q -k "F1 or F2-63 or F2]" \
    > gpurun_out/r2_sanitizer_annotate.log 2>&1
tail -6 gpurun_out/r2_sanitizer_annotate.log
