#!/bin/bash
# exploration: few host threads per rank (what a rank of an 8-GPU node gets), gang count (not a benchmark result)
run() { echo "== $*"; env "$@" python bench.py --steps 2 --warmup 2 --no-cpu-baseline --check-reads 0 --bases-per-step 33554432 > gpurun_out/sw.json 2> gpurun_out/sw_err.txt; python - <<'PY'
import json
for l in open('gpurun_out/sw.json'):
    if l.startswith('{'):
        d=json.loads(l); print("value %.2f Mb/s e2e %.2f Mb/s ms/step %.0f cores busy %.1f stage %s" % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['details']['host_cpu_cores_busy'], {k:int(v) for k,v in d['details']['stage_ms_per_step'].items()}))
PY
tail -1 gpurun_out/sw_err.txt | cut -c1-200
}
run RTK_HOST_THREADS=4 RTK_GANGS2=3
run RTK_HOST_THREADS=4 RTK_GANGS2=1
run RTK_HOST_THREADS=4 RTK_GANGS2=2
