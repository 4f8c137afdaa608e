#!/bin/bash
# exploration (not a benchmark result)
run() { echo "== $*"; env "$@" python bench.py --steps 2 --warmup 2 --no-cpu-baseline --check-reads 0 > gpurun_out/sw.json 2> gpurun_out/sw_err_$1.txt; python - <<'PY'
import json
for l in open('gpurun_out/sw.json'):
    if l.startswith('{'):
        d=json.loads(l); print("value %.2f Mb/s e2e %.2f Mb/s ms/step %.0f cores busy %.1f stage %s" % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['details']['host_cpu_cores_busy'], {k:int(v) for k,v in d['details']['stage_ms_per_step'].items()}))
PY
grep getSeeds gpurun_out/sw_err_$1.txt | tail -2; grep "myers_run" gpurun_out/sw_err_$1.txt | tail -1
}
run RTK_BROKER_PROFILE=1
run RTK_NO_WAIT_BACKOFF=1 RTK_BROKER_PROFILE=1
