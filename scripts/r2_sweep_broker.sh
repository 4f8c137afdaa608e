#!/bin/bash
# quick sweep of broker settings (not a benchmark result; exploration only)
run() { echo "== $*"; env "$@" RTK_BROKER_PROFILE=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/sw.json 2> gpurun_out/sw.err; python - <<'PY'
import json
for l in open('gpurun_out/sw.json'):
    if l.startswith('{'):
        d=json.loads(l); print("value %.2f Mb/s e2e %.2f Mb/s ms/step %.0f reqs/step %.0f stage %s" % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['config']['gpu_requests_per_step'], d['config']['stage_ms_per_step']))
PY
grep -E "tasks=" gpurun_out/sw.err | tail -1 | cut -c1-100; }
run X=1
run RTK_CORRECT_INFLIGHT=131072
run RTK_CORRECT_INFLIGHT=131072 RTK_SERVICE_MIN_BATCH=4096 RTK_SERVICE_LINGER_US=2000
run RTK_CORRECT_INFLIGHT=65536 RTK_SERVICE_THREADS=2,2,2,3 RTK_RG_WARPS_PER_SM=12
