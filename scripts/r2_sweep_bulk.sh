#!/bin/bash
# exploration: gang / wave scheduling variants (not a benchmark result)
run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/sw.json 2> gpurun_out/sw_$1.err; python - <<'PY'
import json
for l in open('gpurun_out/sw.json'):
    if l.startswith('{'):
        d=json.loads(l); print("value %.2f Mb/s e2e %.2f Mb/s ms/step %.0f reqs %.0f stage %s" % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['config']['gpu_requests_per_step'], d['config']['stage_ms_per_step']))
PY
}
run RTK_GANGS=2
run RTK_GANGS=3
run RTK_GANGS=4
run RTK_GANGS=4 RTK_BULK_ALL=1
run RTK_GANGS=6 RTK_BULK_ALL=1
