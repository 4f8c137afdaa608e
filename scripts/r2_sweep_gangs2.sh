#!/bin/bash
# exploration: gang count of the pipelined two-pass call (not a benchmark result)
run() { echo "== $*"; env "$@" python bench.py --steps 2 --warmup 2 --no-cpu-baseline --check-reads 0 > gpurun_out/sw.json 2> gpurun_out/sw_err.txt; python - <<'PY'
import json
for l in open('gpurun_out/sw.json'):
    if l.startswith('{'):
        d=json.loads(l); print("value %.2f Mb/s e2e %.2f Mb/s ms/step %.0f stage %s" % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], {k:int(v) for k,v in d['details']['stage_ms_per_step'].items()}))
PY
tail -2 gpurun_out/sw_err.txt | cut -c1-300
}
python -m pytest tests -m gpu -q -x -k "pipeline" 2>&1 | tail -3
run RTK_GANGS2=3
run RTK_GANGS2=4
run RTK_GANGS2=6
run RTK_GANGS2=8
