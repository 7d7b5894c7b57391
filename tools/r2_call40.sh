#!/bin/bash
# round 2, call 40: the default bench line of the final tree (CPU arm with the reference's compiled Adam)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/c40_bench.json 2> gpurun_out/c40_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c40_bench.json').read().strip().splitlines()[-1])
    print(round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1))
    print(d['cpu_baseline']); a=d.get('api_path'); print({k:(round(v['ms_per_step'],3) if isinstance(v,dict) else v) for k,v in a.items() if k!='what'})
except Exception as e: print('ERR', e)
PY
tail -2 gpurun_out/c40_bench.err
