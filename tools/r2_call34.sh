#!/bin/bash
# round 2, call 34: trainer with bounded + shared dedup; kernel-time breakdown of the graphed API step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_layer_trainer.py -m gpu -q -x > $O/c34_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c34_pytest.log
tail -6 $O/c34_pytest.log
timeout 300 python tools/api_graph_profile.py > $O/c34_api_graph_profile.json 2> $O/c34_prof.err; echo "prof rc=$?"; tail -3 $O/c34_prof.err
head -c 6000 $O/c34_api_graph_profile.json
timeout 300 python bench.py --no-cpu-baseline > $O/c34_bench.json 2> $O/c34_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c34_bench.json').read().strip().splitlines()[-1])
    print(round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1))
    a=d.get('api_path'); print({k:(round(v['ms_per_step'],3) if isinstance(v,dict) else v) for k,v in a.items() if k!='what'})
except Exception as e: print('ERR', e)
PY
