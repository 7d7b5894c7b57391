#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
tail -3 gpurun_out/c9_pytest.log
timeout 200 python tools/unique_timeline.py > gpurun_out/c9_unique_timeline.jsonl 2> gpurun_out/c9_unique_timeline.err; echo "timeline rc=$?"
cat gpurun_out/c9_unique_timeline.jsonl
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tower tile > gpurun_out/c9_bench_tile.json 2> gpurun_out/c9_bench_tile.err
B200_LOOKAHEAD_UNIQUE_BLOCKS=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tower tile > gpurun_out/c9_bench_tile_fg.json 2> gpurun_out/c9_bench_tile_fg.err
B200_LOOKAHEAD_UNIQUE_BLOCKS=2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tower tile > gpurun_out/c9_bench_tile_bps2.json 2> gpurun_out/c9_bench_tile_bps2.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tower tile --lookahead off > gpurun_out/c9_bench_tile_nola.json 2> gpurun_out/c9_bench_tile_nola.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tower fused --lookahead off > gpurun_out/c9_bench_fused_nola.json 2> gpurun_out/c9_bench_fused_nola.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c9_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/c9_prof_step \
    python bench.py --profile-step --warmup 3 --no-cpu-baseline --tower tile > gpurun_out/c9_prof_bench.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench_kernels.py > gpurun_out/c9_kernels.jsonl 2> gpurun_out/c9_kernels.err; echo "kernels rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/c9_kernels.jsonl'):
    d=json.loads(l); print(d['kernel'], d.get('dim'), d.get('unique_ids', d.get('k', d.get('replicas'))), round(d['us'],1), 'us', round(d['frac_of_peak'],3))
PY
