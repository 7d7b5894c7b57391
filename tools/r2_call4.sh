#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layer_trainer.py -m gpu -x -q -k "unique or engine or train_loop" > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
tail -3 gpurun_out/c4_pytest.log
timeout 200 python tools/unique_timeline.py > gpurun_out/c4_unique_timeline.jsonl 2> gpurun_out/c4_unique_timeline.err; echo "timeline rc=$?"
cat gpurun_out/c4_unique_timeline.jsonl
for tw in fused tile; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tower $tw > gpurun_out/c4_bench_$tw.json 2> gpurun_out/c4_bench_$tw.err; echo "bench $tw rc=$?"
done
B200_LOOKAHEAD_UNIQUE_BLOCKS=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tower tile > gpurun_out/c4_bench_tile_fg.json 2> gpurun_out/c4_bench_tile_fg.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --tower tile --lookahead off > gpurun_out/c4_bench_tile_nola.json 2> gpurun_out/c4_bench_tile_nola.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c4_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step']*1e3,1),'us', round(d['value']/1e6,1),'M/s  e2e', round(d['e2e']['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/c4_prof_step \
    python bench.py --profile-step --warmup 3 --no-cpu-baseline --tower tile > gpurun_out/c4_prof_bench.log 2>&1; echo "ncu rc=$?"
