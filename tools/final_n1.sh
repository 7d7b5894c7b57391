# Round-end evidence at N=1: bench (with CPU baseline), reference arm, tower knob A/B, ncu captures.
mkdir -p gpurun_out
python bench.py > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
B200_TOWER_GC=2 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/final_gc2.json 2> gpurun_out/final_gc2.err
python - <<PY
import json
for f in ("final_n1", "final_gc2", "final_ref"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, d["value"], d.get("ms_per_step"), d.get("e2e"), d.get("cpu_baseline"), d.get("roofline"))
    except Exception as e:
        print(f, "ERR", e, open("gpurun_out/%s.err" % f).read()[-1500:])
PY
timeout 600 bash tools/ncu_capture.sh 2>&1 | tail -2
