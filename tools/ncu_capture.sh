# ncu evidence for profiles/ (1 GPU): --set full of every kernel of ONE eager training step
# (bench.py --profile-step brackets it with cudaProfilerStart/Stop), plus the launch list.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_step \
    python bench.py --profile-step --warmup 3 --no-cpu-baseline > gpurun_out/prof_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --profile-step --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
