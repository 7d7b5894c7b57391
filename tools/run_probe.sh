# DMA-interference probe for both exchange modes (GPU box): bash tools/run_probe.sh [N]
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
for m in owner direct; do
  timeout 200 $TR tools/xchg_dma_probe.py $m > gpurun_out/probe_$m.log 2>&1; grep '^{' gpurun_out/probe_$m.log || tail -20 gpurun_out/probe_$m.log
done
