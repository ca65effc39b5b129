#!/bin/bash
# multi-GPU: H2D probe + bench at N GPUs (N = $1)
N=${1:-8}
timeout 120 python scripts/h2d_probe.py > gpurun_out/r2_h2d_probe_$N.json 2> gpurun_out/r2_h2d_probe.err
nvidia-smi topo -m > gpurun_out/r2_topo_$N.txt 2>&1
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_bench_${n}gpu.json 2> gpurun_out/r2_bench_${n}gpu.err
    echo "n=$n rc=$?"
  fi
done
