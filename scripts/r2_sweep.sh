#!/bin/bash
# BASELINE config 5 (CoNIC-scale sweep, 4981 tiles of 256x256, C = 7) at 1 / 4 / 8 GPUs: batched records vs per-image dictionaries
run() {  # n_gpus tag extra-args
  local n=$1 tag=$2; shift 2
  if [ "$n" = 1 ]; then timeout 400 python scripts/conic_sweep.py "$@" > gpurun_out/r2_sweep_$tag.json 2> gpurun_out/r2_sweep_$tag.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 scripts/conic_sweep.py "$@" > gpurun_out/r2_sweep_$tag.json 2> gpurun_out/r2_sweep_$tag.err; fi
  echo "sweep $tag rc=$?"; tail -c 300 gpurun_out/r2_sweep_$tag.json | head -c 300; echo
}
run 8 8gpu
run 8 8gpu_per_image --per-image
run 4 4gpu
run 1 1gpu
run 1 1gpu_per_image --per-image
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 12 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; echo "bench8 rc=$?"
