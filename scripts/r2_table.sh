#!/bin/bash
# north-star table: the four other workloads at 8 / 4 / 2 GPUs of one box (dist_monuseg_1000 is the default bench line: r2_bench_{2,4,8}gpu.json)
port=29530
for n in 8 4 2; do
  for wl in unet_cpm17_256 conic_sweep_256 cdnet_consep_1000 hover_consep_1000; do
    port=$((port+1))
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --workload $wl --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_tab_${wl}_${n}gpu.json 2> gpurun_out/r2_tab_${wl}_${n}gpu.err
    echo "$wl n=$n rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/r2_tab_${wl}_${n}gpu.json').read().strip().splitlines()[-1]);print(round(d['value']),round(d['e2e']['value']))" 2>/dev/null)"
  done
done
