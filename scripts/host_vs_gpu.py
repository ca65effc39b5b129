"""Host enqueue time vs GPU time of one DIST post-process + eval pass, by batch size (device-resident inputs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import tiseg_b200
from tiseg_b200 import _lib, ops
tiles = bench.make_tiles(8, 0)
for B in (4, 8, 16, 32):
    host = bench.stack_batch(tiles, B)
    d = {k: torch.from_numpy(v).cuda() for k, v in host.items()}
    ctx = _lib.get_ctx(0)
    def step():
        with _lib.device_outputs():
            cls = ops.softmax_argmax(d["sem_logit"])
            inst = ops.postproc_dist(d["dist_logit"])
            aji, pq = ops.pair_metrics_bin(inst, d["gt_inst"])
            counts, valid = ops.sem_counts(cls, d["gt_sem"], 2)
    for _ in range(3): step()
    torch.cuda.synchronize()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps): step()
    e1.record(); th = time.perf_counter() - t0
    torch.cuda.synchronize()
    # host-only: enqueue while the GPU is blocked?  approximate by single pass after sync
    torch.cuda.synchronize(); t1 = time.perf_counter(); step(); th1 = time.perf_counter() - t1; torch.cuda.synchronize()
    print("B=%2d launches/pass %d  gpu ms/pass %.3f  host enqueue ms/pass (pipelined) %.3f  (single, idle GPU) %.3f" % (
        B, (ctx.launch_count() - l0) // reps, e0.elapsed_time(e1) / reps, th / reps * 1e3, th1 * 1e3))
