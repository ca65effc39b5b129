"""One DIST post-process + eval pass on the bench batch, REPS times (target of ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import tiseg_b200
from tiseg_b200 import _lib, ops
tiles = bench.make_tiles(8, 0)
host = bench.stack_batch(tiles, int(os.environ.get("BATCH", "64")))
d = {k: torch.from_numpy(v).cuda() for k, v in host.items()}
with _lib.device_outputs():
    for _ in range(int(os.environ.get("REPS", "2"))):
        cls = ops.softmax_argmax(d["sem_logit"])
        inst = ops.postproc_dist(d["dist_logit"])
        aji, pq = ops.pair_metrics_bin(inst, d["gt_inst"])
        counts, valid = ops.sem_counts(cls, d["gt_sem"], 2)
    torch.cuda.synchronize()
print("done", float(aji.sum()))
