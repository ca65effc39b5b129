"""Runs the DIST post-process on the bench batch a few times (target of ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import tiseg_b200
from tiseg_b200 import _lib, ops
tiles = bench.make_tiles(8, 0)
host = bench.stack_batch(tiles, 32)
d = torch.from_numpy(host["dist_logit"]).cuda()
with _lib.device_outputs():
    for _ in range(int(os.environ.get("REPS", "4"))):
        ops.postproc_dist(d)
    torch.cuda.synchronize()
print("done")
