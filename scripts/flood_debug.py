"""Diagnostics: flood work-list census + per-kernel timing of the DIST post-process on the bench tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TISEG_DEBUG_FLOOD"] = "1"
import numpy as np, torch
import bench
import tiseg_b200
from tiseg_b200 import _lib, ops
tiles = bench.make_tiles(8, 0)
host = bench.stack_batch(tiles, 32)
d = torch.from_numpy(host["dist_logit"]).cuda()
ctx = _lib.get_ctx(0)
with _lib.device_outputs():
    for _ in range(3):
        ops.postproc_dist(d)
    torch.cuda.synchronize()
    ctx.timing(True)
    ops.postproc_dist(d)
    rep = ctx.timing_report()
    ctx.timing(False)
for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print("%-36s x%d %.4f ms" % (k, v[0], v[1]))
