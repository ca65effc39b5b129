"""Two identical steps of a bench workload on its default batch (target of the ncu metric passes; the digest keeps the
second).   python scripts/step_pass.py [workload] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import tiseg_b200
from tiseg_b200 import _lib, ops
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "dist_monuseg_1000"]()
B = int(sys.argv[2]) if len(sys.argv) > 2 else wl.batch
tiles = bench.make_tiles(4, 0, wl)
host = bench.stack_batch(tiles, B, wl.dihedral)
d = {k: torch.from_numpy(v).cuda() for k, v in host.items()}
with _lib.device_outputs():
    for _ in range(int(os.environ.get("REPS", "2"))):
        rec = wl.flat(wl.gpu_step(ops, d))
    torch.cuda.synchronize()
print("done", float(rec[:, 0].sum() / rec[:, 1].sum()))
