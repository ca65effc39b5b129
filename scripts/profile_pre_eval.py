import cProfile, pstats, os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tiseg_b200
from tiseg_b200 import datasets, segmentors, synth
C = 7
base = [synth.tile_unet(5, j, 256, 256, C) for j in range(32)]
tiles = [base[i % 32] for i in range(512)]
ds = datasets.CoNICDataset(sem_gts=[t["gt_sem"] for t in tiles], inst_gts=[t["gt_inst"] for t in tiles], names=["%d" % i for i in range(512)])
logits = torch.from_numpy(np.stack([t["sem_logit"][None] for t in tiles])).cuda()
post = segmentors.UNet(C)
def run():
    preds = post.forward_eval(logits)
    r = ds.pre_eval(preds, list(range(512)))
    torch.cuda.synchronize()
    return r
run(); run()
import time
t0 = time.perf_counter(); run(); print("one batch of 512: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:4000])
