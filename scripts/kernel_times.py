"""Per-kernel CUDA-event times of one pass of a BASELINE configuration (3 = HoVer, 5 = CoNIC, 1, 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tiseg_b200
from tiseg_b200 import _lib, ops, synth
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
def stack(tiles, keys, batch):
    return {k: torch.from_numpy(np.stack([tiles[b % len(tiles)][k] for b in range(batch)])).cuda() for k in keys}
if cfg == 3:
    tiles = [synth.tile_hover(3, j) for j in range(4)]
    d = stack(tiles, ["sem_logit", "fore_map", "hv_map", "gt_inst", "gt_sem"], 16)
    def fn():
        cls = ops.softmax_argmax(d["sem_logit"][:, None])
        inst = ops.postproc_hover(d["fore_map"], d["hv_map"])
        ops.pair_metrics_bin(inst, d["gt_inst"]); ops.sem_counts(cls, d["gt_sem"], 3)
elif cfg == 5:
    tiles = [synth.tile_unet(5, j, 256, 256, 7) for j in range(16)]
    d = stack(tiles, ["sem_logit", "gt_inst", "gt_sem"], 512)
    def fn():
        cls = ops.softmax_argmax(d["sem_logit"][:, None])
        sem, inst = ops.postproc_unet(cls, 6, 1, None)
        ops.pair_metrics_multiclass(inst, sem, d["gt_inst"], d["gt_sem"], 7); ops.sem_counts(sem, d["gt_sem"], 7)
elif cfg == 1:
    tiles = [synth.tile_unet(1, j, 256, 256, 2) for j in range(16)]
    d = stack(tiles, ["sem_logit", "gt_inst", "gt_sem"], 512)
    def fn():
        cls = ops.softmax_argmax(d["sem_logit"][:, None])
        sem, inst = ops.postproc_unet(cls, 1, 1, None)
        ops.pair_metrics_bin(inst, d["gt_inst"]); ops.sem_counts(sem, d["gt_sem"], 2)
else:
    tiles = [synth.tile_cdnet(4, j, T=1) for j in range(4)]
    d = stack(tiles, ["sem_logit", "dir_logit", "point_logit", "gt_inst", "gt_sem"], 16)
    def fn():
        r = ops.cdnet_refine(d["sem_logit"], d["dir_logit"], d["point_logit"], if_ddm=True)
        sem, inst = ops.postproc_unet(r["cls"], 2, 3, 2)
        ops.pair_metrics_bin(inst, d["gt_inst"]); ops.sem_counts(sem, d["gt_sem"], 2)
ctx = _lib.get_ctx(0)
with _lib.device_outputs():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ctx.timing(True); fn(); rep = ctx.timing_report(); ctx.timing(False)
tot = sum(v[1] for v in rep.values())
print("config", cfg, "total %.3f ms, %d launches" % (tot, sum(v[0] for v in rep.values())))
for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])[:28]:
    print("%-44s x%-3d %.4f ms" % (k[:44], v[0], v[1]))
