#!/usr/bin/env python
"""Throughput of the five BASELINE.json configurations through the operator API, inputs resident in HBM (one GPU).
Not the driver's bench line (that is bench.py = config 2); a table for DESIGN.md / profiles/.

  python scripts/bench_configs.py [--steps 5] > gpurun_out/configs.json
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tiseg_b200
from tiseg_b200 import _lib, ops, synth

PEAK = 6547.2


def dihedral(a, k):
    a = np.rot90(a, k % 4, axes=(0, 1)) if a.ndim >= 2 else a
    return np.ascontiguousarray(a[:, ::-1] if k >= 4 else a)


def stack(tiles, keys, batch, chw=()):
    out = {}
    for key in keys:
        arrs = []
        for b in range(batch):
            t = tiles[b % len(tiles)][key]
            arrs.append(t)
        out[key] = torch.from_numpy(np.stack(arrs)).cuda()
    return out


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    rows = []

    def report(name, batch, H, W, bpp, ms):
        tiles_s = batch / (ms / 1e3)
        rows.append({"config": name, "tile": [H, W], "tiles_per_step": batch, "ms_per_step": ms, "tiles_per_s": tiles_s,
                     "alg_bytes_per_px": bpp, "pipeline_frac_of_hbm": bpp * H * W * tiles_s / 1e9 / PEAK})
        print(json.dumps(rows[-1]), flush=True)

    with _lib.device_outputs():
        # 1. UNet CPM17 256^2, C = 2
        tiles = [synth.tile_unet(1, j, 256, 256, 2) for j in range(16)]
        d = stack(tiles, ["sem_logit", "gt_inst", "gt_sem"], 512)
        lg = d["sem_logit"][:, None]

        def c1():
            cls = ops.softmax_argmax(lg)
            sem, inst = ops.postproc_unet(cls, 1, 1, None)
            ops.pair_metrics_bin(inst, d["gt_inst"]); ops.sem_counts(sem, d["gt_sem"], 2)
        report("1 UNet CPM17 256x256", 512, 256, 256, 18, timed(c1, a.steps))

        # 2. DIST MoNuSeg 1000^2
        tiles = [synth.tile_dist(2, j) for j in range(8)]
        d = stack(tiles, ["sem_logit", "dist_logit", "gt_inst", "gt_sem"], 32)
        lg2 = d["sem_logit"][:, None]

        def c2():
            cls = ops.softmax_argmax(lg2)
            inst = ops.postproc_dist(d["dist_logit"])
            ops.pair_metrics_bin(inst, d["gt_inst"]); ops.sem_counts(cls, d["gt_sem"], 2)
        report("2 DIST MoNuSeg 1000x1000", 32, 1000, 1000, 22, timed(c2, a.steps))

        # 3. HoVer-Net CoNSeP 1000^2
        tiles = [synth.tile_hover(3, j) for j in range(4)]
        d = stack(tiles, ["sem_logit", "fore_map", "hv_map", "gt_inst", "gt_sem"], 16)
        lg3 = d["sem_logit"][:, None]
        C3 = int(d["sem_logit"].shape[1])

        def c3():
            cls = ops.softmax_argmax(lg3)
            inst = ops.postproc_hover(d["fore_map"], d["hv_map"])
            ops.pair_metrics_bin(inst, d["gt_inst"]); ops.sem_counts(cls, d["gt_sem"], C3)
        report("3 HoVer-Net CoNSeP 1000x1000", 16, 1000, 1000, 38, timed(c3, a.steps))

        # 4. CDNet CoNSeP 1000^2 (T = 1)
        tiles = [synth.tile_cdnet(4, j, T=1) for j in range(4)]
        d = stack(tiles, ["sem_logit", "dir_logit", "point_logit", "gt_inst", "gt_sem"], 16)

        def c4():
            r = ops.cdnet_refine(d["sem_logit"], d["dir_logit"], d["point_logit"], if_ddm=True)
            sem, inst = ops.postproc_unet(r["cls"], 2, 3, 2)
            ops.pair_metrics_bin(inst, d["gt_inst"]); ops.sem_counts(sem, d["gt_sem"], 2)
        report("4 CDNet CoNSeP 1000x1000", 16, 1000, 1000, 62, timed(c4, a.steps))

        # 5. CoNIC-scale sweep: 256^2 tiles, C = 7, binary + per-class AJI / PQ
        tiles = [synth.tile_unet(5, j, 256, 256, 7) for j in range(16)]
        d = stack(tiles, ["sem_logit", "gt_inst", "gt_sem"], 512)
        lg5 = d["sem_logit"][:, None]

        def c5():
            cls = ops.softmax_argmax(lg5)
            sem, inst = ops.postproc_unet(cls, 6, 1, None)
            ops.pair_metrics_multiclass(inst, sem, d["gt_inst"], d["gt_sem"], 7); ops.sem_counts(sem, d["gt_sem"], 7)
        ms = timed(c5, a.steps)
        report("5 CoNIC sweep 256x256 (4981 tiles = %.1f ms at this rate)" % (4981 / 512 * ms), 512, 256, 256, 38, ms)
    json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
