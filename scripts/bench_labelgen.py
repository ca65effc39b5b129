#!/usr/bin/env python
"""Throughput of the train-time label makers on 1000x1000 MoNuSeg-like instance maps, inputs
resident in HBM, one GPU; the CPU side is timed with the reference-shaped label generation restated under tests' oracle
— imported here ONLY when --cpu is given (this script is a measurement tool, not product code).

  python scripts/bench_labelgen.py [--batch 16] [--steps 5] [--cpu] > gpurun_out/labelgen.json
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tiseg_b200
from tiseg_b200 import _lib, ops, synth


def timed(fn, steps, warmup=2):
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
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--size", type=int, default=1000)
    ap.add_argument("--only", default="", help="substring of the label maker to time")
    a = ap.parse_args()
    S = a.size
    tiles = [synth.gt_and_pred(4200 + j, S, S, n=max(1, 900 * S * S // 1000000)) for j in range(4)]
    inst_np = np.stack([tiles[b % 4]["gt_inst"].astype(np.int32) for b in range(a.batch)])
    sem_np = (inst_np > 0).astype(np.uint8)
    inst, sem = torch.from_numpy(inst_np).cuda(), torch.from_numpy(sem_np).cuda()
    with _lib.device_outputs():
        fixed = ops.fix_inst(inst)
        cases = [
            ("fix_inst", lambda: ops.fix_inst(inst)),
            ("HVLabelMake (gen_instance_hv_map)", lambda: ops.gen_instance_hv_map(inst)),
            ("DistanceLabelMake (fix_inst + distance map)", lambda: ops.instance_distance_map(ops.fix_inst(inst), True)),
            ("BoundLabelMake (fix_inst + boundary label)", lambda: ops.bound_label(sem, ops.fix_inst(inst), 2, 3)),
            ("UNetLabelMake (fix_inst + weight map)", lambda: ops.unet_weight_map(ops.fix_inst(inst), 10.0, 5.0)),
            ("DirectionLabelMake (fix_inst + centre / distance / direction / point / weight maps)",
             lambda: ops.direction_labels(ops.fix_inst(inst), 8)),
        ]
        for name, fn in cases:
            if a.only and a.only not in name:
                continue
            ms = timed(fn, a.steps)
            print(json.dumps({"label_maker": name, "tile": [S, S], "tiles_per_step": a.batch, "ms_per_step": ms,
                              "tiles_per_s": a.batch / (ms / 1e3)}), flush=True)


if __name__ == "__main__":
    main()
