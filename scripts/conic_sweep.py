#!/usr/bin/env python
"""BASELINE config 5: the CoNIC-scale sweep — 4981 synthetic 256x256 tiles, C = 7, UNet-family post-process +
CoNICDataset evaluation (binary and per-class AJI / PQ, semantic metrics), sharded over the GPUs of one box.

    python scripts/conic_sweep.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/conic_sweep.py

Tiles are sharded by index (DistributedSampler interleave), every rank runs its shard in batches through the
reference-facing API (segmentors.UNet.forward_eval + CoNICDataset.pre_eval), the per-image records are all-gathered
(parallel.gather_results) and rank 0 evaluates.  Prints one JSON line: tiles/s (device time, max over ranks) and the
dataset metrics, which must not depend on the number of GPUs.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import tiseg_b200
from tiseg_b200 import datasets, ops, parallel, segmentors, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=4981)
    ap.add_argument("--distinct", type=int, default=64, help="distinct synthetic tiles (cycled through 8 dihedral variants)")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--per-image", action="store_true", help="go through pre_eval's per-image dictionaries (the reference's protocol)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    C = 7
    base = [synth.tile_unet(5, j, 256, 256, C) for j in range(a.distinct)]

    def tile(i):            # tile i of the sweep: a dihedral variant of a base tile (deterministic in i)
        t, k = base[i % a.distinct], (i // a.distinct) % 8
        def var(x):
            x = np.rot90(x, k % 4, axes=(-2, -1))
            return np.ascontiguousarray(x[..., ::-1] if k >= 4 else x)
        return {key: var(t[key]) for key in ("sem_logit", "gt_sem", "gt_inst")}

    mine = parallel.shard_indices(a.tiles, rank, world)
    tiles = [tile(i) for i in mine]
    # ground truth and logits are resident in HBM when the clock starts (GT reading is a separate I/O row)
    gsem = torch.from_numpy(np.stack([t["gt_sem"] for t in tiles])).cuda()
    ginst = torch.from_numpy(np.stack([t["gt_inst"] for t in tiles]).astype(np.int32)).cuda()
    batched = not a.per_image
    ds = datasets.CoNICDataset(sem_gts=gsem if batched else list(gsem), inst_gts=ginst if batched else list(ginst),
                               names=["%d" % i for i in mine])
    logits = torch.from_numpy(np.stack([t["sem_logit"][None] for t in tiles])).cuda()       # [n, T=1, C, H, W] resident
    post = segmentors.UNet(C)
    results = []
    if batched:                                                     # untimed warm-up batch (workspace growth)
        ds.pre_eval_records(*post.postprocess(ops.softmax_argmax(logits[:a.batch])), list(range(min(a.batch, len(mine)))))
    else:
        ds.pre_eval(post.forward_eval(logits[:a.batch]), list(range(min(a.batch, len(mine)))))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for lo in range(0, len(mine), a.batch):
        tb = time.perf_counter()
        idx = list(range(lo, min(lo + a.batch, len(mine))))
        if batched:          # segmentor tail + batched records: no per-image Python between the CNN and the gather
            sem_pred, inst_pred = post.postprocess(ops.softmax_argmax(logits[lo:lo + a.batch]))
            tm = time.perf_counter()
            results.append(ds.pre_eval_records(sem_pred, inst_pred, idx))
        else:
            preds = post.forward_eval(logits[lo:lo + a.batch])
            tm = time.perf_counter()
            results.extend(ds.pre_eval(preds, idx))
        if os.environ.get("SWEEP_DEBUG"):
            sys.stderr.write("rank %d batch @%d: forward_eval %.1f ms, pre_eval %.1f ms\n" % (
                rank, lo, (tm - tb) * 1e3, (time.perf_counter() - tm) * 1e3))
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if batched:
        rec = torch.cat(results)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            full_rec = parallel.gather_records(rec, mine, a.tiles)
        else:
            full_rec = rec.cpu().numpy()
        results = parallel.unpack_results(full_rec, C) if rank == 0 else None
    elif world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        results = parallel.gather_results(results, mine, a.tiles, C)
    if rank == 0:
        full = datasets.CoNICDataset(sem_gts=[None] * a.tiles, inst_gts=[None] * a.tiles, names=["%d" % i for i in range(a.tiles)])
        ev, _ = full.evaluate(results, logger="silent")
        print(json.dumps({"workload": "conic_sweep_256", "tiles": a.tiles, "n_gpus": world, "ms": float(ms.item()),
                          "tiles_per_s": a.tiles / (float(ms.item()) / 1e3),
                          "metrics": {k: float(v) for k, v in ev.items()}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
