"""GPU parity of the whole path through the reference-facing API (segmentor tails + Dataset.pre_eval /
evaluate) for the five BASELINE configs, against the same path assembled from the CPU oracle."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import tiseg_b200  # noqa: E402,F401
from tiseg_b200 import datasets, segmentors, synth  # noqa: E402
from oracle import postprocess as opp  # noqa: E402
from refpipe import oracle_pre_eval, same_result  # noqa: E402


def _check(ds_cls, C, tiles, preds, want_preds, multi=False):
    ds = ds_cls(sem_gts=[t['gt_sem'] for t in tiles], inst_gts=[t['gt_inst'] for t in tiles],
                names=["t%d" % i for i in range(len(tiles))])
    for i, (p, w) in enumerate(zip(preds, want_preds)):
        assert np.array_equal(np.asarray(p['inst_pred']), w['inst_pred']), "inst_pred of tile %d differs" % i
        assert np.array_equal(np.asarray(p['sem_pred']), w['sem_pred']), "sem_pred of tile %d differs" % i
    got = ds.pre_eval(preds, list(range(len(tiles))))
    want = [oracle_pre_eval(w, t['gt_sem'], t['gt_inst'], C, multi=multi, name=None if multi else "t%d" % i)
            for i, (w, t) in enumerate(zip(want_preds, tiles))]
    for a, b in zip(want, got):
        if multi:
            b = {k: v for k, v in b.items() if k != 'name'}
        same_result(a, b)
    ev_got, _ = ds.evaluate(got, logger="silent")
    ev_want, _ = ds.evaluate(want, logger="silent")
    assert ev_got == ev_want and len(ev_got) >= 12
    return ev_got


def test_config1_unet_cpm17_256():
    C = 2
    tiles = [synth.tile_unet(1, j, 256, 256, C) for j in range(4)]
    post = segmentors.UNet(C)
    preds = post.forward_eval(np.stack([t['sem_logit'][None] for t in tiles]))
    want = []
    for t in tiles:
        cls = opp.argmax_classes(opp.softmax(t['sem_logit']))
        s, i = opp.unet_family_postprocess(cls, radius=1)
        want.append(dict(sem_pred=s, inst_pred=i))
    ev = _check(datasets.CPM17Dataset, C, tiles, preds, want)
    assert 30 < ev['mAji'] < 100


@pytest.mark.parametrize("size", [256, 1000])
def test_config2_dist_monuseg(size):
    C = 2
    tiles = [synth.tile_dist(2, j, size, size) for j in range(2)]
    post = segmentors.Dist(C)
    preds = post.forward_eval(np.stack([t['sem_logit'][None] for t in tiles]), np.stack([t['dist_logit'] for t in tiles]))
    want = []
    for t in tiles:
        cls = opp.argmax_classes(opp.softmax(t['sem_logit']))
        _, i = opp.dist_postprocess(cls, t['dist_logit'], literal=False)
        want.append(dict(sem_pred=cls.astype(np.uint8), inst_pred=i))
    _check(datasets.MoNuSegDataset, C, tiles, preds, want)


@pytest.mark.parametrize("size", [256, 1000])
def test_config3_hovernet_consep(size):
    C = 3
    tiles = [synth.tile_hover(3, j, size, size) for j in range(2)]
    post = segmentors.HoverNet(C)
    preds = post.forward_eval(np.stack([t['sem_logit'][None] for t in tiles]), np.stack([t['hv_map'] for t in tiles]),
                              np.stack([t['fore_map'] for t in tiles]))
    want = []
    for t in tiles:
        cls = opp.argmax_classes(opp.softmax(t['sem_logit']))
        i, _ = opp.hover_post_proc(t['fore_map'], t['hv_map'])
        want.append(dict(sem_pred=cls.astype(np.uint8), inst_pred=i))
    class ThreeClass(datasets.CoNSePDataset):
        CLASSES = ('background', 'type1', 'type2')
    _check(ThreeClass, C, tiles, preds, want)


@pytest.mark.parametrize("size,T", [(256, 2), (1000, 1)])
def test_config4_cdnet_consep(size, T):
    tiles = [synth.tile_cdnet(4, j, size, size, T=T) for j in range(2)]
    post = segmentors.CDNet(2, test_cfg=dict(if_ddm=True))
    preds = post.forward_eval(np.stack([t['sem_logit'] for t in tiles]), np.stack([t['dir_logit'] for t in tiles]),
                              np.stack([t['point_logit'] for t in tiles]))
    want = []
    for t in tiles:
        sem, _, _ = opp.cdnet_inference_tail(list(t['sem_logit']), list(t['dir_logit']), list(t['point_logit']), True)
        cls = np.argmax(sem, 0).astype(np.int64)
        s, i = opp.unet_family_postprocess(cls, radius=3, edge_id=2)
        want.append(dict(sem_pred=s, inst_pred=i))
    _check(datasets.CoNSePDataset, 2, tiles, preds, want)


def test_config5_conic_sweep_slice():
    C = 7
    tiles = [synth.tile_unet(5, j, 256, 256, C) for j in range(12)]
    post = segmentors.UNet(C)
    preds = post.forward_eval(np.stack([t['sem_logit'][None] for t in tiles]))
    want = []
    for t in tiles:
        cls = opp.argmax_classes(opp.softmax(t['sem_logit']))
        s, i = opp.unet_family_postprocess(cls, radius=1)
        want.append(dict(sem_pred=s, inst_pred=i))
    ev = _check(datasets.CoNICDataset, C, tiles, preds, want, multi=True)
    assert 'Aji.neutrophil' in ev and 'bPQ' in ev


def test_device_resident_path_matches_host_path():
    import torch
    C = 2
    tiles = [synth.tile_dist(2, 10 + j, 256, 256) for j in range(3)]
    post = segmentors.Dist(C)
    lg = np.stack([t['sem_logit'][None] for t in tiles]); dm = np.stack([t['dist_logit'] for t in tiles])
    host = post.forward_eval(lg, dm)
    dev = post.forward_eval(torch.from_numpy(lg).cuda(), torch.from_numpy(dm).cuda())
    ds = datasets.MoNuSegDataset(sem_gts=[t['gt_sem'] for t in tiles], inst_gts=[t['gt_inst'] for t in tiles])
    a = ds.pre_eval(host, [0, 1, 2]); b = ds.pre_eval(dev, [0, 1, 2])
    for x, y in zip(a, b):
        assert dev[0]['inst_pred'].is_cuda
        same_result(x, y)


def test_host_feed_chunks_equal_single_call():
    """parallel.HostFeed (the caller loop for host-resident inputs): 40 host tiles fed in chunks of 8
    alternating over 4 lanes (own stream + own workspace each) give, tile by tile, the records of ONE call on the whole
    batch, and the oracle's records on two of them."""
    import torch
    from tiseg_b200 import _lib, ops, parallel
    from oracle import metrics as om
    H = W = 256
    base = [synth.tile_dist(2, 700 + j, H=H, W=W) for j in range(10)]
    n = 40
    var = lambda a, k: np.ascontiguousarray(np.rot90(a, k % 4, axes=(-2, -1))[..., ::-1] if k >= 4 else np.rot90(a, k % 4, axes=(-2, -1)))
    arrays = {key: np.stack([var(base[i % 10][key], i // 10) for i in range(n)]) for key in ("sem_logit", "dist_logit", "gt_inst", "gt_sem")}
    arrays["sem_logit"] = arrays["sem_logit"][:, None]
    pinned = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in arrays.items()}

    def records(src):
        cls = ops.softmax_argmax(src["sem_logit"])
        inst = ops.postproc_dist(src["dist_logit"])
        aji, pq = ops.pair_metrics_bin(inst, src["gt_inst"])
        counts, valid = ops.sem_counts(cls, src["gt_sem"], 2)
        return inst, aji, pq, counts

    with _lib.device_outputs():
        inst1, aji1, pq1, cnt1 = records(pinned)
    torch.cuda.synchronize()
    feed = parallel.HostFeed(0, lanes=4, chunk=8)
    dev = torch.device("cuda", 0)
    out = dict(inst=torch.zeros(n, H, W, dtype=torch.int32, device=dev), aji=torch.zeros(n, 2, dtype=torch.float64, device=dev),
               pq=torch.zeros(n, 4, dtype=torch.float64, device=dev), cnt=torch.zeros(n, 5, 2, dtype=torch.int64, device=dev))
    cursor = [0]

    def chunk_fn(src, lane):
        lo = cursor[0]
        inst, aji, pq, counts = records(src)
        k = inst.shape[0]
        out["inst"][lo:lo + k], out["aji"][lo:lo + k], out["pq"][lo:lo + k], out["cnt"][lo:lo + k] = inst, aji, pq, counts
        cursor[0] = lo + k

    for _ in range(2):                      # twice: the second pass reuses every lane's workspace
        cursor[0] = 0
        feed.run(pinned, chunk_fn)
    torch.cuda.synchronize()
    assert cursor[0] == n
    assert torch.equal(out["inst"], inst1) and torch.equal(out["aji"], aji1) and torch.equal(out["pq"], pq1) and torch.equal(out["cnt"], cnt1)
    for j in (3, 27):
        _, want_inst = opp.dist_postprocess(None, arrays["dist_logit"][j], literal=False)
        assert np.array_equal(out["inst"][j].cpu().numpy(), want_inst)
        assert tuple(out["aji"][j].cpu().numpy()) == tuple(np.float64(om.pre_eval_bin_aji(want_inst, arrays["gt_inst"][j], literal=False)))
        assert tuple(out["pq"][j].cpu().numpy()) == tuple(np.float64(om.pre_eval_bin_pq(want_inst, arrays["gt_inst"][j], literal=False)))



@pytest.mark.gpu
@pytest.mark.parametrize("multi", [False, True])
def test_batched_records_equal_per_image_dictionaries(multi):
    """Dataset.pre_eval_records (one [n, R] tensor, no per-image Python) carries exactly what pre_eval puts in the
    per-image dictionaries, for stacked CUDA ground truth (contiguous slice) and for list ground truth."""
    import torch
    from tiseg_b200 import ops, parallel
    C = 7 if multi else 2
    tiles = [synth.tile_unet(5 if multi else 1, j, 256, 256, C) for j in range(6)]
    post = segmentors.UNet(C)
    sem_pred, inst_pred = post.postprocess(ops.softmax_argmax(
        torch.from_numpy(np.stack([t['sem_logit'][None] for t in tiles])).cuda()))
    cls = datasets.CoNICDataset if multi else datasets.CPM17Dataset
    gsem = torch.from_numpy(np.stack([t['gt_sem'] for t in tiles])).cuda()
    ginst = torch.from_numpy(np.stack([t['gt_inst'] for t in tiles]).astype(np.int32)).cuda()
    names = ["t%d" % i for i in range(len(tiles))]
    stacked = cls(sem_gts=gsem, inst_gts=ginst, names=names)
    listed = cls(sem_gts=[t['gt_sem'] for t in tiles], inst_gts=[t['gt_inst'] for t in tiles], names=names)
    preds = [dict(sem_pred=sem_pred[i], inst_pred=inst_pred[i]) for i in range(len(tiles))]
    want = parallel.pack_results(listed.pre_eval(preds, list(range(len(tiles)))), C)
    rec = stacked.pre_eval_records(sem_pred, inst_pred, range(len(tiles)))
    assert rec.dtype == torch.float64 and rec.is_cuda and tuple(rec.shape) == want.shape
    assert np.array_equal(rec.cpu().numpy(), want)
    part = stacked.pre_eval_records(sem_pred[2:5], inst_pred[2:5], [2, 3, 4])
    assert np.array_equal(part.cpu().numpy(), want[2:5])
    scattered = listed.pre_eval_records(sem_pred[[4, 1]], inst_pred[[4, 1]], [4, 1])
    assert np.array_equal(scattered.cpu().numpy(), want[[4, 1]])
    back = parallel.unpack_results(rec.cpu().numpy(), C, None if multi else names)
    ev_a, _ = listed.evaluate(back, logger="silent")
    ev_b, _ = listed.evaluate(listed.pre_eval(preds, list(range(len(tiles)))), logger="silent")
    assert ev_a == ev_b
