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
