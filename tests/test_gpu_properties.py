"""Size-independent properties of the CUDA path at BASELINE.json's full sizes and batch shapes (where the CPU oracle
would take minutes): idempotence of the relabelling, self-comparison of the metrics, the flood staying inside its
mask and keeping its markers, linearity of the accumulators over a batch, batch-vs-single-tile equality."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import tiseg_b200  # noqa: E402,F401
from tiseg_b200 import ops, synth  # noqa: E402


@pytest.fixture(scope="module")
def dist_batch():
    tiles = [synth.tile_dist(2, 40 + j) for j in range(3)]
    return {k: np.stack([t[k] for t in tiles]) for k in ("dist_logit", "gt_inst", "gt_sem", "sem_logit")}


def test_label_is_idempotent_and_raster_ordered(dist_batch):
    lab, n = ops.label(dist_batch["gt_inst"], connectivity=2, return_num=True)
    again, n2 = ops.label(lab, connectivity=2, return_num=True)
    assert np.array_equal(lab, again) and np.array_equal(n, n2)
    for j in range(len(lab)):
        flat = lab[j].ravel()
        first = np.full(int(n[j]) + 1, flat.size, np.int64)
        np.minimum.at(first, flat, np.arange(flat.size))
        assert np.all(np.diff(first[1:]) > 0), "ids must follow the raster order of each component's first pixel"
    ri = ops.re_instance(dist_batch["gt_inst"])
    assert np.array_equal(ops.re_instance(ri), ri)


def test_metrics_of_a_map_against_itself(dist_batch):
    inst = ops.postproc_dist(dist_batch["dist_logit"])
    aji, pq = ops.pair_metrics_bin(inst, inst)
    k = np.array([len(np.unique(ops.label(m, connectivity=2))) - 1 for m in inst], np.float64)
    area = (inst > 0).reshape(len(inst), -1).sum(1).astype(np.float64)
    assert np.array_equal(aji[:, 0], area) and np.array_equal(aji[:, 1], area)          # inter == union == foreground
    assert np.array_equal(pq[:, 0], k) and not pq[:, 1:3].any() and np.array_equal(pq[:, 3], k)   # every IoU is exactly 1
    counts, valid = ops.sem_counts(dist_batch["gt_sem"], dist_batch["gt_sem"], 2)
    assert not counts[:, 1:3].any() and np.array_equal(counts[:, 0], counts[:, 3]) and np.all(valid == 1000 * 1000)


def test_flood_stays_in_mask_and_keeps_markers(dist_batch):
    inst, mk, ws = ops.postproc_dist(dist_batch["dist_logit"], debug=True)
    mask = np.clip(dist_batch["dist_logit"], 0, 255).astype(np.int32) > 0
    assert not ws[~mask].any() and not inst[~mask].any()
    assert np.array_equal(ws[mk > 0], mk[mk > 0])                     # seeds keep their labels
    assert np.all((ws > 0) == mask) or ((ws > 0) & ~mask).sum() == 0  # flooded pixels are mask pixels
    for j in range(len(inst)):                                         # every flooded region holds exactly one marker
        assert set(np.unique(ws[j])) - {0} == set(np.unique(mk[j])) - {0}
    # the watershed-line pass only removes pixels, and what remains keeps consecutive ids of surviving regions
    assert not ((inst > 0) & (ws == 0)).any()


def test_batch_equals_tiles_and_accumulators_are_linear(dist_batch):
    whole = ops.postproc_dist(dist_batch["dist_logit"])
    aji_b, pq_b = ops.pair_metrics_bin(whole, dist_batch["gt_inst"])
    for j in range(len(whole)):
        one = ops.postproc_dist(dist_batch["dist_logit"][j])
        assert np.array_equal(one, whole[j])
        a, p = ops.pair_metrics_bin(one, dist_batch["gt_inst"][j])
        assert np.array_equal(a, aji_b[j]) and np.array_equal(p, pq_b[j])
    # a batch in a different order gives the same per-tile records (no cross-tile state)
    perm = [2, 0, 1]
    a2, p2 = ops.pair_metrics_bin(whole[perm], dist_batch["gt_inst"][perm])
    assert np.array_equal(a2, aji_b[perm]) and np.array_equal(p2, pq_b[perm])


def test_conic_shaped_batch_properties():
    """512 tiles of 256^2 with 7 classes in one call (the CoNIC sweep's batch shape)."""
    base = [synth.tile_unet(5, j, 256, 256, 7) for j in range(8)]
    lg = np.stack([base[i % 8]["sem_logit"][None] for i in range(512)])
    cls = ops.softmax_argmax(lg)
    sem, inst = ops.postproc_unet(cls, 6, 1, None)
    assert np.array_equal(sem[:8], sem[8:16]) and np.array_equal(inst[:8], inst[504:512])     # replicas agree
    assert np.array_equal((sem > 0), (inst > 0))
    r = ops.pair_metrics_multiclass(inst, sem, inst, sem, 7)
    assert np.array_equal(r["aji"][..., 0], r["aji"][..., 1])          # self-comparison: inter == union per class
    assert not r["pq"][:, 1:, 1:3].any()                               # (slot 0 counts id 0 on both sides: inst_metrics.py:249-252)
    assert np.array_equal(r["bin_aji"][:, 0], (inst > 0).reshape(512, -1).sum(1).astype(np.float64))
