"""CPU oracle: evaluation path (TEST INFRASTRUCTURE ONLY).

Restates ``tiseg/utils/inst_metrics.py``, ``tiseg/utils/sem_metrics.py`` and
``tiseg/datasets/utils/instance_semantic.py``.  Two forms of the GT x pred pair statistics:

* ``literal=True`` follows the reference's algorithm step by step (one full-image uint8 mask per
  id, bounding-box crops per overlapping pair; inst_metrics.py:25-67, 160-195) and therefore has
  the reference's cost profile — this is what the CPU baseline times;
* ``literal=False`` gets the same integers from one ``np.unique`` over pair keys — used by the
  big randomised tests.  Both are checked against golden vectors produced by the reference's own
  module (tests/golden/make_golden.py).
"""
import numpy as np

from .skimage_port import label as sk_label


# --------------------------------------------------------------------------- A15 / A18 helpers
def re_instance(instance_map):
    """instance_semantic.py:5-15: sorted unique non-zero ids -> 1..K (int32)."""
    ids = np.unique(instance_map)
    ids = ids[ids != 0]
    lut_in = np.searchsorted(ids, instance_map)
    lut_in = np.clip(lut_in, 0, max(len(ids) - 1, 0))
    out = np.zeros(instance_map.shape, np.int32)
    if len(ids):
        hit = ids[lut_in] == instance_map
        out[hit] = (lut_in[hit] + 1).astype(np.int32)
    return out


def assign_sem_class_to_insts(inst_seg, sem_seg, num_classes):
    """instance_semantic.py:68-93: instance -> class with most pixels among classes >= 1 (first
    max; class 0 when no such pixel or for id 0).  dict{class: [ids...]} in first-seen order."""
    ids = [int(v) for v in np.unique(inst_seg)]
    if 0 not in ids:
        ids.insert(0, 0)
    out = {}
    for i in ids:
        cnt = np.bincount(sem_seg[inst_seg == i].astype(np.int64).ravel(), minlength=num_classes)[:num_classes]
        cls = int(np.argmax(cnt[1:]) + 1) if (i != 0 and cnt[1:].sum() > 0) else 0
        out.setdefault(cls, []).append(i)
    return out


# --------------------------------------------------------------------------- pair statistics
def _bbox(m):
    """misc.py:113-131 (max indices are exclusive)."""
    r = np.where(np.any(m, axis=1))[0]
    c = np.where(np.any(m, axis=0))[0]
    return r[0], r[-1] + 1, c[0], c[-1] + 1


def _pairs_literal(pred, gt):
    """inst_metrics.py:15-67 / 145-195: dense [Ng, Np] intersection and 'total' (= |g| + |p|)
    matrices, filled only for overlapping pairs, from per-id full-image masks."""
    p_ids = list(np.unique(pred))
    g_ids = list(np.unique(gt))
    if 0 not in p_ids:
        p_ids.insert(0, 0)
    if 0 not in g_ids:
        g_ids.insert(0, 0)
    p_masks = [(pred == p).astype(np.uint8) for p in p_ids]
    g_masks = [(gt == g).astype(np.uint8) for g in g_ids]
    inter = np.zeros([len(g_ids) - 1, len(p_ids) - 1], np.float64)
    total = np.zeros_like(inter)
    for g in g_ids[1:]:
        gm = g_masks[g]
        r0, r1, c0, c1 = _bbox(gm)
        over = np.unique(pred[r0:r1, c0:c1][gm[r0:r1, c0:c1] > 0])
        for p in over:
            if p == 0:
                continue
            pm = p_masks[p]
            q0, q1, d0, d1 = _bbox(pm)
            a, b, c, d = min(r0, q0), max(r1, q1), min(c0, d0), max(c1, d1)
            gc, pc = gm[a:b, c:d], pm[a:b, c:d]
            total[g - 1, p - 1] = (gc + pc).sum()
            inter[g - 1, p - 1] = (gc * pc).sum()
    area_p = np.array([m.sum() for m in p_masks[1:]], np.float64)
    area_g = np.array([m.sum() for m in g_masks[1:]], np.float64)
    return inter, total, area_g, area_p


def _pairs_fast(pred, gt):
    n_p, n_g = int(pred.max()), int(gt.max())
    area_p = np.bincount(pred.ravel(), minlength=n_p + 1)[1:].astype(np.float64)
    area_g = np.bincount(gt.ravel(), minlength=n_g + 1)[1:].astype(np.float64)
    inter = np.zeros([n_g, n_p], np.float64)
    both = (pred > 0) & (gt > 0)
    if both.any():
        key, cnt = np.unique((gt[both].astype(np.int64) - 1) * n_p + (pred[both].astype(np.int64) - 1),
                             return_counts=True)
        inter.ravel()[key] = cnt
    total = np.where(inter > 0, area_g[:, None] + area_p[None, :], 0.0)
    return inter, total, area_g, area_p


def pre_eval_bin_aji(inst_pred, inst_gt, literal=True):
    """inst_metrics.py:10-92 -> (overall_inter, overall_union)."""
    pred = sk_label(inst_pred.copy())
    gt = sk_label(inst_gt.copy())
    inter, total, area_g, area_p = (_pairs_literal if literal else _pairs_fast)(pred, gt)
    union = total - inter
    iou = inter / (union + 1.0e-6)
    if iou.shape[0] * iou.shape[1] == 0:
        return 0., 0.
    best_p = np.argmax(iou, axis=1)
    best = np.max(iou, axis=1)
    g_sel = np.nonzero(best > 0.0)[0]
    p_sel = best_p[g_sel]
    o_inter = inter[g_sel, p_sel].sum()
    o_union = union[g_sel, p_sel].sum()
    g_used = np.zeros(len(area_g), bool)
    g_used[g_sel] = True
    p_used = np.zeros(len(area_p), bool)
    p_used[p_sel] = True
    for a in area_g[~g_used]:
        o_union += a
    for a in area_p[~p_used]:
        o_union += a
    return o_inter, o_union


def pre_eval_bin_pq(inst_pred, inst_gt, match_iou=0.5, literal=True):
    """inst_metrics.py:138-229 (match_iou >= 0.5 branch; Hungarian branch unreachable with the
    default) -> (tp, fp, fn, iou_sum)."""
    assert match_iou >= 0.5, "oracle covers the reference default branch only"
    pred = sk_label(inst_pred.copy())
    gt = sk_label(inst_gt.copy())
    inter, total, area_g, area_p = (_pairs_literal if literal else _pairs_fast)(pred, gt)
    with np.errstate(invalid="ignore", divide="ignore"):
        iou = np.where(inter > 0, inter / (total - inter), 0.0)
    iou[iou <= match_iou] = 0.0
    pg, pp = np.nonzero(iou)
    paired = iou[pg, pp]
    tp = len(pg)
    fp = len(area_p) - len(set(pp.tolist()))
    fn = len(area_g) - len(set(pg.tolist()))
    return tp, fp, fn, paired.sum()


def _class_map(inst, id_list):
    """inst_metrics.py:115-118: ids of one class renumbered 1..k in list order (int32)."""
    out = np.zeros(inst.shape, np.int32)
    for idx, i in enumerate(id_list):
        out = out + (inst == i).astype(np.int32) * (idx + 1)
    return out


def pre_eval_aji(inst_pred, inst_gt, pred_ids_per_class, gt_ids_per_class, num_classes,
                 reduce_zero_label=True, literal=True):
    """inst_metrics.py:95-135 -> (inter[C-1], union[C-1]) float32."""
    inter = np.zeros(num_classes, np.float32)
    union = np.zeros(num_classes, np.float32)
    for s in set(list(pred_ids_per_class.keys()) + list(gt_ids_per_class.keys())):
        in_p, in_g = s in pred_ids_per_class, s in gt_ids_per_class
        if s == 0:
            union[0] += sum(np.sum(inst_pred == i) for i in pred_ids_per_class[0] if i != 0)
            union[0] += sum(np.sum(inst_gt == i) for i in gt_ids_per_class[0] if i != 0)
        elif in_p and in_g:
            r = pre_eval_bin_aji(_class_map(inst_pred, pred_ids_per_class[s]),
                                 _class_map(inst_gt, gt_ids_per_class[s]), literal)
            inter[s] += r[0]
            union[s] += r[1]
        elif in_p:
            union[s] += sum(np.sum(inst_pred == i) for i in pred_ids_per_class[s] if i != 0)
        else:
            union[s] += sum(np.sum(inst_gt == i) for i in gt_ids_per_class[s] if i != 0)
    return (inter[1:], union[1:]) if reduce_zero_label else (inter, union)


def pre_eval_pq(inst_pred, inst_gt, pred_ids_per_class, gt_ids_per_class, num_classes,
                reduce_zero_label=True, literal=True):
    """inst_metrics.py:232-280 -> (tp, fp, fn, iou)[C-1] float32."""
    tp, fp, fn, iou = (np.zeros(num_classes, np.float32) for _ in range(4))
    for s in set(list(pred_ids_per_class.keys()) + list(gt_ids_per_class.keys())):
        in_p, in_g = s in pred_ids_per_class, s in gt_ids_per_class
        if s == 0:
            fp[0] += len(pred_ids_per_class[0])
            fn[0] += len(gt_ids_per_class[0])
        elif in_p and in_g:
            r = pre_eval_bin_pq(_class_map(inst_pred, pred_ids_per_class[s]),
                                _class_map(inst_gt, gt_ids_per_class[s]), literal=literal)
            tp[s] += r[0]
            fp[s] += r[1]
            fn[s] += r[2]
            iou[s] += r[3]
        elif in_p:
            fp[s] += len(pred_ids_per_class[s])
        else:
            fn[s] += len(gt_ids_per_class[s])
    if reduce_zero_label:
        return tp[1:], fp[1:], fn[1:], iou[1:]
    return tp, fp, fn, iou


# --------------------------------------------------------------------------- A19 semantic counts
def pre_eval_all_semantic_metric(pred_label, target_label, num_classes, ignore_index=255,
                                 reduce_zero_label=True):
    """sem_metrics.py:16-53 -> (TP, TN, FP, FN, Pred, GT) float32 arrays (torch.histc bins ==
    integer class ids; values outside [0, C-1] are not counted)."""
    pred = np.asarray(pred_label).ravel()
    tgt = np.asarray(target_label).ravel()
    keep = tgt != ignore_index
    pred, tgt = pred[keep].astype(np.int64), tgt[keep].astype(np.int64)

    def hist(v):
        v = v[(v >= 0) & (v <= num_classes - 1)]
        return np.bincount(v, minlength=num_classes).astype(np.float32)

    same = pred == tgt
    tp, fp, fn = hist(tgt[same]), hist(pred[~same]), hist(tgt[~same])
    pr, gt = hist(pred), hist(tgt)
    tn = pr.sum() - (tp + fp + fn)
    res = (tp, tn, fp, fn, pr, gt)
    return tuple(r[1:] for r in res) if reduce_zero_label else res


# --------------------------------------------------------------------------- A20 reducers
def to_bin_aji(results):
    """inst_metrics.py:383-403."""
    i = sum(np.sum(r[0]) for r in results)
    u = sum(np.sum(r[1]) for r in results)
    return {"Aji": i / u}


def to_bin_pq(results):
    """inst_metrics.py:457-491."""
    tp = sum(np.sum(r[0]) for r in results)
    fp = sum(np.sum(r[1]) for r in results)
    fn = sum(np.sum(r[2]) for r in results)
    iou = sum(np.sum(r[3]) for r in results)
    dq = tp / (tp + 0.5 * fp + 0.5 * fn)
    sq = iou / (tp + 1.0e-6)
    return {"DQ": dq, "SQ": sq, "PQ": dq * sq}


def to_sem_dice(results):
    """sem_metrics.py:214-303 for metric 'Dice' (sequential float32 sums over images)."""
    tp = sum(r[0] for r in results)
    pr = sum(r[4] for r in results)
    gt = sum(r[5] for r in results)
    return {"Dice": 2 * tp / (pr + gt)}
