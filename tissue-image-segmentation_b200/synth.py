"""Seeded synthetic tiles shared by the tests, the CPU oracle baseline and ``bench.py``.

There is no dataset or checkpoint in this environment, so every workload is generated
(SURVEY.md §8d): a ground-truth instance map of touching, non-overlapping elliptical nuclei, a
perturbed "prediction" (shifted / dropped / spurious / merged nuclei, speckle noise) and the
network outputs a segmentor would have produced for it (semantic logits, distance map, HV map,
direction logits, point map).  Pure numpy/scipy host code; not part of the compute path.

seed convention: ``1000 * config_id + tile_index``.
"""
import numpy as np
from scipy import ndimage as ndi


# --------------------------------------------------------------------------- instance maps
def _ellipse_params(rng, H, W, n):
    cy = rng.uniform(0, H, n)
    cx = rng.uniform(0, W, n)
    a = rng.uniform(4, 12, n)
    b = rng.uniform(4, 12, n)
    th = rng.uniform(0, np.pi, n)
    return np.stack([cy, cx, a, b, th], axis=1)


def _paint(params, ids, H, W):
    """Paint ellipses in list order, only onto background (=> touching, never overlapping)."""
    out = np.zeros((H, W), np.int32)
    for (cy, cx, a, b, th), i in zip(params, ids):
        r = int(np.ceil(max(a, b))) + 1
        y0, y1 = max(int(cy) - r, 0), min(int(cy) + r + 1, H)
        x0, x1 = max(int(cx) - r, 0), min(int(cx) + r + 1, W)
        if y0 >= y1 or x0 >= x1:
            continue
        yy, xx = np.mgrid[y0:y1, x0:x1]
        dy, dx = yy - cy, xx - cx
        u = dx * np.cos(th) + dy * np.sin(th)
        v = -dx * np.sin(th) + dy * np.cos(th)
        m = (u / a) ** 2 + (v / b) ** 2 <= 1.0
        win = out[y0:y1, x0:x1]
        win[m & (win == 0)] = i
    return out


def default_count(H, W):
    """60 nuclei at 256^2 (CoNIC / CPM17 density), 900 at 1000^2 (MoNuSeg / CoNSeP)."""
    return max(1, int(round(900 * (H * W) / 1.0e6)))


def gt_and_pred(seed, H, W, n=None, num_classes=2, speckle=0.001):
    """Returns dict(gt_inst int32, gt_sem uint8, pred_inst int32, pred_sem uint8).

    ``*_sem`` is the class map (instance class in 1..num_classes-1, 0 background)."""
    rng = np.random.default_rng(seed)
    n = default_count(H, W) if n is None else n
    prm = _ellipse_params(rng, H, W, n)
    ids = np.arange(1, n + 1)
    cls = rng.integers(1, max(num_classes, 2), n).astype(np.uint8) if num_classes > 2 else np.ones(n, np.uint8)
    gt_inst = _paint(prm, ids, H, W)

    # prediction = perturbed copy: +-2 px shift, 5 % dropped, 5 % spurious, 5 % merged pairs
    pp = prm.copy()
    pp[:, 0] += rng.integers(-2, 3, n)
    pp[:, 1] += rng.integers(-2, 3, n)
    keep = rng.random(n) >= 0.05
    pid = ids.copy()
    merge = np.nonzero(rng.random(n) < 0.05)[0]
    for j in merge:  # give j the id of its nearest neighbour
        d = (prm[:, 0] - prm[j, 0]) ** 2 + (prm[:, 1] - prm[j, 1]) ** 2
        d[j] = np.inf
        pid[j] = pid[int(np.argmin(d))]
    n_sp = max(1, n // 20)
    sp = _ellipse_params(rng, H, W, n_sp)
    pprm = np.concatenate([pp[keep], sp], 0)
    pids = np.concatenate([pid[keep], np.arange(n + 1, n + 1 + n_sp)])
    pcls = np.concatenate([cls[keep], rng.integers(1, max(num_classes, 2), n_sp).astype(np.uint8)
                           if num_classes > 2 else np.ones(n_sp, np.uint8)])
    pred_inst = _paint(pprm, pids, H, W)

    lut_g = np.zeros(n + 1, np.uint8)
    lut_g[1:] = cls
    lut_p = np.zeros(n + 1 + n_sp, np.uint8)
    lut_p[pids] = pcls
    gt_sem = lut_g[gt_inst]
    pred_sem = lut_p[pred_inst]

    # speckle: isolated false-positive specks and pin-holes (exercise remove_small / fill_holes)
    k = int(speckle * H * W)
    if k:
        ys, xs = rng.integers(0, H, k), rng.integers(0, W, k)
        was_bg = pred_sem[ys, xs] == 0
        pred_sem[ys, xs] = np.where(was_bg, 1, 0).astype(np.uint8)
        new_ids = (n + 1 + n_sp + np.arange(k)).astype(np.int32)   # each speck is its own nucleus
        pred_inst[ys, xs] = np.where(was_bg, new_ids, 0)
    return dict(gt_inst=gt_inst, gt_sem=gt_sem, pred_inst=pred_inst, pred_sem=pred_sem)


# --------------------------------------------------------------------------- network outputs
def sem_logits(rng, cls_map, C, margin=4.0):
    """fp32 [C,H,W]: +margin on the class channel, -margin elsewhere, plus N(0,1)."""
    H, W = cls_map.shape
    lg = rng.standard_normal((C, H, W), dtype=np.float32)
    lg -= np.float32(margin)
    idx = cls_map.astype(np.int64)
    np.put_along_axis(lg, idx[None], np.take_along_axis(lg, idx[None], 0) + np.float32(2 * margin), 0)
    return lg


def chessboard_distance(inst):
    """Per-instance chessboard distance to the instance border (the DIST regression target,
    datasets/ops/distance_map.py:93 with inst_norm=False)."""
    out = np.zeros(inst.shape, np.float32)
    objs = ndi.find_objects(inst)
    H, W = inst.shape
    for i, sl in enumerate(objs, start=1):
        if sl is None:
            continue
        y0, y1 = max(sl[0].start - 2, 0), min(sl[0].stop + 2, H)
        x0, x1 = max(sl[1].start - 2, 0), min(sl[1].stop + 2, W)
        m = inst[y0:y1, x0:x1] == i
        mp = np.pad(m, 1)
        d = ndi.distance_transform_cdt(mp, metric="chessboard")[1:-1, 1:-1]
        win = out[y0:y1, x0:x1]
        win[m] = d[m]
    return out


def dist_map(rng, pred_inst, noise=0.3, smooth=1.0):
    """DIST head output: chessboard distance of the predicted nuclei + N(0, noise), passed
    through a Gaussian (a regression head's output is smooth; white noise on the truncated map
    would shatter every plateau into spurious maxima)."""
    d = chessboard_distance(pred_inst)
    d = d + rng.standard_normal(d.shape, dtype=np.float32) * np.float32(noise)
    return ndi.gaussian_filter(d, smooth).astype(np.float32)


def hv_map(rng, pred_inst, noise=0.05):
    """HoVer-Net HV head output [H,W,2] (horizontal, vertical), datasets/ops/hv_map.py recipe:
    per instance, offsets from the centre of mass scaled to [-1, 1] on each side, plus noise."""
    H, W = pred_inst.shape
    hv = np.zeros((H, W, 2), np.float32)
    objs = ndi.find_objects(pred_inst)
    for i, sl in enumerate(objs, start=1):
        if sl is None:
            continue
        m = pred_inst[sl] == i
        ys, xs = np.nonzero(m)
        if len(ys) == 0:
            continue
        dy = ys - ys.mean()
        dx = xs - xs.mean()
        for arr in (dx, dy):
            neg, pos = arr < 0, arr > 0
            if neg.any():
                arr[neg] /= -arr[neg].min()
            if pos.any():
                arr[pos] /= arr[pos].max()
        win = hv[sl]
        win[ys, xs, 0] = dx
        win[ys, xs, 1] = dy
    hv += rng.standard_normal(hv.shape, dtype=np.float32) * np.float32(noise)
    return hv


def direction_logits(rng, pred_inst, margin=4.0):
    """CDNet direction head [9,H,W]: class 0 background, 1..8 = quantised angle of the vector
    from the pixel to its instance centroid, one-hot * margin + N(0,1); and the point map
    [1,H,W] = Gaussian(sigma 2) at centroids * 255 (datasets/ops/direction_map.py:157)."""
    H, W = pred_inst.shape
    cls = np.zeros((H, W), np.int64)
    pts = np.zeros((H, W), np.float32)
    objs = ndi.find_objects(pred_inst)
    for i, sl in enumerate(objs, start=1):
        if sl is None:
            continue
        m = pred_inst[sl] == i
        ys, xs = np.nonzero(m)
        if len(ys) == 0:
            continue
        cy, cx = ys.mean(), xs.mean()
        ang = np.degrees(np.arctan2(cy - ys, cx - xs)) % 360.0
        q = (np.floor((ang + 22.5) / 45.0).astype(np.int64) % 8) + 1
        cls[sl][ys, xs] = q
        py, px = int(round(cy)) + sl[0].start, int(round(cx)) + sl[1].start
        if 0 <= py < H and 0 <= px < W:
            pts[py, px] = 1.0
    pts = ndi.gaussian_filter(pts, 2.0)
    if pts.max() > 0:
        pts = pts / pts.max() * 255.0
    return sem_logits(rng, cls, 9, margin), pts[None].astype(np.float32)


def three_class_map(inst):
    """Inside / edge class map for the CUNet / CDNet family: 1 inside, 2 = instance pixels that
    touch another label or background in their 3x3 window (edge), 0 background."""
    mx = ndi.maximum_filter(inst, size=3, mode="nearest")
    mn = ndi.minimum_filter(inst, size=3, mode="nearest")
    out = (inst > 0).astype(np.uint8)
    out[(inst > 0) & (mx != mn)] = 2
    return out


# --------------------------------------------------------------------------- per-config tiles
def tile_unet(config_id, tile_index, H=256, W=256, C=2):
    """Config 1 / 5 (UNet family): class logits + GT."""
    seed = 1000 * config_id + tile_index
    t = gt_and_pred(seed, H, W, num_classes=C)
    rng = np.random.default_rng(seed + 500000)
    t["sem_logit"] = sem_logits(rng, t["pred_sem"], C)
    return t


def tile_dist(config_id, tile_index, H=1000, W=1000):
    """Config 2 (DIST): binary sem logits + distance map + GT."""
    seed = 1000 * config_id + tile_index
    t = gt_and_pred(seed, H, W, num_classes=2)
    rng = np.random.default_rng(seed + 500000)
    t["sem_logit"] = sem_logits(rng, t["pred_sem"], 2)
    inst = np.where(t["pred_sem"] > 0, t["pred_inst"], 0)
    t["dist_logit"] = dist_map(rng, inst)
    return t


def tile_hover(config_id, tile_index, H=1000, W=1000):
    """Config 3 (HoVer-Net): 3-class type logits, foreground probability, HV map [H,W,2] + GT."""
    seed = 1000 * config_id + tile_index
    t = gt_and_pred(seed, H, W, num_classes=3)
    rng = np.random.default_rng(seed + 500000)
    t["sem_logit"] = sem_logits(rng, t["pred_sem"], 3)
    inst = np.where(t["pred_sem"] > 0, t["pred_inst"], 0)
    fg = sem_logits(rng, (inst > 0).astype(np.uint8), 2)
    e = np.exp(fg - fg.max(0, keepdims=True))
    t["fore_map"] = np.ascontiguousarray((e / e.sum(0, keepdims=True))[1].astype(np.float32))
    t["hv_map"] = ndi.gaussian_filter(hv_map(rng, inst), (1.0, 1.0, 0)).astype(np.float32)
    return t


def tile_cdnet(config_id, tile_index, H=1000, W=1000, T=1):
    """Config 4 (CDNet): 3-class (bg / inside / edge) logits, 9-class direction logits, point map, per TTA
    variant, + GT."""
    seed = 1000 * config_id + tile_index
    t = gt_and_pred(seed, H, W, num_classes=2)
    rng = np.random.default_rng(seed + 500000)
    inst = np.where(t["pred_sem"] > 0, t["pred_inst"], 0)
    tc = three_class_map(inst)
    t["sem_logit"] = np.stack([sem_logits(rng, tc, 3) for _ in range(T)])
    dirs, pts = zip(*[direction_logits(rng, inst) for _ in range(T)])
    t["dir_logit"] = np.stack(dirs)
    t["point_logit"] = np.stack(pts)
    return t
