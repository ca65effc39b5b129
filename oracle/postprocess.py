"""CPU oracle: per-model test-time post-processing (TEST INFRASTRUCTURE ONLY).

Restates, function by function, what the reference's segmentors do between the network logits
and ``{'sem_pred', 'inst_pred'}``; every function cites the reference lines it follows.
scipy.ndimage and OpenCV are called exactly as the reference calls them; the scikit-image
functions come from ``skimage_port`` (C restatement) and the thin wrappers below.
"""
import math
import warnings

import numpy as np
from scipy import ndimage as ndi

from .skimage_port import label as sk_label
from .skimage_port import watershed as sk_watershed
from .skimage_port import reconstruction_erosion, align_foreground

try:  # OpenCV is only needed by the HoVer-Net path
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


# --------------------------------------------------------------------------- skimage wrappers
def disk(radius):
    """skimage.morphology.disk: {(x, y): x^2 + y^2 <= r^2} as uint8."""
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    return ((X * X + Y * Y) <= radius * radius).astype(np.uint8)


def square(width):
    return np.ones((width, width), np.uint8)


def dilation(image, selem):
    """skimage.morphology.dilation == ndi.grey_dilation(image, footprint=selem) (symmetric selem)."""
    return ndi.grey_dilation(image, footprint=np.asarray(selem)[::-1, ::-1])


def erosion(image, selem):
    """skimage.morphology.erosion == ndi.grey_erosion(image, footprint=selem)."""
    return ndi.grey_erosion(image, footprint=np.asarray(selem))


def remove_small_objects(ar, min_size=64, connectivity=1):
    """skimage.morphology.remove_small_objects (0.18.3 semantics).

    bool input: components of the given connectivity (default 1 -> 4-neighbourhood) found with
    ``ndi.label``; int input: the values themselves are the component ids.  Components with
    ``size < min_size`` are zeroed.  Returns a copy.
    """
    out = np.array(ar, copy=True)
    if min_size == 0:
        return out
    if out.dtype == bool:
        selem = ndi.generate_binary_structure(ar.ndim, connectivity)
        ccs = np.zeros(ar.shape, np.int32)
        ndi.label(ar, selem, output=ccs)
    else:
        ccs = out
    sizes = np.bincount(ccs.ravel())
    too_small = sizes < min_size
    out[too_small[ccs]] = 0
    return out


# --------------------------------------------------------------------------- A1 softmax / argmax
def softmax(logits, axis=0):
    """fp32 softmax over the channel axis (torch ``F.softmax``: exp(x - max) / sum)."""
    x = np.asarray(logits, np.float32)
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m, dtype=np.float32)
    return (e / e.sum(axis=axis, keepdims=True, dtype=np.float32)).astype(np.float32)


def softmax_tta_mean(logit_list):
    """base.py:321-336: softmax each TTA variant, then ``sum(list) / len(list)`` in fp32."""
    acc = None
    for lg in logit_list:
        p = softmax(lg, axis=0)
        acc = p if acc is None else (acc + p).astype(np.float32)
    return (acc / np.float32(len(logit_list))).astype(np.float32)


def split_stitch(patches, H, W, window, overlap):
    """base.py:255-295 ``split_inference`` with the network replaced by its outputs: ``patches`` [M, C, window,
    window] in the order the double loop visits the windows.  -> [C, H, W]."""
    st = window - overlap
    pad_h = st - (H - window) % st if H - window > 0 else window - H
    pad_w = st - (W - window) % st if W - window > 0 else window - W
    H1, W1 = pad_h + H, pad_w + W
    C = patches.shape[1]
    canvas = np.zeros((C, H1, W1), patches.dtype)
    k = 0
    for i in range(0, H1 - overlap, st):
        r_s = i + overlap // 2 if i > 0 else 0
        r_e = i + window - overlap // 2 if i + window < H1 else H1
        for j in range(0, W1 - overlap, st):
            c_s = j + overlap // 2 if j > 0 else 0
            c_e = j + window - overlap // 2 if j + window < W1 else W1
            canvas[:, r_s:r_e, c_s:c_e] = patches[k][:, r_s - i:r_e - i, c_s - j:c_e - j]
            k += 1
    assert k == len(patches), (k, len(patches))
    return canvas[:, (H1 - H) // 2:(H1 - H) // 2 + H, (W1 - W) // 2:(W1 - W) // 2 + W]


def reverse_tta_transform(x, rotate_degree, flip_direction):
    """base.py:365-381 on a [C, H, W] array (numpy's rot90 / flip are torch's)."""
    k = 4 - (rotate_degree // 90) % 4
    if flip_direction == "horizontal":
        x = np.flip(x, axis=-1)
    if flip_direction == "vertical":
        x = np.flip(x, axis=-2)
    if flip_direction == "diagonal":
        x = np.flip(x, axis=(-2, -1))
    return np.ascontiguousarray(np.rot90(x, k=k, axes=(-2, -1)))


def argmax_classes(prob):
    """``sem_logit.argmax(dim=1)``: first maximum wins."""
    return np.argmax(prob, axis=0).astype(np.int64)


# --------------------------------------------------------------------------- A2 UNet family
def unet_family_postprocess(pred, radius=1, edge_id=None):
    """unet.py:71-93 (radius 1), micronet.py:185-207; with ``edge_id`` (= num_classes) the
    variants cunet.py:70-93, cdnet.py:96-119, fullnet.py:190-213, cmicronet.py:186-209
    (radius 3) which first zero the edge class in place.

    For each class id present (ascending, 0 skipped): mask -> binary_fill_holes ->
    remove_small_objects(5) -> measure.label -> dilation(disk(radius)) -> + cur -> overwrite
    into inst_pred; cur += len(unique(dilated)); sem_pred[dilated > 0] = id.
    """
    if edge_id is not None:
        pred[pred == edge_id] = 0
    inst_pred = np.zeros(pred.shape, np.int32)
    sem_pred = np.zeros(pred.shape, np.uint8)
    cur = 0
    for sem_id in np.unique(pred):
        if sem_id == 0:
            continue
        m = ndi.binary_fill_holes(pred == sem_id)
        m = remove_small_objects(m, 5)
        lab = dilation(sk_label(m), disk(radius))
        hit = lab > 0
        lab[hit] += cur
        inst_pred[hit] = 0
        inst_pred += lab.astype(np.int32)
        cur += len(np.unique(lab))
        sem_pred[hit] = sem_id
    return sem_pred, inst_pred


def dcan_postprocess(cell_pred, cont_pred, radius=3):
    """dcan.py:193-217: contour prediction splits cells (in place), then the UNet-family loop."""
    cell_pred[cont_pred > 0] = 0
    return unet_family_postprocess(cell_pred, radius=radius)


# --------------------------------------------------------------------------- A8 DIST
def _h_reconstruction_erosion(prob_img, h, literal=True):
    """dist.py:43-57.  ``literal`` keeps the reference's per-pixel ``np.vectorize`` (its real cost
    on the CPU baseline); the fast form is value-identical."""
    if literal:
        # int(x): numpy >= 2 would wrap uint8(255) + 1; the reference's pinned numpy 1.20 promotes
        shifted = np.vectorize(lambda x, lamb=h: min(255, int(x) + lamb))(prob_img)
    else:
        shifted = np.minimum(255, prob_img.astype(np.float64) + h)
    return reconstruction_erosion(shifted, prob_img).astype(np.ubyte)


def _find_maxima(img, mask, literal=True):
    """dist.py:60-71 with convertuint8=False, inverse=False."""
    res = _h_reconstruction_erosion(img, 1, literal) - img
    res[mask == 0] = 0
    return res


def _arrange_label(mat):
    """dist.py:101-111: relabel with the most frequent value as background."""
    val, counts = np.unique(mat, return_counts=True)
    bg = val[np.argmax(counts)]
    return sk_label(mat, background=bg)


def _generate_wsl(ws):
    """dist.py:83-98: 255 where the 3x3 window of a non-zero pixel holds >= 2 different
    non-zero labels."""
    se = square(3)
    ero = ws.copy()
    ero[ero == 0] = ero.max() + 1
    ero = erosion(ero, se)
    ero[ws == 0] = 0
    grad = dilation(ws, se) - ero
    grad[ws == 0] = 0
    grad[grad > 0] = 255
    return grad.astype(np.uint8)


def dist_dynamic_watershed(p_img, lamb=0.0, p_thresh=0.5, literal=True):
    """dist.py:114-129 ``dynamic_watershed_alias``."""
    b_img = (p_img > p_thresh) + 0
    probs_inv = 255 - p_img.astype(np.uint8)
    hrecons = _h_reconstruction_erosion(probs_inv, lamb, literal)
    markers = sk_label(_find_maxima(hrecons, b_img, literal))
    ws = sk_watershed(hrecons, markers, mask=b_img)
    arranged = _arrange_label(ws)
    wsl = _generate_wsl(arranged)
    arranged[wsl > 0] = 0
    return arranged


def dist_postprocess(sem_pred, dist_logit, literal=True):
    """dist.py:275-284: clip to [0, 255], truncate to int32, dynamic watershed (lambda = 0.0,
    threshold 0.5).  ``sem_pred`` passes through untouched."""
    d = np.copy(dist_logit)
    d[d > 255] = 255
    d[d < 0] = 0
    d = d.astype("int32")
    return sem_pred, dist_dynamic_watershed(d, 0.0, 0.5, literal)


# --------------------------------------------------------------------------- A11 HoVer-Net
def hover_post_proc(fore_map, hv_map, fx=1, scale_factor=1):
    """hovernet.py:283-365 with OpenCV / scipy called as the reference does."""
    assert cv2 is not None, "OpenCV is required for the HoVer-Net oracle"
    raw_h, raw_w = hv_map.shape[:2]
    fore_map = cv2.resize(fore_map, (0, 0), fx=scale_factor, fy=scale_factor)
    hv_map = cv2.resize(hv_map, (0, 0), fx=scale_factor, fy=scale_factor)
    h_raw, v_raw = hv_map[:, :, 0], hv_map[:, :, 1]

    blb = np.array(fore_map >= 0.5, dtype=np.int32)
    blb = ndi.label(blb)[0]
    blb = remove_small_objects(blb, min_size=10)
    blb[blb > 0] = 1

    def mm(x):
        return cv2.normalize(x, None, alpha=0, beta=1, norm_type=cv2.NORM_MINMAX, dtype=cv2.CV_32F)

    h_dir, v_dir = mm(h_raw), mm(v_raw)
    ksize = int((20 * fx) + 1)
    obj_size = math.ceil(10 * (fx ** 2))
    sobelh = 1 - mm(cv2.Sobel(h_dir, cv2.CV_64F, 1, 0, ksize=ksize))
    sobelv = 1 - mm(cv2.Sobel(v_dir, cv2.CV_64F, 0, 1, ksize=ksize))

    overall = np.maximum(sobelh, sobelv)
    overall = overall - (1 - blb)
    overall[overall < 0] = 0
    dist = (1.0 - overall) * blb
    dist = -cv2.GaussianBlur(dist, (3, 3), 0)
    overall = np.array(overall >= 0.4, dtype=np.int32)

    marker = blb - overall
    marker[marker < 0] = 0
    marker = ndi.binary_fill_holes(marker).astype("uint8")
    kernel = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5))
    marker = cv2.morphologyEx(marker, cv2.MORPH_OPEN, kernel)
    marker = ndi.label(marker)[0]
    marker = remove_small_objects(marker, min_size=obj_size)
    out = sk_watershed(dist, markers=marker, mask=blb)
    out = cv2.resize(out, (raw_w, raw_h), interpolation=cv2.INTER_NEAREST)
    return out, dict(blb=blb, dist=dist, marker=marker, overall=overall)


# --------------------------------------------------------------------------- A12 CDNet DDM
_DIR9 = np.array([[0, 0], [0, -1], [-1, -1], [-1, 0], [-1, 1], [0, 1], [1, 1], [1, 0], [1, -1]], np.float32)


def direction_differential_map(dir_map, direction_classes=9):
    """direct_diff_map.py:95-167 for 9 direction classes, fp32 arithmetic like the torch code:
    label -> 2-vector, cosine similarity with the 8 circularly shifted neighbours, min over the 8,
    background -> 1, ``1 - round``, min-max normalise (all-zero map returned as is)."""
    assert direction_classes == 9
    dm = np.asarray(dir_map)
    va = _DIR9[dm]  # H, W, 2
    a0, a1 = va[..., 0], va[..., 1]
    shifts = [(1, 0), (1, 1), (0, 1), (-1, 1), (-1, 0), (-1, -1), (0, -1), (1, -1)]
    na = np.sqrt(a0 * a0 + a1 * a1, dtype=np.float32)
    best = None
    for sv, sh in shifts:
        f0 = np.roll(a0, (sv, sh), axis=(0, 1))
        f1 = np.roll(a1, (sv, sh), axis=(0, 1))
        num = a0 * f0 + a1 * f1
        den = na * np.sqrt(f0 * f0 + f1 * f1, dtype=np.float32) + np.float32(0.000001)
        c = (num / den).astype(np.float32)
        best = c if best is None else np.minimum(best, c)
    best[dm == 0] = 1
    lv = (1 - np.round(best)).astype(np.float32)
    mx, mn = lv.max(), lv.min()
    if mx == 0:
        return lv
    with np.errstate(invalid="ignore", divide="ignore"):
        return ((lv - mn) / (mx - mn)).astype(np.float32)


def cdnet_inference_tail(sem_logit_list, dir_logit_list, point_logit_list, if_ddm=True):
    """cdnet.py:175-217 (after the CNN): softmax + TTA mean of sem; point mean; per variant
    ``dir[:,0] *= sem[:,0]`` -> argmax -> DDM; mean DDM; optional ``_ddm_enhencement``
    (cdnet.py:354-367).  Inputs are lists of [C,H,W] fp32 logits.  Returns (sem_prob[C,H,W],
    dir_map of the first variant, dd_map)."""
    sem = softmax_tta_mean(sem_logit_list)
    point = None
    for p in point_logit_list:
        point = p.astype(np.float32) if point is None else (point + p).astype(np.float32)
    point = (point / np.float32(len(point_logit_list))).astype(np.float32)
    dd_sum, dir_maps = None, []
    for dl in dir_logit_list:
        d = softmax(dl, axis=0)
        d[0] = d[0] * sem[0]
        dir_map = np.argmax(d, axis=0)
        dir_maps.append(dir_map)
        dd = direction_differential_map(dir_map, 9)
        dd_sum = dd if dd_sum is None else (dd_sum + dd).astype(np.float32)
    dd_map = (dd_sum / np.float32(len(dir_logit_list))).astype(np.float32)
    if if_ddm:
        pl = point[0]
        with np.errstate(invalid="ignore", divide="ignore"):
            point_map = (pl / pl.max()) > 0.2
        dd2 = (dd_map - (dd_map * point_map.astype(np.float32))).astype(np.float32)
        sem = sem.copy()
        sem[-1] = ((sem[-1] + dd2).astype(np.float32) * (np.float32(1) + dd2)).astype(np.float32)
    return sem, dir_maps[0], dd_map


def ddm_enhancement(sem_prob, dd_map, point, mode=0):
    """``_ddm_enhencement``: mode 0 = cdnet.py:354-367, mode 1 = multi_task_cdnet.py:548-564.  sem_prob [C,H,W] fp32
    probabilities, dd_map [H,W], point [H,W] (channel 0 of the TTA-mean point map).  fp32, the reference's operation order;
    returns a new array."""
    sem = np.array(sem_prob, np.float32, copy=True)
    one = np.float32(1)
    with np.errstate(invalid="ignore", divide="ignore"):
        if mode == 0:
            point_map = (point / point.max()) > np.float32(0.2)
            dd2 = (dd_map - (dd_map * point_map.astype(np.float32))).astype(np.float32)
            sem[-1] = ((sem[-1] + dd2).astype(np.float32) * (one + dd2)).astype(np.float32)
        else:
            dist_map = (point + np.float32(0.2)).astype(np.float32)
            t = (dist_map / dist_map.max()).astype(np.float32)
            fprob = (t * t).astype(np.float32)                         # torch ``** 2``
            fmap = fprob > np.float32(0.6)
            w0 = (one - fprob).astype(np.float32)
            dd1 = (dd_map - (dd_map * fmap.astype(np.float32))).astype(np.float32)
            e = ((sem[-1] * (one + dd1)).astype(np.float32) * w0).astype(np.float32)
            e[e >= one] = np.float32(0.95)
            sem[-1] = e                                                # (:562 compares a bool map with 0.8: never true)
    return sem


def regression_dir_map(angle_rad, background, num_angles=8):
    """multi_task_cdnet.py:304-315: the regression head's angle (radians, fp32) -> direction classes.  Clamp to [0, 2 pi],
    degrees, (180, 360] -> (-180, 0], background 0, class = 1 + bin of align_angle (direction_calculation.py:60-73;
    angle_to_vector + vector_to_label snap to the bin centre and bin again, which is the same bin), background class 0."""
    a = np.array(angle_rad, np.float32, copy=True)
    a[a < 0] = 0
    a[a > np.float32(2 * np.pi)] = np.float32(2 * np.pi)
    deg = (a * np.float32(180)) / np.float32(np.pi)
    deg[deg > 180] -= 360
    deg[background] = 0
    d = direction_bins(deg, num_angles)
    d[background] = -1
    return d + 1


def mtcdnet_inference_tail(tc_logit_list, sem_logit_list, dir_logit_list, point_logit_list, if_ddm=True, use_regression=False,
                           num_angles=8):
    """multi_task_cdnet.py:262-330 after the CNN: softmax + TTA mean of tc and sem; point mean;
    per variant ``dir[:,0] *= tc[:,0]`` -> argmax (or, with use_regression, the angle head -> classes, :304-315) -> DDM;
    mean DDM; optional ``_ddm_enhencement`` (:548-564) on tc.
    Inputs: lists of [C,H,W] fp32 logits (already reverse-transformed).  Returns (tc_prob, sem_prob, dir_map of the first
    variant, dd_map)."""
    tc = softmax_tta_mean(tc_logit_list)
    sem = softmax_tta_mean(sem_logit_list)
    point = None
    for p in point_logit_list:
        point = p.astype(np.float32) if point is None else (point + p).astype(np.float32)
    point = (point / np.float32(len(point_logit_list))).astype(np.float32)
    dd_sum, dir_maps = None, []
    for dl in dir_logit_list:
        if use_regression:
            dir_map = regression_dir_map(dl[0], np.argmax(tc, axis=0) == 0, num_angles)
        else:
            d = softmax(dl, axis=0)
            d[0] = d[0] * tc[0]
            dir_map = np.argmax(d, axis=0)
        dir_maps.append(dir_map)
        dd = direction_differential_map(dir_map, num_angles + 1)
        dd_sum = dd if dd_sum is None else (dd_sum + dd).astype(np.float32)
    dd_map = (dd_sum / np.float32(len(dir_logit_list))).astype(np.float32)
    if if_ddm:
        tc = ddm_enhancement(tc, dd_map, point[0], mode=1)
    return tc, sem, dir_maps[0], dd_map


def tta_plain_mean(reversed_list):
    """``sum(list) / len(list)`` of already reverse-transformed fp32 maps (the regression heads: dist.py:398-406)."""
    acc = None
    for x in reversed_list:
        acc = np.asarray(x, np.float32) if acc is None else (acc + x).astype(np.float32)
    return (acc / np.float32(len(reversed_list))).astype(np.float32)


def three_class_gt(sem_gt_w_bound, num_classes):
    """multi_task_cdnet_debug.py:164-168: 0 background, 1 inside, boundary label (== num_classes) -> 2."""
    tc = np.array(sem_gt_w_bound, copy=True)
    tc[(tc != 0) * (tc != num_classes)] = 1
    tc[tc > 1] = 2
    return tc


# --------------------------------------------------------------------------- A13 multi-task
def multitask_postprocess(inner_pred, sem_pred, variant="unet"):
    """multi_task_unet.py:84-106, multi_task_cunet.py:86-108 (variant 'cunet': tc map with edge
    class 2 zeroed first, returns canvas), multi_task_cdnet.py:222-243 (variant 'cdnet': same but
    returns the RAW sem_pred).  sem canvas: per class remove_small_objects(5) THEN
    binary_fill_holes; instances: measure.label(connectivity=1) + align_foreground(.., 20)."""
    canvas = np.zeros(sem_pred.shape, np.uint8)
    for sem_id in np.unique(sem_pred):
        if sem_id == 0:
            continue
        m = remove_small_objects(sem_pred == sem_id, 5)
        m = ndi.binary_fill_holes(m)
        canvas[m > 0] = sem_id
    bin_pred = inner_pred.copy()
    if variant in ("cunet", "cdnet"):
        bin_pred[bin_pred == 2] = 0
    inst = sk_label(bin_pred, connectivity=1)
    inst = align_foreground(np.ascontiguousarray(inst), canvas > 0, 20)
    return (sem_pred if variant == "cdnet" else canvas), inst


warnings.filterwarnings("ignore", category=DeprecationWarning, module="scipy")


# --------------------------------------------------------------------------- mudslide_watershed
_DIRX = [0, 0, -1, -1, -1, 0, 1, 1, 1]          # postprocess.py:37-38: (row, col) offset of direction k = 1..8
_DIRY = [0, -1, -1, 0, 1, 1, 1, 0, -1]


def graph_degree(graph):
    """postprocess.py:12-28 ``get_graph_degree``: every pixel with a direction k adds one to the pixel BEHIND it
    (position minus the direction vector)."""
    n, m = graph.shape
    degree = np.zeros((n, m), np.int16)
    for i in range(n):
        for j in range(m):
            k = int(graph[i, j])
            if k > 0:
                nx, ny = i - _DIRX[k], j - _DIRY[k]
                if 0 <= nx < n and 0 <= ny < m:
                    degree[nx, ny] += 1
    return degree


def mudslide_prepare(seg, dir_graph, contour, degree):
    """postprocess.py:31-120 ``prepare`` (plain-Python restatement of the numba kernel, same visiting order).
    Mutates ``seg`` and ``dir_graph``; returns ``level``."""
    h, w = seg.shape
    vis = np.zeros((h, w), np.int16)
    level = np.ones((h, w), np.int16)
    hfa = np.zeros((h, w), np.int16)
    seg[degree > 0] = 0                                              # :50-53
    Q = []
    for i in range(h):                                               # :55-78
        for j in range(w):
            ok1 = 0
            if seg[i, j] == 1:
                for k in range(1, 9):
                    nx, ny = i + _DIRX[k], j + _DIRY[k]
                    if nx < 0 or nx >= h or ny < 0 or ny >= w or seg[nx, ny] != 1:
                        ok1 = 1
            if ok1 == 1:                                             # ok2 can never become 1 (:66-67)
                Q.append((i, j)); vis[i, j] = 1
            if contour[i, j] > 0 and vis[i, j] == 0:
                Q.append((i, j)); vis[i, j] = 1
            k = int(dir_graph[i, j])
            if k > 0:
                nx, ny = i + _DIRX[k], j + _DIRY[k]
                if 0 <= nx < h and 0 <= ny < w:
                    hfa[nx, ny] = 1
    it = 1
    while Q:                                                         # :80-119
        NQ = []
        it += 1
        for (x, y) in Q:
            k = int(dir_graph[x, y])
            if k != 0:
                nx, ny = x + _DIRX[k], y + _DIRY[k]
                if 0 <= nx < h and 0 <= ny < w and seg[nx, ny] > 0:
                    if vis[nx, ny] == 0:
                        NQ.append((nx, ny)); vis[nx, ny] = it
                    if vis[nx, ny] == it:
                        level[nx, ny] = min(level[nx, ny], level[x, y] - 1)
                        if dir_graph[nx, ny] == 0:
                            dir_graph[nx, ny] = dir_graph[x, y]
        for (x, y) in Q:
            for k in range(1, 9):
                nx, ny = x + _DIRX[k], y + _DIRY[k]
                if 0 <= nx < h and 0 <= ny < w and seg[nx, ny] > 0 and vis[nx, ny] == 0 and hfa[nx, ny] == 0:
                    NQ.append((nx, ny)); vis[nx, ny] = it
                    if dir_graph[nx, ny] == 0:
                        dir_graph[nx, ny] = k
                        level[nx, ny] = min(level[nx, ny], level[x, y] - 1)
                    if level[x, y] <= -1:
                        level[nx, ny] = min(level[nx, ny], level[x, y])
        Q = NQ
    return level


def mudslide_watershed(seg, dir_graph, fore):
    """postprocess.py:158-181.  ``dir_graph`` is modified in place like the reference does.  -> (pred, boundary)."""
    seg = ndi.binary_fill_holes(seg)
    fore = ndi.binary_fill_holes(fore)
    fore = remove_small_objects(fore, 20)
    seg[fore == 0] = 0
    contour = (fore > 0) ^ (seg > 0)
    dir_graph_pos = remove_small_objects(dir_graph > 0, 20)
    dir_graph[dir_graph_pos == 0] = 0
    small_area = remove_small_objects(seg, 60) ^ seg
    du = graph_degree(dir_graph) > 1
    du = remove_small_objects(du, 3)
    level = mudslide_prepare(seg, dir_graph, contour, du)
    pred = level <= 0
    boundary = level > 0
    pred = remove_small_objects(pred, 15, connectivity=1)
    pred = pred ^ small_area
    return pred, boundary


# --------------------------------------------------------------------------- label generation
def _bounding_box(img):
    """hv_map.py:5-15 / distance_map.py:11-20."""
    rows, cols = np.any(img, axis=1), np.any(img, axis=0)
    rmin, rmax = np.where(rows)[0][[0, -1]]
    cmin, cmax = np.where(cols)[0][[0, -1]]
    return [rmin, rmax + 1, cmin, cmax + 1]


def _expanded_box(inst_map, h, w):
    box = _bounding_box(inst_map)
    box[0] -= 2; box[2] -= 2; box[1] += 2; box[3] += 2
    box[0] = max(box[0], 0); box[2] = max(box[2], 0); box[1] = min(box[1], h); box[3] = min(box[3], w)
    return box


def gen_instance_hv_map(inst_gt):
    """hv_map.py:18-97."""
    x_map = np.zeros(inst_gt.shape[:2], np.float32)
    y_map = np.zeros(inst_gt.shape[:2], np.float32)
    h, w = inst_gt.shape[:2]
    for inst_id in np.unique(inst_gt):
        if inst_id == 0:
            continue
        inst_map = np.array(inst_gt == inst_id, np.uint8)
        box = _expanded_box(inst_map, h, w)
        inst_map = inst_map[box[0]:box[1], box[2]:box[3]]
        if inst_map.shape[0] < 2 or inst_map.shape[1] < 2:
            continue
        com = list(ndi.center_of_mass(inst_map))
        com[0] = int(com[0] + 0.5); com[1] = int(com[1] + 0.5)
        xr = np.arange(1, inst_map.shape[1] + 1) - com[1]
        yr = np.arange(1, inst_map.shape[0] + 1) - com[0]
        ix, iy = np.meshgrid(xr, yr)
        ix[inst_map == 0] = 0; iy[inst_map == 0] = 0
        ix = ix.astype("float32"); iy = iy.astype("float32")
        if np.min(ix) < 0:
            ix[ix < 0] /= -np.amin(ix[ix < 0])
        if np.min(iy) < 0:
            iy[iy < 0] /= -np.amin(iy[iy < 0])
        if np.max(ix) > 0:
            ix[ix > 0] /= np.amax(ix[ix > 0])
        if np.max(iy) > 0:
            iy[iy > 0] /= np.amax(iy[iy > 0])
        x_map[box[0]:box[1], box[2]:box[3]][inst_map > 0] = ix[inst_map > 0]
        y_map[box[0]:box[1], box[2]:box[3]][inst_map > 0] = iy[inst_map > 0]
    return np.dstack([x_map, y_map])


def fix_inst(inst_gt):
    """distance_map.py:41-57 ``_fix_inst``: per id drop the 4-connected pieces below 5 px, split into 8-connected
    components, renumber consecutively in (id, raster) order."""
    cur = 0
    new = np.zeros_like(inst_gt)
    for inst_id in np.unique(inst_gt):
        if inst_id == 0:
            continue
        m = remove_small_objects(inst_gt == inst_id, 5)
        rem = sk_label(np.array(m, np.uint8))
        rem[rem > 0] += cur
        new[rem > 0] = rem[rem > 0]
        cur += len(np.unique(rem[rem > 0]))
    return new


def instance_distance_map(inst_gt, inst_norm=True):
    """distance_map.py:67-106 (after ``_fix_inst``): per-instance scipy chessboard distance on the expanded crop."""
    dist = np.zeros(inst_gt.shape, np.float32)
    h, w = inst_gt.shape[:2]
    for inst_id in np.unique(inst_gt):
        if inst_id == 0:
            continue
        inst_map = (inst_gt == inst_id).astype(np.uint8)
        box = _expanded_box(inst_map, h, w)
        inst_map = inst_map[box[0]:box[1], box[2]:box[3]]
        if inst_map.shape[0] < 2 or inst_map.shape[1] < 2:
            continue
        d = ndi.distance_transform_cdt(inst_map).astype("float32")
        if inst_norm:
            mx = np.amax(d)
            if mx <= 0:
                continue
            d = d / np.amax(d)
        dist[box[0]:box[1], box[2]:box[3]][inst_map > 0] = d[inst_map > 0]
    return dist


def diamond(radius):
    """skimage.morphology.diamond: {(x, y): |x| + |y| <= r} as uint8."""
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    return (np.abs(X) + np.abs(Y) <= radius).astype(np.uint8)


def bound_label(sem_gt, inst_gt, edge_id=2, radius=(3, 3)):
    """bound_map.py:62-88 (after ``_fix_inst``) -> (sem_gt, sem_gt_w_bound)."""
    sem_gt = sem_gt.copy()
    sem_gt[inst_gt == 0] = 0
    out = np.zeros_like(sem_gt)
    out += sem_gt
    for inst_id in np.unique(inst_gt):
        if inst_id == 0:
            continue
        m = inst_gt == inst_id
        bound = dilation(m, diamond(radius[0])) & (~erosion(m, diamond(radius[1])))
        out[bound > 0] = edge_id
    return sem_gt, out


def unet_weight_map(inst_gt, w0=10.0, sigma=5.0):
    """unet_map.py:53-98 (after ``_fix_inst``; wc = None) -> (eroded instance map, weight map + 1)."""
    inner = np.zeros(inst_gt.shape[:2], np.int32)
    for inst_id in np.unique(inst_gt):
        if inst_id == 0:
            continue
        m = erosion((inst_gt == inst_id).astype(np.uint8), diamond(1))
        inner[m > 0] = inst_id
    ids = [i for i in np.unique(inner) if i > 0]
    if len(ids) <= 1:
        return inner, np.zeros(inner.shape[:2]) + 1
    stacked = np.zeros(inner.shape[:2] + (len(ids), ))
    for k, inst_id in enumerate(ids):
        stacked[..., k] = ndi.distance_transform_edt(np.array(inner != inst_id, np.uint8))
    near1 = np.amin(stacked, axis=2)
    near2 = stacked - np.expand_dims(near1, axis=2)
    near2[near2 == 0] = np.inf
    near2 = np.amin(near2, axis=2)
    near2[inner > 0] = 0
    near2 = near2 + near1
    eve = (1.0 + stacked) / (1.0 + np.expand_dims(near1, axis=2))
    eve[eve != 1] = 0
    eve = np.sum(eve, axis=2)
    near2[eve > 1] = near1[eve > 1]
    pen = (near1 + near2) / sigma
    pen = w0 * np.exp(-pen**2 / 2)
    pen[inner > 0] = 0
    return inner, pen + 1


# --------------------------------------------------------------------------- DirectionLabelMake (direction_map.py)
def _centerness_point(mask, H, W):
    """calculate_centerpoint (center_calculation.py:8-54): FCOS centerness by binary search along eight directions
    (float64 arithmetic, round-half-even like Python's round); first maximum in raster order."""
    import math
    dirs = [(math.sin(2 * math.pi / 8 * i), math.cos(2 * math.pi / 8 * i)) for i in range(8)]
    best, bx, by = -1, -1, -1
    ys, xs = np.nonzero(mask)
    for i, j in zip(ys.tolist(), xs.tolist()):
        max_d, min_d = 0, 10000000
        for k in range(8):
            lo, hi = 0, 1000000
            while abs(lo - hi) > 0.1:
                mid = (lo + hi) / 2
                xo, yo = round(i + dirs[k][0] * mid), round(j + dirs[k][1] * mid)
                if xo >= 0 and yo < W and yo >= 0 and xo < H and mask[xo][yo] > 0:
                    lo = mid
                else:
                    hi = mid
            max_d, min_d = max(max_d, hi), min(min_d, lo)
        c = min_d / max_d
        if c > best:
            best, bx, by = c, i, j
    return bx, by


def sobel_kernel_11():
    """Sobel.kernel (gradient_calculation.py:13-38): [2, 11, 11] float32, channel 0 = d/d(row), channel 1 = d/d(column)"""
    k = np.zeros((2, 11, 11), np.float32)
    for j in range(11):
        for i in range(11):
            j_, i_ = j - 5, i - 5
            if j_ == 0 and i_ == 0:
                continue
            k[0, j, i] = j_ / float(i_ * i_ + j_ * j_)
            k[1, j, i] = i_ / float(i_ * i_ + j_ * j_)
    return k


def direction_bins(angle_deg, num_angles):
    """align_angle (direction_calculation.py:60-73) index: bin 0 wraps around +-180, bin i is centred on -180 + step * i"""
    step = 360 / num_angles
    a = np.asarray(angle_deg)
    idx = np.zeros(a.shape, np.int64)
    for i in range(1, num_angles):
        middle = -180 + step * i
        idx[(a > (middle - step / 2)) & (a <= (middle + step / 2))] = i
    return idx


def direction_label_make(inst_gt, sem_gt, num_angles=8, to_center=True):
    """DirectionLabelMake.__call__ (direction_map.py:36-84) -> dict(sem_gt, inst_gt, dist_gt, point_gt, dir_gt, reg_dir_gt,
    loss_weight_map, centers).  The gradient is the 11x11 correlation in float32 (the reference: torch F.conv2d, whose
    accumulation order is the backend's: float tolerance); dir_gt quantises its angle, so it can differ from the
    reference where the angle sits on a bin edge."""
    from scipy.ndimage import gaussian_filter, distance_transform_edt, correlate, grey_dilation
    inst = fix_inst(inst_gt)
    sem = np.array(sem_gt, copy=True)
    sem[inst == 0] = 0
    H, W = inst.shape
    dmap = np.zeros((H, W), np.float32)
    grad = np.zeros((H, W, 2), np.float32)
    point = np.zeros((H, W), np.float32)
    ker = sobel_kernel_11()
    centers = []
    yy, xx = np.mgrid[0:H, 0:W]
    for k in np.unique(inst):
        if k == 0:
            continue
        m = (inst == k).astype(np.uint8)
        cy, cx = _centerness_point(m, H, W)
        centers.append((int(k), cy, cx))
        point[cy, cx] = 1
        if to_center:
            d = np.sqrt(((yy - cy) ** 2 + (xx - cx) ** 2).astype(np.float64)) * m
            di = (1 - d / (d.max() + 0.0000001)) * m
        else:
            d = distance_transform_edt(m) * m
            di = (d / (d.max() + 0.0000001)) * m
        dmap += di
        d32 = di.astype(np.float32)
        g = np.stack([correlate(d32, ker[c], mode="constant", cval=0.0) for c in range(2)], -1).astype(np.float32)
        g[m == 0] = 0
        grad[m != 0] = 0
        grad += g
    point_g = gaussian_filter(point * 255, sigma=2, order=0).astype(np.float32)
    dist = (dmap ** 0.5) * 10
    angle = np.degrees(np.arctan2(grad[:, :, 0], grad[:, :, 1]))
    a0 = angle.copy()
    a0[inst == 0] = 0
    dir_map = direction_bins(a0, num_angles)
    dir_map[inst == 0] = -1
    dir_map = dir_map + 1
    reg = angle.copy()
    reg[reg < 0] += 360
    reg[inst == 0] = 0
    reg = reg / 180 * np.pi
    if num_angles == 8:
        dd = direction_differential_map(dir_map, num_angles + 1)
        w = dd * (10 - dist)
        w = grey_dilation(w, footprint=disk(1))
        w = w.astype(np.float32) * 2 + 1.0
    else:
        w = np.zeros_like(dir_map)
    return dict(sem_gt=sem, inst_gt=inst, dist_gt=dist, point_gt=point_g, dir_gt=dir_map, reg_dir_gt=reg.astype(np.float32),
                loss_weight_map=w, centers=centers, angle=angle, grad=grad)
