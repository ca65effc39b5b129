"""Test-time tails of the reference's segmentors (``tiseg/models/segmentors/*.py``): everything between the
CNN heads' raw outputs and ``[{'sem_pred': ..., 'inst_pred': ...}]``, with the reference's method names and
argument meaning.  The CNN (``calculate``) is not here — it stays in PyTorch / cuDNN; these classes take its
logits.  All compute runs in the CUDA library; inputs may be numpy arrays or CUDA tensors, single tiles
``[H, W]`` or batches ``[N, H, W]``.

    post = UNet(num_classes=2)                     # unet.py
    out = post.forward_eval(sem_logits)            # [T, C, H, W] logits of the T TTA variants
    out[0]['sem_pred'], out[0]['inst_pred']

Dtypes: ``sem_pred`` uint8, ``inst_pred`` int32 (the reference returns int32 / int64 with the same values).
"""
import numpy as np

from . import ops
from ._lib import is_torch


def _as_class_map(pred):
    """argmax output (int64 in the reference) -> uint8 working copy."""
    if is_torch(pred):
        import torch
        return pred.to(torch.uint8).contiguous()
    return np.ascontiguousarray(pred, dtype=np.uint8)


def _wrap(sem_pred, inst_pred):
    if sem_pred.ndim == 2:
        return [{'sem_pred': sem_pred, 'inst_pred': inst_pred}]
    return [{'sem_pred': s, 'inst_pred': i} for s, i in zip(sem_pred, inst_pred)]


class _UNetFamily:
    """unet.py:71-93 and siblings: per class fill holes -> remove_small_objects(5) -> label -> dilation."""
    radius = 1
    has_edge = False

    def __init__(self, num_classes, test_cfg=None):
        self.num_classes = num_classes
        self.test_cfg = dict(test_cfg or {})

    def postprocess(self, pred):
        edge = self.num_classes if self.has_edge else None
        radius = self.test_cfg.get('radius', self.radius)
        work = _as_class_map(pred)
        max_class = self.num_classes if self.has_edge else self.num_classes - 1
        sem_pred, inst_pred = ops.postproc_unet(work, max_class, radius, edge)
        if edge is not None and work is not pred:
            pred[pred == edge] = 0                      # the reference zeroes the edge class in the caller's map
        return sem_pred, inst_pred

    def forward_eval(self, sem_logit):
        """sem_logit: raw logits ``[T, C, H, W]`` (or ``[N, T, C, H, W]``) of the TTA variants."""
        sem_pred = ops.softmax_argmax(sem_logit)
        return _wrap(*self.postprocess(sem_pred))


class UNet(_UNetFamily):            # unet.py:71-93 (radius 1)
    pass


class MicroNet(_UNetFamily):        # micronet.py:185-207
    pass


class CUNet(_UNetFamily):           # cunet.py:70-93: 3-class map, edge class = num_classes
    radius, has_edge = 3, True


class FullNet(CUNet):               # fullnet.py:190-213
    pass


class CMicroNet(CUNet):             # cmicronet.py:186-209
    pass


class DCAN(_UNetFamily):            # dcan.py:193-217: the contour head splits touching cells
    radius = 3

    def postprocess(self, cell_pred, cont_pred):
        work = _as_class_map(cell_pred)
        sem_pred, inst_pred = ops.postproc_unet(work, self.num_classes - 1, self.test_cfg.get('radius', self.radius),
                                                None, kill=cont_pred)
        if work is not cell_pred:
            cell_pred[cont_pred > 0] = 0
        return sem_pred, inst_pred

    def forward_eval(self, cell_logit, cont_logit):
        return _wrap(*self.postprocess(ops.softmax_argmax(cell_logit), ops.softmax_argmax(cont_logit)))


class CDNet(CUNet):
    """cdnet.py: direction-guided refinement (inference tail :183-217, ``_ddm_enhencement`` :354-367) followed by
    the CUNet-style postprocess with radius 3 (:96-119)."""

    def __init__(self, num_classes, num_angles=8, test_cfg=None):
        super().__init__(num_classes, test_cfg)
        self.num_angles = num_angles

    def inference_tail(self, sem_logit, dir_logit, point_logit):
        """raw head outputs of the T TTA variants -> (refined sem probabilities, dir_map of variant 0)."""
        r = ops.cdnet_refine(sem_logit, dir_logit, point_logit, if_ddm=self.test_cfg.get('if_ddm', False))
        return r['sem_prob'], r['dir_map'], r['cls']

    @classmethod
    def _ddm_enhencement(cls, sem_logit, dd_map, point_logit):
        raise NotImplementedError("fused into ops.cdnet_refine (tiseg_cdnet_refine); call inference_tail")

    def forward_eval(self, sem_logit, dir_logit, point_logit):
        _, dir_map, sem_pred = self.inference_tail(sem_logit, dir_logit, point_logit)
        sem_pred, inst_pred = self.postprocess(sem_pred)
        out = _wrap(sem_pred, inst_pred)
        return out


class Dist:
    """dist.py:262-284: semantic argmax passes through, instances from the distance-map watershed."""

    def __init__(self, num_classes=2, test_cfg=None):
        self.num_classes = num_classes
        self.test_cfg = dict(test_cfg or {})

    def postprocess(self, sem_pred, dist_logit):
        return sem_pred, ops.postproc_dist(dist_logit)

    def forward_eval(self, sem_logit, dist_logit):
        """sem_logit ``[T, C, H, W]`` raw; dist_logit ``[H, W]`` = TTA mean of the distance head (dist.py:398-406)."""
        sem_pred = ops.softmax_argmax(sem_logit)
        return _wrap(*self.postprocess(sem_pred, dist_logit))


class HoverNet:
    """hovernet.py:267-365."""

    def __init__(self, num_classes=3, test_cfg=None):
        self.num_classes = num_classes
        self.test_cfg = dict(test_cfg or {})

    def hover_post_proc(self, fore_map, hv_map, fx=1, scale_factor=1):
        if fx != 1:
            raise NotImplementedError("hover_post_proc: only fx = 1 (ksize 21) is implemented")
        return ops.postproc_hover(fore_map, hv_map, scale_factor=scale_factor)

    def forward_eval(self, sem_logit, hv_map, fore_prob):
        """sem_logit ``[T, C, H, W]`` raw; hv_map ``[H, W, 2]``; fore_prob ``[H, W]`` (softmax channel 1)."""
        sem_pred = ops.softmax_argmax(sem_logit)
        inst_pred = self.hover_post_proc(fore_prob, hv_map, scale_factor=self.test_cfg.get('scale_factor', 1))
        return _wrap(sem_pred, inst_pred)


class MultiTaskUNet:
    """multi_task_unet.py:68-106 (inner map has no edge class)."""
    edge_id = None
    returns_canvas = True

    def __init__(self, num_classes, test_cfg=None):
        self.num_classes = num_classes
        self.test_cfg = dict(test_cfg or {})

    def postprocess(self, inner_pred, sem_pred):
        canvas, inst_pred = ops.postproc_multitask(inner_pred, sem_pred, self.num_classes - 1, self.edge_id, 20)
        return (canvas if self.returns_canvas else sem_pred), inst_pred

    def forward_eval(self, inner_logit, sem_logit):
        return _wrap(*self.postprocess(ops.softmax_argmax(inner_logit), ops.softmax_argmax(sem_logit)))


class MultiTaskCUNet(MultiTaskUNet):    # multi_task_cunet.py:69-108: three-class inner map, edge = 2
    edge_id = 2


class MultiTaskCDNet(MultiTaskCUNet):   # multi_task_cdnet.py:206-243: returns the RAW sem_pred, not the canvas
    returns_canvas = False
