"""Reference-side pipelines assembled from the CPU oracle (TEST INFRASTRUCTURE ONLY): what
``Dataset.pre_eval`` of the reference would return for one image (custom.py:252-283, conic.py:157-196)."""
import numpy as np
import torch

from oracle import metrics as om


def oracle_pre_eval(pred, sem_gt, inst_gt, num_classes, multi=False, name=None, literal=False):
    sem_pred, inst_pred = np.asarray(pred['sem_pred']), np.asarray(pred['inst_pred'])
    sem_res = tuple(torch.from_numpy(x) for x in om.pre_eval_all_semantic_metric(sem_pred, sem_gt, num_classes))
    ip, ig = om.re_instance(inst_pred), om.re_instance(om.re_instance(inst_gt))
    d = {}
    if name is not None:
        d['name'] = name
    aji = om.pre_eval_bin_aji(ip, ig, literal=literal)
    pq = om.pre_eval_bin_pq(ip, ig, literal=literal)
    d['bin_aji_pre_eval_res'] = (np.float64(aji[0]), np.float64(aji[1]))
    d['bin_pq_pre_eval_res'] = (int(pq[0]), int(pq[1]), int(pq[2]), np.float64(pq[3]))
    d['sem_pre_eval_res'] = sem_res
    if multi:
        dp = om.assign_sem_class_to_insts(ip, sem_pred, num_classes)
        dg = om.assign_sem_class_to_insts(ig, sem_gt, num_classes)
        d['aji_pre_eval_res'] = om.pre_eval_aji(ip, ig, dp, dg, num_classes, literal=literal)
        d['pq_pre_eval_res'] = om.pre_eval_pq(ip, ig, dp, dg, num_classes, literal=literal)
    return d


def same_result(a, b):
    assert set(a) == set(b), (sorted(a), sorted(b))
    for k in a:
        va, vb = a[k], b[k]
        if k == 'name':
            assert va == vb
            continue
        assert len(va) == len(vb), k
        for x, y in zip(va, vb):
            x = x.numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
            y = y.numpy() if isinstance(y, torch.Tensor) else np.asarray(y)
            assert x.dtype == y.dtype or x.ndim == 0, (k, x.dtype, y.dtype)
            assert np.array_equal(x, y), (k, x, y)
