#!/usr/bin/env python
"""Golden vectors for MultiTaskCDNet.inference with use_regression = True (the two `_jour_regression` configs): the method
is executed FROM ITS SOURCE TEXT (tiseg/models/segmentors/multi_task_cdnet.py:246-330, with _ddm_enhencement :548-564,
BaseSegmentor.tta_transform / reverse_tta_transform, generate_direction_differential_map, and angle_to_vector /
vector_to_label of tiseg/datasets/utils/direction_calculation.py loaded by path) with a recorder in place of the network.

Stand-ins: np.float / np.int / np.bool restored as the builtins (removed in numpy 1.24; direction_calculation.py uses
them); `Tensor.cuda()` is the identity here (no GPU in the build container).

Writes tests/golden/reg_ref.npz.   python tests/golden/make_golden_reg.py
"""
import os, sys, types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import make_golden as mg          # noqa: E402
import make_golden_r2 as r2       # noqa: E402
import make_golden_dir as gd      # noqa: E402


def main():
    import torch
    import torch.nn.functional as F
    for alias, t in (("float", float), ("int", int), ("bool", bool)):
        if not hasattr(np, alias):
            setattr(np, alias, t)
    torch.Tensor.cuda = lambda self, *a, **k: self
    misc, inst_mod, sem_mod, isem, pp, ddm = mg.load_reference()
    for name in ("refpkg", "refpkg.datasets"):
        gd._stub(name).__path__ = []
    u = gd._stub("refpkg.datasets.utils", calculate_centerpoint=None, calculate_gradient=None)
    u.__path__ = []
    gd._stub("refpkg.datasets.utils.center_calculation", calculate_centerpoint=None)
    gd._stub("refpkg.datasets.utils.gradient_calculation", calculate_gradient=None)
    dc = gd.load("refpkg.datasets.utils.direction_calculation", os.path.join(mg.REF, "tiseg/datasets/utils/direction_calculation.py"),
                 "refpkg.datasets.utils")
    benv = {"torch": torch, "F": F, "resize": None, "np": np}
    b_tta = r2.ref_method("tiseg/models/segmentors/base.py", "BaseSegmentor", "tta_transform", benv)
    b_rev = r2.ref_method("tiseg/models/segmentors/base.py", "BaseSegmentor", "reverse_tta_transform", benv)
    cenv = dict(benv)
    cenv.update(generate_direction_differential_map=ddm.generate_direction_differential_map, angle_to_vector=dc.angle_to_vector,
                vector_to_label=dc.vector_to_label)
    mt_inference = r2.ref_method("tiseg/models/segmentors/multi_task_cdnet.py", "MultiTaskCDNet", "inference", cenv)
    mt_enh = r2.ref_method("tiseg/models/segmentors/multi_task_cdnet.py", "MultiTaskCDNet", "_ddm_enhencement", cenv)
    out = {}
    cases = [(40, 44, ["none", "horizontal", "vertical"], [0], True, 4), (36, 36, ["none"], [0], False, 2),
             (32, 48, ["none", "diagonal"], [0, 90], True, 3)]
    for j, (H, W, flips, rots, if_ddm, Csem) in enumerate(cases):
        rng = np.random.default_rng(9850 + j)
        rec = []

        def whole(img, meta, rescale):
            outs = [torch.from_numpy((rng.standard_normal((1, c, img.shape[2], img.shape[3])) * 2).astype(np.float16)
                                     .astype(np.float32)) for c in (3, Csem, 1, 1)]
            # the regression head: angles in radians, some outside [0, 2 pi], some exactly on class edges / the wrap
            ang = (rng.random((1, 1, img.shape[2], img.shape[3])) * 7.6 - 0.6).astype(np.float32)
            edge = rng.random(ang.shape) < 0.05
            ang[edge] = (rng.integers(0, 17, ang.shape)[edge] * (np.pi / 8)).astype(np.float32)
            outs[2] = torch.from_numpy(ang)
            outs[3] = outs[3].abs()
            rec.append([o.numpy()[0].astype(np.float16 if k != 2 else np.float32) for k, o in enumerate(outs)])
            return tuple(o.clone() for o in outs)
        me = types.SimpleNamespace(num_classes=Csem, num_angles=8, whole_inference=whole, use_regression=True, if_ddm=if_ddm)
        me.test_cfg = r2.Cfg(mode="whole", rotate_degrees=rots, flip_directions=flips)
        me.tta_transform = lambda img, r, f: b_tta(me, img, r, f)
        me.reverse_tta_transform = lambda img, r, f: b_rev(me, img, r, f)
        me._ddm_enhencement = lambda a, b, c: mt_enh(me, a, b, c)
        tc_out, sem_out, dir_out = mt_inference(me, torch.zeros(1, 3, H, W), None, False)
        out["g%d_flips" % j] = np.array([f for _ in rots for f in flips])
        out["g%d_rots" % j] = np.array([r for r in rots for _ in flips])
        out["g%d_if_ddm" % j] = np.array(if_ddm)
        for t, (a, b, c, d) in enumerate(rec):
            out["g%d_tc%d" % (j, t)], out["g%d_sem%d" % (j, t)] = a, b
            out["g%d_dir%d" % (j, t)], out["g%d_pt%d" % (j, t)] = c, d
        out["g%d_tc_out" % j], out["g%d_sem_out" % j] = tc_out.numpy()[0], sem_out.numpy()[0]
        out["g%d_dir_out" % j] = dir_out.numpy()[0]
        print(j, tc_out.shape, dir_out.shape, dir_out.dtype, np.unique(dir_out.numpy()))
    out["n_cases"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, "reg_ref.npz"), **out)


if __name__ == "__main__":
    main()
