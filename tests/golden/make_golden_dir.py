#!/usr/bin/env python
"""Golden vectors for DirectionLabelMake (train-time label generation), produced by the reference's OWN
files, imported by path:

    tiseg/datasets/ops/direction_map.py        DirectionLabelMake (the class under test, unmodified)
    tiseg/datasets/utils/center_calculation.py calculate_centerpoint (numba, the real numba here)
    tiseg/datasets/utils/gradient_calculation.py  calculate_gradient (torch F.conv2d, 11x11 kernel)
    tiseg/datasets/utils/direction_calculation.py angle_to_vector / vector_to_label
    tiseg/models/utils/direct_diff_map.py      generate_direction_differential_map

Stand-ins (this image has neither mmcv nor scikit-image; numpy is 2.x):
  * np.float / np.int / np.bool (removed in numpy 1.24) are restored as the builtins for the reference's own lines
    (direction_calculation.py:64-65, 78, 100);
  * skimage.measure.label / morphology.remove_small_objects -> the oracle's C port (as in make_golden.py);
    morphology.dilation(image, selem) -> scipy.ndimage.grey_dilation(image, footprint=selem) — what skimage 0.18.3's
    dilation calls for an odd-sized footprint; morphology.selem.disk -> the oracle's disk.

Writes tests/golden/dirlabel_ref.npz.  Run in the build container (needs /root/reference):  python tests/golden/make_golden_dir.py
"""
import importlib.util, os, sys, types
import numpy as np
from scipy import ndimage as ndi

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load(modname, path, package=None):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    for alias, t in (("float", float), ("int", int), ("bool", bool)):
        if not hasattr(np, alias):
            setattr(np, alias, t)
    from oracle import postprocess as opp
    from oracle import skimage_port as skp
    sk = _stub("skimage")
    sk.measure = _stub("skimage.measure", label=skp.label)
    sk.morphology = _stub("skimage.morphology", remove_small_objects=opp.remove_small_objects,
                          dilation=lambda image, selem=None: ndi.grey_dilation(image, footprint=selem[::-1, ::-1]))
    sk.morphology.selem = _stub("skimage.morphology.selem", disk=opp.disk)
    # package skeleton so that the relative imports of direction_map.py resolve to the reference's own files
    for name in ("refpkg", "refpkg.datasets", "refpkg.datasets.ops", "refpkg.models"):
        _stub(name).__path__ = []
    cc = load("refpkg.datasets.utils.center_calculation", os.path.join(REF, "tiseg/datasets/utils/center_calculation.py"))
    gc = load("refpkg.datasets.utils.gradient_calculation", os.path.join(REF, "tiseg/datasets/utils/gradient_calculation.py"))
    utils = _stub("refpkg.datasets.utils", calculate_centerpoint=cc.calculate_centerpoint, calculate_gradient=gc.calculate_gradient)
    utils.__path__ = []
    dc = load("refpkg.datasets.utils.direction_calculation", os.path.join(REF, "tiseg/datasets/utils/direction_calculation.py"),
              "refpkg.datasets.utils")
    utils.angle_to_vector, utils.vector_to_label = dc.angle_to_vector, dc.vector_to_label
    ddm = load("refpkg.models.utils.direct_diff_map", os.path.join(REF, "tiseg/models/utils/direct_diff_map.py"))
    _stub("refpkg.models.utils", generate_direction_differential_map=ddm.generate_direction_differential_map)
    dm = load("refpkg.datasets.ops.direction_map", os.path.join(REF, "tiseg/datasets/ops/direction_map.py"), "refpkg.datasets.ops")
    return dm


CASES = [(9950, 64, 80, 10, 8, True), (9951, 96, 96, 18, 8, True), (9952, 50, 61, 6, 4, True), (9953, 72, 64, 9, 16, True),
         (9954, 40, 40, 1, 8, True), (9955, 64, 64, 8, 8, False)]


def case_inputs(seed, H, W, n):
    import tiseg_b200  # noqa: F401
    from tiseg_b200 import synth
    t = synth.gt_and_pred(seed, H, W, n=n, num_classes=4)
    inst = t["gt_inst"].astype(np.int32)
    if n == 1:
        inst[:] = 0
        inst[6:30, 8:33] = 5
        inst[2:5, 2:4] = 9            # below the small-object limit of _fix_inst together with the next one
        inst[33:36, 30:38] = 9        # one id, two components
    sem = t["gt_sem"].astype(np.uint8)
    sem[(inst > 0) & (sem == 0)] = 1
    return inst, sem


def main():
    dm = load_reference()
    out = {"n_cases": np.int64(len(CASES))}
    for j, (seed, H, W, n, A, to_center) in enumerate(CASES):
        inst, sem = case_inputs(seed, H, W, n)
        data = dict(sem_gt=sem.copy(), inst_gt=inst.copy(), seg_fields=[])
        res = dm.DirectionLabelMake(to_center=to_center, num_angles=A)(data)
        p = "d%d_" % j
        out[p + "inst"], out[p + "sem"] = inst, sem
        out[p + "num_angles"], out[p + "to_center"] = np.int64(A), np.int64(to_center)
        out[p + "fixed"] = dm.DirectionLabelMake()._fix_inst(inst)
        for key in ("sem_gt", "dist_gt", "point_gt", "dir_gt", "reg_dir_gt", "loss_weight_map"):
            out[p + key] = np.asarray(res[key])
        print(j, (H, W), "A", A, "instances", int(out[p + "fixed"].max()), {k: (out[p + k].dtype, out[p + k].shape) for k in ("dist_gt", "point_gt", "dir_gt", "reg_dir_gt", "loss_weight_map")})
    np.savez_compressed(os.path.join(HERE, "dirlabel_ref.npz"), **out)


if __name__ == "__main__":
    main()
