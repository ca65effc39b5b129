"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports
exactly the symbols include/tiseg_b200.h declares; without a GPU it fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from tiseg_b200 import _lib
    return _lib


def _declared():
    src = open(os.path.join(ROOT, "include", "tiseg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tiseg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound(lib):
    names = _declared()
    assert len(names) >= 20
    so = ctypes.CDLL(lib.SO_PATH)
    for n in names:
        assert hasattr(so, n), "declared in tiseg_b200.h but not exported: " + n
    assert sorted(lib.SIGNATURES) == names, "python binding table and header disagree"
    assert lib.load().tiseg_version() >= 100


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from tiseg_b200 import ops
    with pytest.raises(lib.TisegError, match="no CUDA device|NOGPU|status 3"):
        ops.label(np.zeros((4, 4), np.int32))
    # the product never imports the oracle
    pkg = os.path.join(ROOT, "tissue-image-segmentation_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, f)).read().replace("CPU oracle", ""), f


def test_label_maker_classes_mirror_the_reference_constructors():
    """Host-side mirror of tiseg/datasets/ops: same class names and constructor arguments; the one configuration the
    reference itself cannot run (UNetLabelMake with class weights, unet_map.py:120-124) is refused, not emulated."""
    import inspect
    from tiseg_b200 import label_makers as lm
    assert list(inspect.signature(lm.BoundLabelMake.__init__).parameters)[1:] == ["edge_id", "selem_radius"]
    assert list(inspect.signature(lm.DistanceLabelMake.__init__).parameters)[1:] == ["inst_norm"]
    assert list(inspect.signature(lm.UNetLabelMake.__init__).parameters)[1:] == ["wc", "w0", "sigma"]
    assert lm.BoundLabelMake(selem_radius=2).radius == (2, 2)
    with pytest.raises(NotImplementedError):
        lm.UNetLabelMake(wc={1: 2.0})
