"""CPU oracle for the tiseg test-time post-process + evaluation path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import it, and there only as the
checker / the reported baseline.  The product (``tiseg_b200``) never falls back to it.

What it is: a restatement of the reference's algorithm (clownrat6/Tissue-Image-Segmentation,
files cited per function) in numpy / scipy.ndimage / OpenCV exactly as the reference calls
them, plus a plain-C restatement (``skimage_port.c``) of the five scikit-image 0.18.3 functions
the reference depends on and that are not installable here (``measure.label``,
``morphology.remove_small_objects``, ``morphology.dilation/erosion``,
``morphology.reconstruction``, ``segmentation.watershed``) and of the numba kernel
``align_foreground``.

Pinning status: the reference has no tests, golden vectors or fixtures.  The
parts of the reference that import in the authoring container (``tiseg/utils/inst_metrics.py``,
``tiseg/utils/misc.py``, ``tiseg/models/utils/postprocess.py``,
``tiseg/models/utils/direct_diff_map.py``, ``tiseg/datasets/utils/instance_semantic.py``) were run
there to produce ``tests/golden/*.npz`` (script: ``tests/golden/make_golden.py``) and this oracle
is checked against those.  The scikit-image boundary itself is PARITY UNPINNED (no wheel, no
source): see ``skimage_port.c`` header for the one defined tie-break.
"""
from .skimage_port import label, watershed, reconstruction_erosion, align_foreground  # noqa: F401
