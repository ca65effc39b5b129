import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
if os.environ.get("WSDBG"): os.environ["TISEG_DEBUG_FLOOD"] = "1"
import numpy as np
from scipy import ndimage as ndi
import tiseg_b200
from tiseg_b200 import ops
from oracle import skimage_port as sk
from test_gpu_ops import _random_ws_case
rng = np.random.default_rng(50)
for (H, W, levels, nmark) in [(1, 9, 3, 2), (20, 31, 4, 5)]:
    img, mk, mask = _random_ws_case(rng, H, W, levels, nmark)
    for rep in range(2):
        got = ops.watershed(img, mk, mask); want = sk.watershed(img, mk, mask)
        bad = (got != want)
        print(H, W, "rep", rep, "mismatch", bad.sum(), "blobs", ndi.label(mask)[1], "levels", np.unique(img[mask > 0]))
        if bad.sum():
            lab, _ = ndi.label(mask)
            print(" bad blob ids", np.unique(lab[bad]), "sizes", np.bincount(lab.ravel())[np.unique(lab[bad])])
    got = ops.watershed(img, mk); want = sk.watershed(img, mk)
    print(H, W, "unmasked mismatch", (got != want).sum())
