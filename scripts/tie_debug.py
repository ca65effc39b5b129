import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import tiseg_b200
from tiseg_b200 import ops
from oracle import postprocess as opp
from test_gpu_ops import _ulp_spaced_logits
T, C = 1, 3
rng = np.random.default_rng(4100 + 10 * T + C)
for magnitude in (0.05, 3.0, 0.4):
    lg = _ulp_spaced_logits(rng, 2, T, C, 48, 64, magnitude)
    cls = ops.softmax_argmax(lg)
    for n in range(2):
        p = opp.softmax_tta_mean(list(lg[n]))
        want = opp.argmax_classes(p).astype(np.uint8)
        bad = np.argwhere(cls[n] != want)
        print("mag", magnitude, "n", n, "mismatches", len(bad))
        for (y, x) in bad[:6]:
            v = lg[n, 0, :, y, x]
            e = np.exp(v - v.max(), dtype=np.float32)
            ec = torch.exp(torch.from_numpy(v - v.max()).cuda()).cpu().numpy()
            print("  logits", [float.hex(float(t)) for t in v], "gap(ulp-ish)", (v.max() - v), "np e", [float.hex(float(t)) for t in e],
                  "cuda e", [float.hex(float(t)) for t in ec], "np p", [float.hex(float(t)) for t in p[:, y, x]], "gpu", cls[n, y, x], "np", want[y, x])
