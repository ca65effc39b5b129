"""Probe: do CUDA expf (what torch's CUDA softmax and this library call), numpy's fp32 exp and the correctly rounded
exp agree on tiny negative arguments (the argmax tie band of softmax)?  Prints counts of disagreements."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch

out = {}
# d = -k * 2^-e for a dense set of tiny arguments
ks = np.arange(1, 4097, dtype=np.float64)
ds = []
for e in range(20, 34):
    ds.append(-(ks * 2.0 ** -e))
d = np.unique(np.concatenate(ds).astype(np.float32))
cr = np.exp(d.astype(np.float64)).astype(np.float32)          # correctly rounded (double rounding negligible here)
npx = np.exp(d)
cux = torch.exp(torch.from_numpy(d).cuda()).cpu().numpy()
tcx = torch.exp(torch.from_numpy(d)).numpy()
out["n"] = int(d.size)
out["numpy_vs_correct"] = int((npx != cr).sum())
out["cuda_vs_correct"] = int((cux != cr).sum())
out["torchcpu_vs_correct"] = int((tcx != cr).sum())
out["cuda_vs_numpy"] = int((cux != npx).sum())
bad = np.flatnonzero(cux != npx)[:10]
out["examples"] = [(float(d[i]), float(npx[i]), float(cux[i]), float(cr[i])) for i in bad]
# where does exp(d) == 1 stop?
out["numpy_last_one"] = float(d[npx == 1.0].min()) if (npx == 1.0).any() else None
out["cuda_last_one"] = float(d[cux == 1.0].min()) if (cux == 1.0).any() else None
out["numpy_first_below"] = float(d[npx < 1.0].max())
out["cuda_first_below"] = float(d[cux < 1.0].max())
# random logits, C = 2, 3, 7: class maps torch-CUDA softmax+argmax vs numpy fp32 softmax+argmax
rng = np.random.default_rng(0)
for C in (2, 3, 7):
    x = rng.normal(0, 2, (C, 4_000_000)).astype(np.float32)
    m = x.max(0, keepdims=True); e = np.exp(x - m); p = e / e.sum(0, keepdims=True)
    a_np = p.argmax(0)
    xt = torch.from_numpy(x).cuda()
    a_cu = torch.softmax(xt[None], 1)[0].argmax(0).cpu().numpy()
    a_lg = x.argmax(0)
    out["C%d_cuda_vs_numpy" % C] = int((a_np != a_cu).sum())
    out["C%d_logitargmax_vs_numpy" % C] = int((a_np != a_lg).sum())
print(json.dumps(out, indent=1))
json.dump(out, open("gpurun_out/exp_band_probe.json", "w"), indent=1)
