#!/bin/bash
# round-end sanity on the GPU box: parity suite, the bench line, smoke()
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final.json").read())
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"])
PY
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
