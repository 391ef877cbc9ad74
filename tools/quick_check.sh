#!/bin/bash
# Quick GPU check (run under gpurun): parity suite, then the short device-resident bench; prints ms/step and per-kernel ms.
#   bash tools/quick_check.sh <tag> [extra bench args]
tag=${1:-q}; shift
o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q -x > $o/${tag}_pytest.log 2>&1; tail -2 $o/${tag}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-config4 --no-csd5 "$@" > $o/${tag}_bench.log 2>&1
python - <<PY
import json
l=[x for x in open("$o/${tag}_bench.log") if x.startswith("{")]
if not l:
    print(open("$o/${tag}_bench.log").read()[-2000:])
else:
    d=json.loads(l[-1]); print("ms_per_step", d["ms_per_step"], "dense", d.get("ms_per_step_dense_layout"), {k:round(v["ms_per_launch"],4) for k,v in d["kernels"].items()})
PY
