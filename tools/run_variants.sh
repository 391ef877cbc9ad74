#!/bin/bash
# Per-kernel times of the short bench for the product library and every build/variants/libspecgpu_*.so (run under gpurun).
run() {
  SPECGPU_LIB=$1 SPECGPU_BENCH_NOASSERT=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-config4 --no-csd5 2>/dev/null | python -c "
import json,sys
l=[x for x in sys.stdin if x.startswith('{')]
d=json.loads(l[-1]); print('$2', 'one-at-a-time', round(d['ms_per_step_one_shot_at_a_time'],4), 'inflight', round(d['ms_per_step'],4), {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})"
}
run "" product
for f in build/variants/libspecgpu_*.so; do run $f $(basename $f .so | sed s/libspecgpu_//); done
