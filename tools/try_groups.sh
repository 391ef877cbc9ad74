run() { timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-config4 --no-csd5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})"; }
run base
for g in 20 10; do for pin in 0 64 100; do SPECGPU_PIPELINE_GROUP=$g SPECGPU_PIPELINE_LANES=1 SPECGPU_L2_PIN_MB=$pin run "g$g-pin$pin"; done; done
SPECGPU_PIPELINE_GROUP=20 SPECGPU_L2_PIN_MB=64 run "g20-2lanes"
