#!/bin/bash
# One measurement pass on the GPU box (run under gpurun): GPU tests, both bench arms, launch list, one --set full capture.
#   bash tools/measure_round.sh <tag>
# Outputs land in gpurun_out/<tag>_*; tools/summarize_profiles.py turns them into profiles/<tag>_*.
tag=${1:-rXX}
o=gpurun_out
mkdir -p $o
SHORT="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-config4 --no-csd5"
python -m pytest tests -m gpu -q > $o/${tag}_pytest.log 2>&1; tail -1 $o/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/${tag}_smoke.log 2>&1; tail -1 $o/${tag}_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_ref.log 2>&1; tail -c 400 $o/${tag}_bench_ref.log
python bench.py --steps 20 --warmup 3 > $o/${tag}_bench.log 2>&1; tail -c 600 $o/${tag}_bench.log
$SHORT > $o/${tag}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv $SHORT > $o/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:stft_kernel|gram_tma|gram_eig|gram_simt|eig_jacobi|eig_sort|svd_rank1' -s 28 -c 7 \
    -o $o/${tag}_prof $SHORT > $o/${tag}_ncu_full.log 2>&1
ls -la $o/${tag}_*
