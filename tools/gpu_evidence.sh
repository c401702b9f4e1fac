#!/bin/bash
# One B200 box: full GPU test suite, smoke, bench lines (configs 2/3/4), ncu launch list + full captures.
# Usage (from the repo root, through gpurun): bash tools/gpu_evidence.sh <tag>
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/gpu_$TAG.txt; nproc >> $O/gpu_$TAG.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> $O/smoke_$TAG.log
timeout 400 python bench.py --steps 300 --warmup 10 > $O/bench_$TAG.json 2> $O/bench_$TAG.err
timeout 400 python bench.py --steps 200 --warmup 10 --cpu-sample 6 --config synthetic_1920x1440_500feat_10kf > $O/bench_cfg3_$TAG.json 2> $O/bench_cfg3_$TAG.err
timeout 600 python bench.py --steps 60 --warmup 5 --cpu-sample 2 --config synthetic_3840x2880_2000feat_20kf > $O/bench_cfg4_$TAG.json 2> $O/bench_cfg4_$TAG.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_$TAG.csv python bench.py --steps 8 --warmup 3 --cpu-sample 2 --pool 16 > $O/bench_under_ncu_$TAG.log 2>&1
# full captures: a handful of launches per kernel, exported to CSV on the box (gpurun_out/ is capped at 64 MiB, the .ncu-rep is not kept)
timeout 500 ncu --set full --clock-control none -k regex:"chol_kernel|accumulate_kernel|evaluate_kernel|marg_finish_kernel|marg_accumulate_kernel|lk_kernel|min_eig_kernel|candidates_kernel|clahe_apply_kernel|pyr_down_kernel|set_mask_select_kernel|select_kernel|ransac_refit_kernel" -s 120 -c 40 -f -o /tmp/prof_full_$TAG python bench.py --steps 8 --warmup 3 --cpu-sample 2 --pool 16 > $O/bench_under_ncu2_$TAG.log 2>&1
ncu -i /tmp/prof_full_$TAG.ncu-rep --page raw --csv > $O/prof_full_raw_$TAG.csv 2>/dev/null
ls -la /tmp/prof_full_$TAG.ncu-rep
tail -3 $O/pytest_gpu_$TAG.log; tail -2 $O/smoke_$TAG.log; head -c 600 $O/bench_$TAG.json; echo; tail -n 2 $O/bench_cfg3_$TAG.err; tail -n 2 $O/bench_cfg4_$TAG.err; ls -la $O | tail -20
