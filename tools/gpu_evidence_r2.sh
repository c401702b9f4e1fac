#!/bin/bash
# Round-2 evidence on one B200 box: full GPU test suite, smoke, the config-2 bench line (+ the reference arm), the ncu launch list of a
# short bench run and one `ncu --set full` capture of the solver chain's kernels.   Usage: bash tools/gpu_evidence_r2.sh <tag>
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/gpu_$TAG.txt; nproc >> $O/gpu_$TAG.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> $O/smoke_$TAG.log
timeout 600 python bench.py --steps 300 --warmup 10 > $O/bench_$TAG.json 2> $O/bench_$TAG.err
timeout 400 python bench.py --impl reference --steps 60 --warmup 3 > $O/bench_ref_$TAG.json 2>> $O/bench_$TAG.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file $O/launches_$TAG.csv python bench.py --steps 8 --warmup 3 --cpu-sample 2 --pool 16 --no-extras > $O/bench_under_ncu_$TAG.log 2>&1
timeout 500 ncu --set full --import-source on --clock-control none -k regex:"chol_kernel|accumulate_kernel|evaluate_kernel|marg_finish_kernel|lk_kernel|min_eig_kernel" -c 24 -f -o $O/ncu_full_$TAG python bench.py --steps 4 --warmup 3 --cpu-sample 2 --pool 16 --no-extras > $O/bench_under_ncu2_$TAG.log 2>&1
ncu -i $O/ncu_full_$TAG.ncu-rep --page raw --csv > $O/ncu_full_$TAG.csv 2>/dev/null
tail -3 $O/pytest_gpu_$TAG.log; tail -2 $O/smoke_$TAG.log; head -c 300 $O/bench_$TAG.json; echo; wc -l $O/launches_$TAG.csv
