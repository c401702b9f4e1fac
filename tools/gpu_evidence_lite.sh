#!/bin/bash
# One B200 box, short form of tools/gpu_evidence.sh: full GPU test suite, smoke, the config-2 bench line and the ncu launch
# list of a short bench run (which also exercises the initialisation kernels through bench.py's `initialisation` leg).
# Usage (from the repo root, through gpurun): bash tools/gpu_evidence_lite.sh <tag>
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/gpu_$TAG.txt; nproc >> $O/gpu_$TAG.txt
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> $O/smoke_$TAG.log
timeout 400 python bench.py --steps 300 --warmup 10 > $O/bench_$TAG.json 2> $O/bench_$TAG.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file $O/launches_$TAG.csv python bench.py --steps 8 --warmup 3 --cpu-sample 2 --pool 16 > $O/bench_under_ncu_$TAG.log 2>&1
tail -3 $O/pytest_gpu_$TAG.log; tail -2 $O/smoke_$TAG.log; head -c 400 $O/bench_$TAG.json; echo; wc -l $O/launches_$TAG.csv
