#!/bin/bash
# final one-GPU pass of round 1 with layout CF as the default: GPU test suite, default
# bench line, smoke(), ncu launch list of a short bench
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
( time timeout 80 python -m pytest tests -m gpu -q --durations=8 ) > $OUT/pytest_gpu_final.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu_final.log
( time timeout 120 python bench.py ) > $OUT/bench_n1.json 2> $OUT/bench_n1.err
echo "bench exit: $?" >> $OUT/bench_n1.err
( time timeout 60 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.log 2>&1
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file $OUT/launches_cf.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra --no-dense \
    > $OUT/launches_bench.log 2>&1
ls -la $OUT
