#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
( time timeout 600 python bench.py --no-dense --no-extra --no-cpu-baseline ) > $OUT/bench_n1_short.json 2> $OUT/bench_n1_short.err
( time SDP_OVERLAP=0 timeout 600 python bench.py --no-dense --no-extra --no-cpu-baseline ) > $OUT/bench_n1_short_nooverlap.json 2> $OUT/bench_n1_short_nooverlap.err
ls -la $OUT
