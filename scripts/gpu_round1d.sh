#!/bin/bash
# last one-GPU pass of round 1 (budget: < 10 GPU-minutes): layout CF on config #5 (parity
# against BF on the full grid + timings), then the GPU test suite and a short bench with
# layout CF as the default wherever it applies
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
( time timeout 200 python scripts/dev_column.py ) > $OUT/column.txt 2>&1
echo "dev_column exit: $?" >> $OUT/column.txt
export SDP_COLUMN_HOIST=1 SDP_COLUMN_BANDS=auto
( time timeout 330 python -m pytest tests -m gpu -q --durations=12 ) > $OUT/pytest_gpu_cf.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu_cf.log
( time timeout 150 python bench.py --no-dense --no-extra --no-cpu-baseline ) > $OUT/bench_cf.json 2> $OUT/bench_cf.err
echo "bench exit: $?" >> $OUT/bench_cf.err
ls -la $OUT
