#!/bin/bash
# one-GPU pass #2: parity tests, bench, e2e breakdown, recursion / policy-iteration timings,
# dense vs factored timing of configs #3/#4 (hoist kernel change)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
( time timeout 600 python bench.py ) > $OUT/bench_n1.json 2> $OUT/bench_n1.err
timeout 300 python scripts/dev_e2e_breakdown.py > $OUT/e2e_breakdown.txt 2>&1
timeout 600 python scripts/dev_recursion_timing.py --searev > $OUT/recursion_timing.txt 2>&1
COMPRESS=on timeout 600 python scripts/dev_factored.py ar1 searev > $OUT/dev_factored.txt 2>&1
ls -la $OUT
