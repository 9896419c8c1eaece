#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
( time timeout 150 python scripts/dev_column2.py ) > $OUT/column3.txt 2>&1
echo "exit: $?" >> $OUT/column3.txt
( time timeout 60 python -m pytest tests/test_parity.py -m gpu -q -k "column" ) > $OUT/pytest_column2.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_column2.log
