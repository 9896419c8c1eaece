#!/bin/bash
# second one-GPU pass on layout CF: TMA bulk copy of the column tables, 640/768-thread
# variants, row bands; the CF GPU tests; one full ncu capture of the CF kernels
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
( time timeout 170 python scripts/dev_column.py ) > $OUT/column2.txt 2>&1
echo "dev_column exit: $?" >> $OUT/column2.txt
( time timeout 90 python -m pytest tests/test_parity.py -m gpu -q -k "column" ) > $OUT/pytest_column.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_column.log
COLUMN=on timeout 200 ncu --set full --clock-control none --import-source on \
    -k regex:'k_sweep_fact_column|k_column_table' -s 4 -c 2 -o $OUT/prof_fact_column_large -f \
    python scripts/ncu_target.py large on 4 > $OUT/ncu_column.log 2>&1
ncu -i $OUT/prof_fact_column_large.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_summary.py \
    > $OUT/ncu_fact_column_config5.txt 2>&1
ls -la $OUT
