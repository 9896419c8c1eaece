#!/bin/bash
# one-GPU validation pass of the round: parity tests, bench (both arms), ncu launch list,
# one full ncu capture per streaming kernel, slab-size scaling of the sweep kernel
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
( time timeout 600 python bench.py ) > $OUT/bench_n1.json 2> $OUT/bench_n1.err
( time timeout 400 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
[ -n "${SKIP_SLAB:-}" ] || timeout 600 python scripts/dev_slab_scaling.py > $OUT/slab_scaling.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra --no-dense \
    > $OUT/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_fact_tiled -s 2 -c 1 \
    -o $OUT/prof_fact_tiled_large -f python scripts/ncu_target.py large on 4 > $OUT/ncu_large.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_fact_hoist -s 2 -c 1 \
    -o $OUT/prof_fact_hoist_ar1 -f python scripts/ncu_target.py ar1 on 4 > $OUT/ncu_ar1.log 2>&1
ls -la $OUT
