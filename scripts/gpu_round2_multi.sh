#!/bin/bash
# multi-GPU pass for the next round: parity of the sharded sweep (incl. layout CF cut by rows
# and by columns), then the bench with both cuts.   bash scripts/gpu_round2_multi.sh N
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time SDP_CHECK_COLUMN_AXIS=1 timeout 300 $TR --master-port 29511 tests/multi_gpu_check.py ) > $OUT/multi_gpu_check_n$N.log 2>&1
echo "exit: $?" >> $OUT/multi_gpu_check_n$N.log
for AXIS in rows columns; do
( time SDP_SLAB_AXIS=$AXIS timeout 200 $TR --master-port 29513 bench.py --gpus $N --no-dense --no-extra \
    --no-cpu-baseline ) > $OUT/bench_n${N}_$AXIS.json 2> $OUT/bench_n${N}_$AXIS.err
done
ls -la $OUT
