#!/bin/bash
# two-GPU pass: sharded parity through the public API (peer-memory exchange and NCCL),
# bench at N=2 (default = fused combine + all-gather over peer memory; SDP_P2P=0 = NCCL)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 400 $TR --master-port 29511 tests/multi_gpu_check.py ) > $OUT/multi_gpu_check_n$N.log 2>&1
echo "exit: $?" >> $OUT/multi_gpu_check_n$N.log
if [ -z "${SKIP_NCCL:-}" ]; then
( time SDP_P2P=0 timeout 400 $TR --master-port 29512 tests/multi_gpu_check.py ) > $OUT/multi_gpu_check_nccl_n$N.log 2>&1
echo "exit: $?" >> $OUT/multi_gpu_check_nccl_n$N.log
fi
( time timeout 500 $TR --master-port 29513 bench.py --gpus $N ) > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
if [ -z "${SKIP_NCCL:-}" ]; then
( time SDP_P2P=0 timeout 500 $TR --master-port 29514 bench.py --gpus $N --no-dense ) > $OUT/bench_nccl_n$N.json 2> $OUT/bench_nccl_n$N.err
fi
ls -la $OUT
