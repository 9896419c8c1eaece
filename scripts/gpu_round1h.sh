#!/bin/bash
# two-GPU sanity run of layout CF (row slabs, fused combine + exchange with the column mapping)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
export SDP_P2P_TIMEOUT_S=20
timeout 50 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-dense --no-extra --no-cpu-baseline \
    > $OUT/bench_n2_cf.json 2> $OUT/bench_n2_cf.err
echo "exit: $?" >> $OUT/bench_n2_cf.err
