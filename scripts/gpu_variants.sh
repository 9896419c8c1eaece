#!/bin/bash
# kernel tuning variants built side by side (stodynprog_b200/_lib/variants), config #5
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
V=stodynprog_b200/_lib/variants
( echo "== default"; COMPRESS=on,off timeout 300 python scripts/dev_factored.py large 2>&1 | grep -v "^large distinct"
  for v in bf_minb4 bf_ub2; do echo "== $v"; SDP_B200_LIB=$PWD/$V/libsdp_$v.so COMPRESS=on timeout 300 python scripts/dev_factored.py large 2>&1 | grep "{"; done
  for v in tma128_5 tma128_6; do echo "== $v"; SDP_B200_LIB=$PWD/$V/libsdp_$v.so COMPRESS=off timeout 300 python scripts/dev_factored.py large 2>&1 | grep "{"; done
) > $OUT/variants.txt 2>&1
cat $OUT/variants.txt
