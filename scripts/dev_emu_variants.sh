# layout CF on config #5: the whole grid on one GPU (three row bands walked column by column) and the
# 8 column shards of an 8-GPU run, streaming pass only (pre-pass + column sweep)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( for cfg in ${CFGS:-"SDP_PDL=1" "SDP_PDL=0"}; do
   echo "== $cfg"
   env $cfg BANDS=auto AXES=columns timeout 300 python scripts/dev_shard_emulation.py ${SHARDS:-8} 2>&1 | grep -v "Warning\|OPTS"
done ) > gpurun_out/r2_emu_final.txt 2>&1
cat gpurun_out/r2_emu_final.txt
