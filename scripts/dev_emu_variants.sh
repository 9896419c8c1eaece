# kernel variants of layout CF on config #5: the whole grid on one GPU with three row bands (the
# N=1 default) and the 8 column shards of an 8-GPU run; one / two rows per lane, CTA shapes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( for cfg in "SDP_COLUMN_PAIRS=0 OPTS=col_threads=768,col_dynamic=1" "SDP_COLUMN_PAIRS=1 OPTS=col_threads=768,col_dynamic=1" "SDP_COLUMN_PAIRS=1 OPTS=col_threads=640,col_dynamic=1" "SDP_COLUMN_PAIRS=1 OPTS=col_threads=768,col_dynamic=0" "SDP_COLUMN_PAIRS=1 OPTS=col_threads=512,col_dynamic=1"; do
   echo "== $cfg"
   env $cfg BANDS=auto AXES=columns timeout 300 python scripts/dev_shard_emulation.py ${SHARDS:-8} 2>&1 | grep -v "Warning\|OPTS"
done ) > gpurun_out/r2_emu_variants_pairs.txt 2>&1
cat gpurun_out/r2_emu_variants_pairs.txt
