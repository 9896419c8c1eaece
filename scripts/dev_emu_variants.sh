# kernel variants of layout CF on config #5: the whole grid on one GPU with three row bands (the
# N=1 default) and with one band, and the 8 column shards of an 8-GPU run
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( for cfg in "OPTS=col_threads=640,col_dynamic=0" "OPTS=col_threads=640,col_dynamic=1" "OPTS=col_threads=768,col_dynamic=0" "OPTS=col_threads=768,col_dynamic=1" "OPTS=col_threads=640,col_dynamic=0,col_pf=1" "OPTS=col_threads=512,col_dynamic=1"; do
   env $cfg BANDS=auto timeout 300 python scripts/dev_shard_emulation.py 1 2>&1 | grep -v Warning
   env $cfg BANDS=1 AXES=columns timeout 300 python scripts/dev_shard_emulation.py ${SHARDS:-1} 2>&1 | grep -v "Warning\|OPTS"
done ) > gpurun_out/r2_emu_variants_bands.txt 2>&1
cat gpurun_out/r2_emu_variants_bands.txt
