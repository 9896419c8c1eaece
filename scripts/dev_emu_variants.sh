cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( for cfg in "OPTS= CHUNK=" "OPTS=col_dynamic=1 CHUNK=" "OPTS=col_dynamic=1 CHUNK=16" "OPTS= CHUNK=16" "OPTS=col_dynamic=1 CHUNK=64" "OPTS=col_threads=768 CHUNK=" "OPTS=col_threads=768,col_dynamic=1 CHUNK="; do
   env $cfg AXES=columns timeout 300 python scripts/dev_shard_emulation.py 8 2>&1 | grep -v Warning
done ) > gpurun_out/r2_emu_variants.txt 2>&1
cat gpurun_out/r2_emu_variants.txt
