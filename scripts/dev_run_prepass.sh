# developer GPU session: col_prepass 2 (wait for the table, then take items) against 3 (first item's
# loads before the wait): parity of the column kernels, then one-GPU and 1/8-shard kernel times
cd ${GRAFT_REPO_ROOT:-/root/repo}
timeout 300 python -m pytest tests -x -q -m gpu -k "column" > gpurun_out/r2d_pytest.txt 2>&1; tail -3 gpurun_out/r2d_pytest.txt
: > gpurun_out/r2_emu_prepass.txt
for V in "col_prepass=2" "col_prepass=3" "col_prepass=2" "col_prepass=3"; do
  echo "== $V" >> gpurun_out/r2_emu_prepass.txt
  OPTS=$V AXES=columns timeout 200 python scripts/dev_shard_emulation.py 8 2>&1 | grep "^N=" >> gpurun_out/r2_emu_prepass.txt
done
cat gpurun_out/r2_emu_prepass.txt
