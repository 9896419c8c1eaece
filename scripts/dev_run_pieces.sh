# developer GPU session: parity of the column-piece path, then the e2e timeline for several cuts
cd ${GRAFT_REPO_ROOT:-/root/repo}
timeout 300 python -m pytest tests -x -q -m gpu -k "pieces or config5_full" > gpurun_out/r2c_pytest.txt 2>&1; tail -3 gpurun_out/r2c_pytest.txt
: > gpurun_out/r2_e2e_timeline.txt
for V in "SDP_COLUMN_PIECES=0.3,0.3,0.3,0.1" "SDP_COLUMN_PIECES=0.3,0.3,0.25,0.15" "SDP_COLUMN_PIECES=0.3,0.3,0.2,0.12,0.08"; do
  echo "== $V" >> gpurun_out/r2_e2e_timeline.txt
  env $V timeout 100 python scripts/dev_e2e_timeline.py >> gpurun_out/r2_e2e_timeline.txt 2>&1
done
cat gpurun_out/r2_e2e_timeline.txt
