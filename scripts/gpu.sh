#!/bin/bash
# One parameterised GPU session script (replaces the one-shot command lists of round 1):
#     gpurun [--gpus N] --timeout S -- 'bash scripts/gpu.sh STEP [STEP ...]'
# Steps (each writes its own files under gpurun_out/):
#   tests            pytest -m gpu + __graft_entry__.smoke()
#   bench[:ARGS]     bench.py on NGPU GPUs (NGPU env, default 1); ARGS = extra flags, '+' for spaces
#   ref              bench.py --impl reference (short)
#   check            tests/multi_gpu_check.py on NGPU GPUs, rows and columns
#   emu[:N,N]        scripts/dev_shard_emulation.py (shards of an N-GPU run timed on one GPU)
#   ncu:WHICH        ncu --set full of the streaming kernel of WHICH in {large,ar1}, + launch list
#   sanitize         compute-sanitizer memcheck / racecheck on small grids
#   timeline[:VARS]  where the time of one value_iteration call goes (per variant of environment settings)
#   emuopts:OPTS     streaming-kernel time on one GPU and on one shard of eight per set of library options
#   py:SCRIPT[:ARGS] any script under scripts/
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
N=${NGPU:-1}
TAG=${TAG:-r2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
PORT=29511
run_py() { if [ "$N" -gt 1 ]; then PORT=$((PORT+1)); timeout 900 $TR --master-port $PORT "$@"; else timeout 900 python "$@"; fi; }
for STEP in "$@"; do
  NAME=${STEP%%:*}
  ARGS=""; [ "$STEP" != "$NAME" ] && ARGS=$(echo "${STEP#*:}" | tr '+' ' ')
  case $NAME in
    tests)
      ( time timeout 900 python -m pytest tests -x -q -m gpu ) > $OUT/${TAG}_pytest_gpu.txt 2>&1
      echo "pytest exit: $?" >> $OUT/${TAG}_pytest_gpu.txt
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $OUT/${TAG}_pytest_gpu.txt 2>&1
      tail -5 $OUT/${TAG}_pytest_gpu.txt ;;
    bench)
      SUF=$(echo "$ARGS" | tr -c 'A-Za-z0-9\n' '_')
      ( time run_py bench.py --gpus $N $ARGS ) > $OUT/${TAG}_bench_n${N}${SUF}.json 2> $OUT/${TAG}_bench_n${N}${SUF}.err
      echo "exit: $?" >> $OUT/${TAG}_bench_n${N}${SUF}.err
      head -c 1500 $OUT/${TAG}_bench_n${N}${SUF}.json; echo; tail -3 $OUT/${TAG}_bench_n${N}${SUF}.err ;;
    ref)
      ( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
      cat $OUT/${TAG}_bench_reference.json ;;
    check)
      PORT=$((PORT+1))
      ( time SDP_CHECK_COLUMN_AXIS=1 SDP_CHECK_BENCH_GRID=1 timeout 600 $TR --master-port $PORT tests/multi_gpu_check.py ) > $OUT/${TAG}_multi_gpu_check_n$N.log 2>&1
      echo "exit: $?" >> $OUT/${TAG}_multi_gpu_check_n$N.log
      grep -v "^\[W\|Warning" $OUT/${TAG}_multi_gpu_check_n$N.log | tail -30 ;;
    emu)
      ( time timeout 900 python scripts/dev_shard_emulation.py $(echo ${ARGS:-8} | tr ',' ' ') ) > $OUT/${TAG}_shard_emulation.txt 2>&1
      cat $OUT/${TAG}_shard_emulation.txt ;;
    ncu)
      # ncu:NAME[:ENV=V,ENV=V]  NAME in large (default tables of config #5), shard8 (one rank of eight),
      # largedense, ar1, ar1dense.  Full-set capture of every kernel of one sweep + a launch list.
      W=${ARGS%% *}; W=${W:-large}
      case $W in
        large)      TGT="large on 3";  ENVS="" ;;
        large1band) TGT="large on 3";  ENVS="BANDS=1" ;;
        shard8)     TGT="large on 3";  ENVS="COL_SHARD=3/8" ;;
        largedense) TGT="large off 3"; ENVS="" ;;
        ar1)        TGT="ar1 on 3";    ENVS="" ;;
        ar1dense)   TGT="ar1 off 3";   ENVS="" ;;
      esac
      env $ENVS timeout 900 ncu --set full --clock-control none --import-source on \
          -k regex:'k_sweep|k_column_table|k_combine' --launch-skip ${NCU_SKIP:-2} -c ${NCU_COUNT:-4} -f \
          -o $OUT/${TAG}_ncu_$W python scripts/ncu_target.py $TGT > $OUT/${TAG}_ncu_$W.log 2>&1
      ncu -i $OUT/${TAG}_ncu_$W.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_summary.py > $OUT/${TAG}_ncu_$W.txt
      [ -z "${KEEP_REP:-}" ] && rm -f $OUT/${TAG}_ncu_$W.ncu-rep     # (gpurun brings back at most 64 MiB)
      env $ENVS timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
          --log-file $OUT/${TAG}_launches_$W.csv python scripts/ncu_target.py $TGT >> $OUT/${TAG}_ncu_$W.log 2>&1
      grep "== kernel\|gpu__time_duration\|dram__bytes_read.sum \|wavefronts.avg.pct\|fp64_cycles_active.avg.pct_of_peak_sustained_elapsed\|issue_active" $OUT/${TAG}_ncu_$W.txt | head -40 ;;
    sanitize)
      for TOOL in ${SAN_TOOLS:-memcheck racecheck}; do
        L=$OUT/${TAG}_sanitizer_${TOOL}_n$N.log
        if [ "$N" -gt 1 ]; then
          PORT=$((PORT+1))
          ( time timeout 1200 compute-sanitizer --tool $TOOL --target-processes all --report-api-errors no --error-exitcode 9 \
              $TR --master-port $PORT scripts/sanitizer_target.py ) > $L 2>&1
        else
          ( time timeout 1200 compute-sanitizer --tool $TOOL --error-exitcode 9 python scripts/sanitizer_target.py ) > $L 2>&1
        fi
        echo "exit: $?" >> $L
        grep -v "Warning\|warn" $L | tail -25
      done ;;
    timeline)
      # timeline[:ENV=V;ENV=V ...]  the e2e timeline of value_iteration (scripts/dev_e2e_timeline.py) once
      # per variant (environment settings separated by ';', variants by spaces), e.g.
      #   timeline:SDP_COLUMN_PIECES=0.3,0.3,0.3,0.1+SDP_SMALL_COMBINE=0
      : > $OUT/${TAG}_e2e_timeline.txt
      for V in ${ARGS:-DEFAULT=1}; do
        echo "== $V" >> $OUT/${TAG}_e2e_timeline.txt
        env $(echo "$V" | tr ';' ' ') timeout 200 python scripts/dev_e2e_timeline.py >> $OUT/${TAG}_e2e_timeline.txt 2>&1
      done
      cat $OUT/${TAG}_e2e_timeline.txt ;;
    emuopts)
      # emuopts:OPT=V,OPT=V[+OPT=V ...]  one-GPU and 1/8-shard kernel times (columns) per set of library options
      : > $OUT/${TAG}_emu_options.txt
      for V in $ARGS; do
        echo "== $V" >> $OUT/${TAG}_emu_options.txt
        OPTS=$V AXES=columns timeout 300 python scripts/dev_shard_emulation.py 8 2>&1 | grep "^N=" >> $OUT/${TAG}_emu_options.txt
      done
      cat $OUT/${TAG}_emu_options.txt ;;
    py)
      S=${ARGS%% *}; A=""; [ "$ARGS" != "$S" ] && A=${ARGS#* }
      ( time run_py scripts/$S $A ) > $OUT/${TAG}_$(basename $S .py).txt 2>&1
      tail -40 $OUT/${TAG}_$(basename $S .py).txt ;;
    *) echo "unknown step $STEP" ;;
  esac
done
ls -la $OUT | tail -30
