#!/bin/bash
# One gpurun call that collects the evidence kept under profiles/ (tests, bench, launch lists, ncu --set full captures,
# sanitizer logs).  Usage on the GPU box:  bash scripts/gpu_evidence.sh <tag> [tests] [bench] [launches] [ncu] [sanitize]
# Everything lands in gpurun_out/<tag>_*; numbers printed by a run under ncu / compute-sanitizer are never bench values.
tag=${1:-r2}; shift
what=${*:-tests bench launches ncu sanitize}
out=gpurun_out; mkdir -p $out
has() { [[ " $what " == *" $1 "* ]]; }
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"

if has tests; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_gputests.log 2>&1; echo "tests exit $?" | tee -a $out/${tag}_gputests.log
fi
if has bench; then
  timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?"
  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench.err; echo "ref exit $?"
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches_bench.csv \
      python bench.py --steps 16 --warmup 3 --preload 0.02 --mix 8 --no-dense > $out/${tag}_launches_bench.log 2>&1; echo "launches exit $?"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file $out/${tag}_launches_rollout.csv \
      python scripts/prof_rollout.py > $out/${tag}_launches_rollout.log 2>&1; echo "rollout launches exit $?"
fi
if has ncu; then
  timeout 600 $NCU -k regex:mnv_env_kernel -c 2 -f -o $out/${tag}_step python scripts/prof_step.py > $out/${tag}_ncu.log 2>&1
  timeout 600 $NCU -k regex:mnv_env_dense -c 2 -f -o $out/${tag}_dense python scripts/prof_step.py 16384 dense >> $out/${tag}_ncu.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:iqn_ --launch-skip 12 -c 4 -f -o $out/${tag}_update python scripts/prof_update.py >> $out/${tag}_ncu.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:iqn_ --launch-skip 3 -c 4 -f -o $out/${tag}_act python scripts/prof_act_tc.py >> $out/${tag}_ncu.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mnv_reset --launch-skip 4 -c 2 -f -o $out/${tag}_reset python scripts/reset_lab.py >> $out/${tag}_ncu.log 2>&1
  echo "ncu done"
fi
if has sanitize; then
  # memcheck also covers one captured rollout + learn vector step (control-block entry points, tcgen05 act kernel); the race /
  # sync tools run the env + host-boundary paths only (they slow the IQN kernels down by orders of magnitude)
  timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_env.py > $out/${tag}_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"
  timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_env.py --no-iqn > $out/${tag}_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"
  timeout 600 compute-sanitizer --tool synccheck python scripts/sanitize_env.py --no-iqn > $out/${tag}_sanitizer_synccheck.log 2>&1; echo "synccheck exit $?"
fi
ls -la $out | tail -40
