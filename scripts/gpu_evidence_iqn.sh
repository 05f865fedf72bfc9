bash scripts/gpu_evidence.sh r2d tests bench
out=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file $out/r2d_launches_rollout.csv python scripts/prof_rollout.py > $out/r2d_launches_rollout.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:iqn_ --launch-skip 10 -c 4 -f -o $out/r2d_update python scripts/prof_update.py > $out/r2d_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:iqn_ --launch-skip 2 -c 4 -f -o $out/r2d_act python scripts/prof_act_tc.py >> $out/r2d_ncu.log 2>&1
tail -3 $out/r2d_gputests.log
