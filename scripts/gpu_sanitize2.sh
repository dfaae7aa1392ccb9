# compute-sanitizer over the kernels added in round 2 (scripts/san_round2.py): memcheck, racecheck (shared-memory hazards of
# the team phases), synccheck (named barriers)
# (synccheck wants the library built with -DDAQP_B200_SYNCCHECK:
#    python -c "from daqp_b200 import build; build.build(force=True, extra=['-DDAQP_B200_SYNCCHECK'])"   see team_ops.cuh)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  extra=""; [ $tool = racecheck ] && extra="--racecheck-report hazard --print-limit 2000"
  timeout 1500 compute-sanitizer --tool $tool $extra --error-exitcode 3 python scripts/san_round2.py > gpurun_out/san2_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Error|hazard" gpurun_out/san2_$tool.log | sort | uniq -c | sort -rn | head -8
done
