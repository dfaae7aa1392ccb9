mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --racecheck-report hazard --print-limit 5000 python -m pytest tests -q -m gpu -x -k "(golden and (g1_n10_m20 or duplicate_rows or equalities or warm_wrong)) or (workspace_sequence and (n10 or soft)) or soft_constraints" > gpurun_out/san_race.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san_race.log | tail -3
grep -E "in (ldp_kernel|common|setup_kernel|update_kernel).cuh:[0-9]+" -o gpurun_out/san_race.log | sort | uniq -c | sort -rn | head -30
