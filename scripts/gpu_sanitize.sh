# compute-sanitizer passes over a few small parity tests (memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards)
mkdir -p gpurun_out
K='golden and (g1_n10_m20 or g1_n12_m30_ms6 or duplicate_rows or warm_wrong or equalities)'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -q -m gpu -x -k "$K" > gpurun_out/san_mem.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_mem.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -q -m gpu -x -k "workspace_sequence and (n10 or soft)" > gpurun_out/san_mem2.log 2>&1; echo "memcheck(ws) rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_mem2.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --racecheck-report hazard --print-limit 5000 --error-exitcode 3 python -m pytest tests -q -m gpu -x -k "golden and (g1_n10_m20 or duplicate_rows)" > gpurun_out/san_race.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san_race.log | tail -3; grep -E "hazard" gpurun_out/san_race.log | sort | uniq -c | sort -rn | head -12
