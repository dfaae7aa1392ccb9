#!/bin/bash
# knob sweep of the team kernel on C4 (13320 problems = 30 waves), one JSON line per setting
mkdir -p gpurun_out; : > gpurun_out/tune_c4.jsonl
for cfg in "0 3" "8 3" "16 3" "48 3" "64 3" "80 3" "112 3" "0 2" "48 2"; do
  set -- $cfg
  DAQP_B200_TUNE=$1 DAQP_B200_WARPS=$2 timeout 300 python scripts/bench_c4.py --n 13320 --reps 1 2>/dev/null | sed "s/^/{\"warps\": $2, \"r\": /; s/$/}/" >> gpurun_out/tune_c4.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/tune_c4.jsonl'):
    d = json.loads(l); r = d['r']
    print(f"tune={r['tune']:>4s} warps={d['warps']} cold solve {r['C4_cold']['solve_ms']:8.1f} ms  warm solve {r['C4_warm']['solve_ms']:8.1f} ms  setup {r['C4_cold']['setup_ms']:.1f}")
PY
