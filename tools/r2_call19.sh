#!/bin/bash
# Round 2, GPU call 19 (1 GPU): compute-sanitizer on the final kernels (memcheck, racecheck, initcheck), smoke()
set -u
OUT=gpurun_out/r2c19
mkdir -p "$OUT"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE ok')" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" >> "$OUT/rc.txt"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_frames.py > "$OUT/sanitizer_memcheck.log" 2>&1; echo "memcheck rc=$?" >> "$OUT/rc.txt"
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_frames.py --small > "$OUT/sanitizer_racecheck.log" 2>&1; echo "racecheck rc=$?" >> "$OUT/rc.txt"
timeout 400 compute-sanitizer --tool initcheck --error-exitcode 7 python tools/sanitize_frames.py --small > "$OUT/sanitizer_initcheck.log" 2>&1; echo "initcheck rc=$?" >> "$OUT/rc.txt"
cat "$OUT/rc.txt"; tail -2 "$OUT/smoke.log"; tail -3 "$OUT/sanitizer_memcheck.log"; tail -3 "$OUT/sanitizer_racecheck.log"; tail -3 "$OUT/sanitizer_initcheck.log"
