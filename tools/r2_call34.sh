#!/bin/bash
# Round 2, GPU call 34: compute-sanitizer over the kernels as shipped (window resolve + collapsed runs came after call 19's logs)
set -u
OUT=gpurun_out/r2c34
mkdir -p "$OUT"
timeout 90 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_frames.py > "$OUT/sanitizer_memcheck.log" 2>&1; echo "memcheck rc=$?" >> "$OUT/rc.txt"
timeout 100 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_frames.py --small --enhance > "$OUT/sanitizer_racecheck.log" 2>&1; echo "racecheck rc=$?" >> "$OUT/rc.txt"
timeout 40 compute-sanitizer --tool initcheck --error-exitcode 7 python tools/sanitize_frames.py --small --enhance > "$OUT/sanitizer_initcheck.log" 2>&1; echo "initcheck rc=$?" >> "$OUT/rc.txt"
cat "$OUT/rc.txt"; tail -2 "$OUT/sanitizer_memcheck.log"; tail -2 "$OUT/sanitizer_racecheck.log"; tail -2 "$OUT/sanitizer_initcheck.log"
