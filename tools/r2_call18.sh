#!/bin/bash
# Round 2, GPU call 18: sweep of LINON_QUIET_AHEAD (regular quiet steps between two plans on the final approach), C3 and C5 frame
set -u
OUT=gpurun_out/r2c18
mkdir -p "$OUT"
for a in 3 5 8 12 16 24; do
  LINON_DEV_DEFINES="-DLINON_QUIET_AHEAD=$a" timeout 200 python tools/render_once.py --frames 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ahead', $a, 'C3 march_ms', round(d['march_ms'],3), 'checks', d['culling_checks'], 'face', d['face_tests'], d['cubin_hash'])" >> "$OUT/sweep.txt"
done
for a in 3 8 16; do
  LINON_DEV_DEFINES="-DLINON_QUIET_AHEAD=$a" timeout 200 python tools/render_once.py --frames 3 --size 1920x1080 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ahead', $a, '1080p enhance march_ms', round(d['march_ms'],3))" >> "$OUT/sweep.txt"
done
cat "$OUT/sweep.txt"
