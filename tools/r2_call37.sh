#!/bin/bash
# Round 2, GPU call 37: literal edits of main_view (focal length, viewport height, reach of linear_mode) on the device
set -u
OUT=gpurun_out/r2c37
mkdir -p "$OUT"
timeout 45 python -m pytest tests/test_gpu_units.py -q -m gpu -p no:cacheprovider -rs -k "shading_block and (camera or linear_reach)" > "$OUT/tests.log" 2>&1; echo "tests rc=$?"
tail -4 "$OUT/tests.log"
