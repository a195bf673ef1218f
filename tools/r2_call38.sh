#!/bin/bash
# Round 2, GPU call 38 (last seconds of the budget): edited sample_colors on the device
set -u
OUT=gpurun_out/r2c38
mkdir -p "$OUT"
timeout 22 python -m pytest tests/test_gpu_units.py -q -m gpu -p no:cacheprovider -k "edited_sample_colors" > "$OUT/tests.log" 2>&1; echo "tests rc=$?"
tail -3 "$OUT/tests.log"
