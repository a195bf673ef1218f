#!/bin/bash
# Round 2, GPU call 36: the "depth" shading edit again (its production-math colour check was ill-posed: t / max_dist in an adaptive frame)
set -u
OUT=gpurun_out/r2c36
mkdir -p "$OUT"
timeout 60 python -m pytest tests/test_gpu_units.py -q -m gpu -p no:cacheprovider -rs -k "shading_block and (depth or declined)" > "$OUT/tests.log" 2>&1; echo "tests rc=$?"
tail -4 "$OUT/tests.log"
