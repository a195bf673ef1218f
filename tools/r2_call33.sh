#!/bin/bash
# Round 2, GPU call 33: the collapsed-runs identity test at its final threshold (the frames were already bit-equal in call 32)
set -u
OUT=gpurun_out/r2c33
mkdir -p "$OUT"
timeout 150 python -m pytest tests/test_gpu_zz_quiet_zone.py -q -m gpu -p no:cacheprovider -k "collapsed_runs" > "$OUT/tests.log" 2>&1; echo "tests rc=$?"
tail -3 "$OUT/tests.log"
