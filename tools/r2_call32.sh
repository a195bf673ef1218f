#!/bin/bash
# Round 2, GPU call 32 (last): collapsed runs vs step-by-step frames on the device, smoke()
set -u
OUT=gpurun_out/r2c32
mkdir -p "$OUT"
timeout 200 python -m pytest tests/test_gpu_zz_quiet_zone.py -q -m gpu -p no:cacheprovider -k "collapsed_runs" > "$OUT/tests.log" 2>&1; echo "tests rc=$?"
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE ok')" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"
tail -3 "$OUT/tests.log"; tail -2 "$OUT/smoke.log" | cut -c1-300
