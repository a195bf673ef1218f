#!/bin/bash
# Round 2, GPU call 15: reproduce the order-dependent failure of the culling identity test with diagnostics
set -u
OUT=gpurun_out/r2c15
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_zz_quiet_zone.py -q -m gpu -p no:cacheprovider -k "fullsize or culling_never" > "$OUT/tests_a.log" 2>&1; echo "a rc=$?" >> "$OUT/rc.txt"
timeout 900 python -m pytest tests/test_gpu_units.py tests/test_gpu_zz_quiet_zone.py -q -m gpu -p no:cacheprovider -k "test_gpu_units or culling_never" > "$OUT/tests_b.log" 2>&1; echo "b rc=$?" >> "$OUT/rc.txt"
cat "$OUT/rc.txt"; grep -n "AssertionError: {" "$OUT"/tests_*.log | cut -c1-900; tail -2 "$OUT/tests_a.log" "$OUT/tests_b.log"
