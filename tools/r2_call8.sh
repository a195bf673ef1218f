#!/bin/bash
# Round 2, GPU call 8: culling on/off identity probe (old and new check path), overlay tile kernel tests, C3 time
set -u
OUT=gpurun_out/r2c8
mkdir -p "$OUT"
timeout 300 python tools/dev/cull_diff.py > "$OUT/cull_diff_new.json" 2> "$OUT/cull_diff_new.err"; echo "cull new rc=$?" >> "$OUT/rc.txt"
timeout 300 python tools/dev/cull_diff.py tools/dev/_old > "$OUT/cull_diff_old.json" 2> "$OUT/cull_diff_old.err"; echo "cull old rc=$?" >> "$OUT/rc.txt"
timeout 600 python -m pytest tests/test_gpu_zz_quiet_zone.py -q -m gpu -p no:cacheprovider -k "overlay" > "$OUT/tests_overlay.log" 2>&1; echo "tests overlay rc=$?" >> "$OUT/rc.txt"
timeout 300 python tools/render_once.py --frames 4 > "$OUT/render_once_c3.json" 2>&1
cat "$OUT/rc.txt"; tail -3 "$OUT/tests_overlay.log"; cat "$OUT/cull_diff_new.json" | cut -c1-3000; echo; cat "$OUT/cull_diff_old.json" | cut -c1-3000; tail -1 "$OUT/render_once_c3.json" | cut -c1-300
