#!/bin/bash
# Round 2, second GPU call (1 GPU):  gpurun --timeout 1800 -- 'bash tools/r2_call2.sh'
# The whole -m gpu suite with the parameterised mirage family + the cancellation-free eta + linon_sharded_* (shards on one
# device), the bench line, the ncu capture of the NEW shipped march binary, and the overlay candidates side by side.
set -u
OUT=gpurun_out/r2c2
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -q -m gpu -rxXs -p no:cacheprovider --durations=15 > "$OUT/tests_gpu.log" 2>&1; echo "tests rc=$?" >> "$OUT/rc.txt"
timeout 600 python bench.py --steps 10 --warmup 3 > "$OUT/bench1.json" 2> "$OUT/bench1.err"; echo "bench rc=$?" >> "$OUT/rc.txt"
timeout 120 tools/ubench/overlay_variants > "$OUT/overlay_variants.log" 2>&1; echo "overlay variants rc=$?" >> "$OUT/rc.txt"
timeout 120 tools/ubench/overlay_variants 3840 2160 5 >> "$OUT/overlay_variants.log" 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:linon_march_tiles -s 2 -c 1 -o "$OUT/march_c3_quiet" -f \
    python tools/render_once.py --frames 3 > "$OUT/ncu_march.log" 2>&1; echo "ncu march rc=$?" >> "$OUT/rc.txt"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:linon_march_tiles -s 2 -c 1 -o "$OUT/march_c2_adaptive_q2" -f \
    python tools/render_once.py --size 1920x1080 --flags 0 --frames 3 --quiet 2 > "$OUT/ncu_march_adaptive_q2.log" 2>&1; echo "ncu adaptive q2 rc=$?" >> "$OUT/rc.txt"
ls -la "$OUT"; cat "$OUT/rc.txt"; tail -5 "$OUT/tests_gpu.log"
