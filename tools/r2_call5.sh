#!/bin/bash
# Round 2, GPU call 5 (1 GPU):  gpurun --timeout 1800 -- 'bash tools/r2_call5.sh'
# Full -m gpu suite after: march-level hot reload, display hand-off, advisor fixes, noise floor of the step control, quiet zone
# level 2 by default; bench line; ncu captures of the shipped march (C3) and of an adaptive frame with the final step control.
set -u
OUT=gpurun_out/r2c5
mkdir -p "$OUT"
timeout 1300 python -m pytest tests -q -m gpu -rxXs -p no:cacheprovider --durations=12 > "$OUT/tests_gpu.log" 2>&1; echo "tests rc=$?" >> "$OUT/rc.txt"
timeout 600 python bench.py --steps 10 --warmup 3 > "$OUT/bench1.json" 2> "$OUT/bench1.err"; echo "bench rc=$?" >> "$OUT/rc.txt"
timeout 300 python tools/render_once.py --frames 3 > "$OUT/render_once_c3.json" 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:linon_march_tiles -s 2 -c 1 -o "$OUT/march_c2_adaptive" -f \
    python tools/render_once.py --size 1920x1080 --flags 0 --frames 3 > "$OUT/ncu_march_adaptive.log" 2>&1; echo "ncu adaptive rc=$?" >> "$OUT/rc.txt"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_frames.py > "$OUT/sanitizer_memcheck.log" 2>&1; echo "memcheck rc=$?" >> "$OUT/rc.txt"
cat "$OUT/rc.txt"; tail -4 "$OUT/tests_gpu.log"; cat "$OUT/render_once_c3.json" | tail -1 | cut -c1-300
