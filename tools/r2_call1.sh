#!/bin/bash
# Round 2, first GPU call (run under gpurun from the repo root):  gpurun --timeout 1500 -- 'bash tools/r2_call1.sh'
# The flipped defaults (quiet zone level 1, direct overlay) on hardware: tests, the bench line with its parity / variants /
# configs blocks, compute-sanitizer on small frames, launch list + full ncu captures of the shipped kernels.
set -u
OUT=gpurun_out/r2c1
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/smi.txt" 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" >> "$OUT/rc.txt"
timeout 600 python bench.py --steps 5 --warmup 3 > "$OUT/bench1.json" 2> "$OUT/bench1.err"; echo "bench rc=$?" >> "$OUT/rc.txt"
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_zz_quiet_zone.py -q -m gpu -rxXs -p no:cacheprovider -x --durations=10 > "$OUT/tests_new.log" 2>&1; echo "tests rc=$?" >> "$OUT/rc.txt"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_frames.py > "$OUT/sanitizer_memcheck.log" 2>&1; echo "memcheck rc=$?" >> "$OUT/rc.txt"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_frames.py --small > "$OUT/sanitizer_racecheck.log" 2>&1; echo "racecheck rc=$?" >> "$OUT/rc.txt"
# launch list of the bench command (never a bench value) and full captures of the shipped march + overlay, and of an adaptive frame
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file "$OUT/launches_bench.csv" \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-configs --c5-frames 0 > "$OUT/ncu_launches.log" 2>&1; echo "launches rc=$?" >> "$OUT/rc.txt"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:linon_march_tiles -s 2 -c 1 -o "$OUT/march_c3_quiet" -f \
    python tools/render_once.py --frames 3 > "$OUT/ncu_march.log" 2>&1; echo "ncu march rc=$?" >> "$OUT/rc.txt"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:overlay -s 2 -c 1 -o "$OUT/overlay_c3_direct" -f \
    python tools/render_once.py --frames 3 > "$OUT/ncu_overlay.log" 2>&1; echo "ncu overlay rc=$?" >> "$OUT/rc.txt"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linon_march_tiles -s 2 -c 1 -o "$OUT/march_c2_adaptive" -f \
    python tools/render_once.py --size 1920x1080 --flags 0 --frames 3 > "$OUT/ncu_march_adaptive.log" 2>&1; echo "ncu adaptive rc=$?" >> "$OUT/rc.txt"
ls -la "$OUT"; cat "$OUT/rc.txt"
