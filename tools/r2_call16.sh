#!/bin/bash
# Round 2, GPU call 16 (1 GPU): the whole -m gpu suite with the TMA overlay, one-ray-per-CTA sample kernel and the 1.5 max_dist
# quiet radius as defaults; bench line; ncu captures of the shipped march (C3) and of the tile overlay; hash probe of this box
set -u
OUT=gpurun_out/r2c16
mkdir -p "$OUT"
python tools/dev/hash_probe.py "$OUT/plain" > "$OUT/hashes.txt" 2>&1
timeout 1500 python -m pytest tests -q -m gpu -rxXs -p no:cacheprovider --durations=12 > "$OUT/tests_gpu.log" 2>&1; echo "tests rc=$?" >> "$OUT/rc.txt"
timeout 700 python bench.py --steps 10 --warmup 3 > "$OUT/bench1.json" 2> "$OUT/bench1.err"; echo "bench rc=$?" >> "$OUT/rc.txt"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:linon_march_tiles -s 2 -c 1 -o "$OUT/march_c3" -f \
    python tools/render_once.py --frames 3 > "$OUT/ncu_march.log" 2>&1; echo "ncu march rc=$?" >> "$OUT/rc.txt"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:overlay_tile_kernel -s 2 -c 1 -o "$OUT/overlay_c3_tile" -f \
    python tools/render_once.py --frames 3 > "$OUT/ncu_overlay.log" 2>&1; echo "ncu overlay rc=$?" >> "$OUT/rc.txt"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 3 --no-configs --no-parity --no-cpu-baseline --c5-frames 0 > "$OUT/bench_under_ncu.log" 2>&1; echo "launch list rc=$?" >> "$OUT/rc.txt"
cat "$OUT/rc.txt" "$OUT/hashes.txt"; tail -5 "$OUT/tests_gpu.log"; tail -1 "$OUT/ncu_march.log" | cut -c1-300
