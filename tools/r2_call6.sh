#!/bin/bash
# Round 2, GPU call 6 (1 GPU):  gpurun --timeout 1500 -- 'bash tools/r2_call6.sh'
# NVRTC pinned by path (bench and tests now time the same cubin), overlay TMA candidates, outline probe, the two fixed tests.
set -u
OUT=gpurun_out/r2c6
mkdir -p "$OUT"
timeout 200 tools/ubench/overlay_variants > "$OUT/overlay_variants.log" 2>&1; echo "overlay variants rc=$?" >> "$OUT/rc.txt"
timeout 200 tools/ubench/overlay_variants 3840 2160 5 >> "$OUT/overlay_variants.log" 2>&1
timeout 100 tools/ubench/overlay_variants 1000 777 1 4 >> "$OUT/overlay_variants.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"v_tma|v0_shipped|v_copy_floor" -o "$OUT/overlay_variants" -f \
    tools/ubench/overlay_variants 3840 2160 1 0 > "$OUT/ncu_overlay_variants.log" 2>&1; echo "ncu overlay rc=$?" >> "$OUT/rc.txt"
timeout 200 python tools/dev/outline_probe.py > "$OUT/outline_probe.json" 2> "$OUT/outline_probe.err"; echo "outline probe rc=$?" >> "$OUT/rc.txt"
timeout 400 python -m pytest tests/test_gpu_units.py -q -m gpu -p no:cacheprovider -k "march_level_constants or outline" > "$OUT/tests_fixed.log" 2>&1; echo "tests rc=$?" >> "$OUT/rc.txt"
timeout 600 python bench.py --steps 10 --warmup 3 > "$OUT/bench1.json" 2> "$OUT/bench1.err"; echo "bench rc=$?" >> "$OUT/rc.txt"
cat "$OUT/rc.txt"; tail -3 "$OUT/tests_fixed.log"; cat "$OUT/overlay_variants.log"; cat "$OUT/outline_probe.json" | cut -c1-1500
