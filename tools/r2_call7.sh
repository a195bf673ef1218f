#!/bin/bash
# Round 2, GPU call 7 (1 GPU):  gpurun --timeout 1800 -- 'bash tools/r2_call7.sh'
# Overlay TMA candidates (alignment fixed, interior fast path), the march with the boxes-clear check path: time, parity, ncu.
set -u
OUT=gpurun_out/r2c7
mkdir -p "$OUT"
timeout 200 tools/ubench/overlay_variants > "$OUT/overlay_variants.log" 2>&1; echo "overlay variants rc=$?" >> "$OUT/rc.txt"
timeout 200 tools/ubench/overlay_variants 3840 2160 5 >> "$OUT/overlay_variants.log" 2>&1
timeout 100 tools/ubench/overlay_variants 1000 777 1 4 >> "$OUT/overlay_variants.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"v_tma|v0_shipped|v_copy_floor" -o "$OUT/overlay_variants" -f \
    tools/ubench/overlay_variants 3840 2160 1 0 > "$OUT/ncu_overlay_variants.log" 2>&1; echo "ncu overlay rc=$?" >> "$OUT/rc.txt"
timeout 300 python tools/render_once.py --frames 4 > "$OUT/render_once_c3.json" 2>&1
timeout 300 python tools/render_once.py --frames 4 --size 8192x8192 > "$OUT/render_once_c5.json" 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_zz_quiet_zone.py -q -m gpu -x -p no:cacheprovider --durations=8 > "$OUT/tests_parity.log" 2>&1; echo "tests rc=$?" >> "$OUT/rc.txt"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:linon_march_tiles -s 2 -c 1 -o "$OUT/march_c3" -f \
    python tools/render_once.py --frames 3 > "$OUT/ncu_march.log" 2>&1; echo "ncu march rc=$?" >> "$OUT/rc.txt"
cat "$OUT/rc.txt"; tail -3 "$OUT/tests_parity.log"; cat "$OUT/overlay_variants.log" | grep -v "strip\|pair"; tail -1 "$OUT/render_once_c3.json" | cut -c1-400; tail -1 "$OUT/render_once_c5.json" | cut -c1-300
