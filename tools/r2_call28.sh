#!/bin/bash
# Round 2, GPU call 28 (1 GPU): final defaults (collapsed quiet runs, window resolve, one ray per CTA in the sample kernel): identity with
# culling off, the whole -m gpu suite, the bench line, ncu capture of the shipped march
set -u
OUT=gpurun_out/r2c28
mkdir -p "$OUT"
python tools/dev/hash_probe.py "$OUT/plain" > "$OUT/hashes.txt" 2>&1
timeout 300 python tools/dev/cull_diff.py > "$OUT/cull_diff.json" 2> "$OUT/cull_diff.err"; echo "cull rc=$?" >> "$OUT/rc.txt"
python -c "
import json
d=json.load(open('$OUT/cull_diff.json'))
print({k:(v['cull1_vs_cull0']['hit_diff'], v['cull1_vs_cull0']['colour_diff'], v['cull1_vs_cull0']['endpoint_diff']) for k,v in d.items() if k!='pkg'})" >> "$OUT/sweep.txt"
timeout 1500 python -m pytest tests -q -m gpu -rxXs -p no:cacheprovider --durations=8 > "$OUT/tests_gpu.log" 2>&1; echo "tests rc=$?" >> "$OUT/rc.txt"
timeout 700 python bench.py --steps 10 --warmup 3 > "$OUT/bench1.json" 2> "$OUT/bench1.err"; echo "bench rc=$?" >> "$OUT/rc.txt"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:linon_march_tiles -s 2 -c 1 -o "$OUT/march_c3" -f \
    python tools/render_once.py --frames 3 > "$OUT/ncu_march.log" 2>&1; echo "ncu march rc=$?" >> "$OUT/rc.txt"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 3 --no-configs --no-parity --no-cpu-baseline --c5-frames 0 > "$OUT/bench_under_ncu.log" 2>&1; echo "launch list rc=$?" >> "$OUT/rc.txt"
cat "$OUT/sweep.txt" "$OUT/rc.txt" "$OUT/hashes.txt"; tail -4 "$OUT/tests_gpu.log"
