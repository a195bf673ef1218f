#!/bin/bash
# Round 2, GPU call 10: initcheck / racecheck of the fixed-step tile kernel (culling on and off, quiet zone), dirty-memory stress
set -u
OUT=gpurun_out/r2c10
mkdir -p "$OUT"
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:${LD_LIBRARY_PATH:-}
timeout 500 compute-sanitizer --tool initcheck --track-unused-memory no python tools/dev/cull_diff.py . 96x64 > "$OUT/initcheck.log" 2>&1; echo "initcheck rc=$?" >> "$OUT/rc.txt"
timeout 500 compute-sanitizer --tool racecheck python tools/dev/cull_diff.py . 96x64 > "$OUT/racecheck.log" 2>&1; echo "racecheck rc=$?" >> "$OUT/rc.txt"
timeout 500 compute-sanitizer --tool synccheck python tools/dev/cull_diff.py . 96x64 > "$OUT/synccheck.log" 2>&1; echo "synccheck rc=$?" >> "$OUT/rc.txt"
for i in $(seq 1 12); do
  CULL_DIFF_DIRTY=1 timeout 100 python tools/dev/cull_diff.py . 160x120 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print({k:(v['cull1_vs_cull0']['hit_diff'], v['cull1_vs_cull0']['colour_diff'], v['cull1_vs_cull0']['endpoint_diff']) for k,v in d.items() if k!='pkg'})" >> "$OUT/loop_dirty.log"
done
cat "$OUT/rc.txt"; tail -6 "$OUT/initcheck.log" | cut -c1-300; tail -4 "$OUT/racecheck.log" | cut -c1-300; tail -4 "$OUT/synccheck.log" | cut -c1-300; sort "$OUT/loop_dirty.log" | uniq -c
