#!/bin/bash
# Round 2, GPU call 22: window resolve with the refined start - time, identity with culling off
set -u
OUT=gpurun_out/r2c22
mkdir -p "$OUT"
for defs in "-DLINON_WINDOW_REFINE=0" "-DLINON_WINDOW_REFINE=1" "-DLINON_WINDOW_REFINE=1 -DLINON_WINDOW=64" "-DLINON_WINDOW_REFINE=1 -DLINON_WINDOW_CHECKS=2" "-DLINON_WINDOW_REFINE=1 -DLINON_WINDOW_CHECKS=3" "-DLINON_WINDOW_REFINE=1 -DLINON_WINDOW=48 -DLINON_WINDOW_CHECKS=6"; do
  LINON_DEV_DEFINES="$defs" timeout 200 python tools/render_once.py --frames 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$defs', 'C3 march_ms', round(d['march_ms'],3), 'checks', d['culling_checks'], 'face', d['face_tests'], 'regs', d['regs'])" >> "$OUT/sweep.txt"
done
timeout 300 python tools/dev/cull_diff.py > "$OUT/cull_diff.json" 2> "$OUT/cull_diff.err"; echo "cull rc=$?" >> "$OUT/rc.txt"
python -c "
import json
d=json.load(open('$OUT/cull_diff.json'))
print({k:(v['cull1_vs_cull0']['hit_diff'], v['cull1_vs_cull0']['colour_diff'], v['cull1_vs_cull0']['endpoint_diff']) for k,v in d.items() if k!='pkg'})" >> "$OUT/sweep.txt"
cat "$OUT/sweep.txt" "$OUT/rc.txt"
