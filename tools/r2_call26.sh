#!/bin/bash
# Round 2, GPU call 26: collapsed quiet runs (advance_coord) and one ray per CTA in the sample kernel - time, identity, tests
set -u
OUT=gpurun_out/r2c26
mkdir -p "$OUT"
for defs in "-DLINON_COLLAPSE_RUNS=0" "-DLINON_COLLAPSE_RUNS=1"; do
  LINON_DEV_DEFINES="$defs" timeout 200 python tools/render_once.py --frames 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$defs', 'C3 march_ms', round(d['march_ms'],3), 'sample_rays_ms', round(d['sample_rays_ms'],3), 'regs', d['regs'])" >> "$OUT/sweep.txt"
  LINON_DEV_DEFINES="$defs" timeout 200 python tools/render_once.py --frames 3 --size 8192x8192 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$defs', 'C5 march_ms', round(d['march_ms'],3))" >> "$OUT/sweep.txt"
  LINON_DEV_DEFINES="$defs" timeout 200 python tools/render_once.py --frames 3 --size 1920x1080 --flags 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$defs', '1080p adaptive-off march_ms', round(d['march_ms'],3), 'sample_rays_ms', round(d['sample_rays_ms'],3))" >> "$OUT/sweep.txt"
done
timeout 300 python tools/dev/cull_diff.py > "$OUT/cull_diff.json" 2> "$OUT/cull_diff.err"; echo "cull rc=$?" >> "$OUT/rc.txt"
python -c "
import json
d=json.load(open('$OUT/cull_diff.json'))
print({k:(v['cull1_vs_cull0']['hit_diff'], v['cull1_vs_cull0']['colour_diff'], v['cull1_vs_cull0']['endpoint_diff']) for k,v in d.items() if k!='pkg'})" >> "$OUT/sweep.txt"
timeout 1500 python -m pytest tests -q -m gpu -rxXs -p no:cacheprovider --durations=6 > "$OUT/tests_gpu.log" 2>&1; echo "tests rc=$?" >> "$OUT/rc.txt"
cat "$OUT/sweep.txt" "$OUT/rc.txt"; tail -4 "$OUT/tests_gpu.log"
