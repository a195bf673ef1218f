#!/bin/bash
# Round 2, GPU call 27: run length from which quiet runs are collapsed
set -u
OUT=gpurun_out/r2c27
mkdir -p "$OUT"
for m in 32 64 96 160 256 512; do
  LINON_DEV_DEFINES="-DLINON_COLLAPSE_MIN=$m" timeout 200 python tools/render_once.py --frames 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('min', $m, 'C3 march_ms', round(d['march_ms'],3), 'sample_rays_ms', round(d['sample_rays_ms'],3))" >> "$OUT/sweep.txt"
done
for m in 96 256; do
  LINON_DEV_DEFINES="-DLINON_COLLAPSE_MIN=$m" timeout 200 python tools/render_once.py --frames 3 --size 8192x8192 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('min', $m, 'C5 march_ms', round(d['march_ms'],3))" >> "$OUT/sweep.txt"
  LINON_DEV_DEFINES="-DLINON_COLLAPSE_MIN=$m" timeout 200 python tools/render_once.py --frames 3 --size 1920x1080 --flags 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('min', $m, '1080p adaptive-off march_ms', round(d['march_ms'],3), 'sample_rays_ms', round(d['sample_rays_ms'],3))" >> "$OUT/sweep.txt"
done
timeout 300 python tools/dev/cull_diff.py > "$OUT/cull_diff.json" 2> "$OUT/cull_diff.err"
python -c "
import json
d=json.load(open('$OUT/cull_diff.json'))
print({k:(v['cull1_vs_cull0']['hit_diff'], v['cull1_vs_cull0']['colour_diff'], v['cull1_vs_cull0']['endpoint_diff']) for k,v in d.items() if k!='pkg'})" >> "$OUT/sweep.txt"
cat "$OUT/sweep.txt"
