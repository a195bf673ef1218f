#!/bin/bash
# Round 2, GPU call 21: window resolve of wall hits in quiet runs - time (C3, C5 frame, 1080p), sweep, identity with culling off
set -u
OUT=gpurun_out/r2c21
mkdir -p "$OUT"
for defs in "-DLINON_WINDOW=0" "-DLINON_WINDOW=32" "-DLINON_WINDOW=64" "-DLINON_WINDOW=32 -DLINON_WINDOW_CHECKS=8" "-DLINON_WINDOW=16 -DLINON_WINDOW_CHECKS=4" "-DLINON_WINDOW=64 -DLINON_WINDOW_CHECKS=8"; do
  LINON_DEV_DEFINES="$defs" timeout 200 python tools/render_once.py --frames 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$defs', 'C3 march_ms', round(d['march_ms'],3), 'checks', d['culling_checks'], 'face', d['face_tests'], 'regs', d['regs'])" >> "$OUT/sweep.txt"
done
timeout 200 python tools/render_once.py --frames 3 --size 8192x8192 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C5 march_ms', round(d['march_ms'],3))" >> "$OUT/sweep.txt"
timeout 300 python tools/dev/cull_diff.py > "$OUT/cull_diff.json" 2> "$OUT/cull_diff.err"; echo "cull rc=$?" >> "$OUT/rc.txt"
python -c "
import json
d=json.load(open('$OUT/cull_diff.json'))
print({k:(v['cull1_vs_cull0']['hit_diff'], v['cull1_vs_cull0']['colour_diff'], v['cull1_vs_cull0']['endpoint_diff']) for k,v in d.items() if k!='pkg'})" >> "$OUT/sweep.txt"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_zz_quiet_zone.py -q -m gpu -p no:cacheprovider -x > "$OUT/tests.log" 2>&1; echo "tests rc=$?" >> "$OUT/rc.txt"
cat "$OUT/sweep.txt" "$OUT/rc.txt"; tail -3 "$OUT/tests.log"
