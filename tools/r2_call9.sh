#!/bin/bash
# Round 2, GPU call 9: is the culling identity flaky?  The failing test in a loop, the probe in a loop, racecheck on the configuration
set -u
OUT=gpurun_out/r2c9
mkdir -p "$OUT"
for i in 1 2 3 4 5 6; do
  timeout 200 python -m pytest tests/test_gpu_zz_quiet_zone.py -q -m gpu -p no:cacheprovider -k "culling_never_changes" 2>&1 | tail -2 >> "$OUT/loop_tests.log"
done
for i in 1 2 3 4; do
  timeout 200 python tools/dev/cull_diff.py 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print({k:(v['cull1_vs_cull0']['hit_diff'], v['cull1_vs_cull0']['colour_diff'], v['cull1_vs_cull0']['endpoint_diff']) for k,v in d.items() if k!='pkg'})" >> "$OUT/loop_probe_new.log"
  timeout 200 python tools/dev/cull_diff.py tools/dev/_old 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print({k:(v['cull1_vs_cull0']['hit_diff'], v['cull1_vs_cull0']['colour_diff'], v['cull1_vs_cull0']['endpoint_diff']) for k,v in d.items() if k!='pkg'})" >> "$OUT/loop_probe_old.log"
done
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_quiet_zone.py -q -m gpu -p no:cacheprovider -x > "$OUT/tests_parity.log" 2>&1; echo "tests rc=$?" >> "$OUT/rc.txt"
cat "$OUT/loop_tests.log" "$OUT/loop_probe_new.log" "$OUT/loop_probe_old.log" "$OUT/rc.txt"; tail -3 "$OUT/tests_parity.log"
