#!/bin/bash
# Round 2, GPU call 11: one ray per CTA in the sample-ray kernel, initcheck, unit tests, the bench line with the TMA overlay
set -u
OUT=gpurun_out/r2c11
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_units.py tests/test_gpu_zz_quiet_zone.py -q -m gpu -p no:cacheprovider -x --durations=6 > "$OUT/tests_units.log" 2>&1; echo "tests rc=$?" >> "$OUT/rc.txt"
timeout 700 python bench.py --steps 10 --warmup 3 > "$OUT/bench1.json" 2> "$OUT/bench1.err"; echo "bench rc=$?" >> "$OUT/rc.txt"
timeout 400 compute-sanitizer --tool initcheck python tools/dev/cull_diff.py . 96x64 > "$OUT/initcheck.log" 2>&1; echo "initcheck rc=$?" >> "$OUT/rc.txt"
cat "$OUT/rc.txt"; tail -3 "$OUT/tests_units.log"; tail -4 "$OUT/initcheck.log" | cut -c1-300; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c11/bench1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['ncu_profile'], d['roofline_overlay']['frac'], d['roofline_overlay']['launch_ms'], d['roofline_overlay']['kernel'][:30])
print({k:(v['ms_per_step'],) for k,v in d['variants'].items()}, {k:v.get('green') for k,v in d['parity'].items() if isinstance(v,dict)})
for l in d['configs']['lines']:
    print(l['config'], l['function'], l['variant'][:30], l['march_ms'], l['sample_rays_ms'], l.get('exponents_above_threshold'), l.get('parity',{}).get('green'))
P
