#!/bin/bash
# First GPU call of a round that starts with unmeasured, opt-in paths (run under gpurun from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
# Everything lands in gpurun_out/first/ ; nothing here changes defaults.  Each step has its own time limit.
set -u
OUT=gpurun_out/first
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
# 1. the opt-in paths on hardware (quiet zone, peer stores in one process, direct overlay kernel), with the reasons of every xfail / xpass
timeout 900 python -m pytest tests/test_gpu_zz_quiet_zone.py -q -m gpu -rxXs -p no:cacheprovider --runxfail > "$OUT/quiet_zone_tests.log" 2>&1
# 2. C3: default path (with the probe), quiet zone, quiet zone with the tighter sigmoid radius
timeout 600 python bench.py --steps 5 --warmup 3 > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
timeout 300 python bench.py --steps 5 --warmup 3 --quiet-zone 1 --no-cpu-baseline > "$OUT/bench_quiet.json" 2> "$OUT/bench_quiet.err"
# 3. C1 / C2 / C4 with the option at level 2 (adaptive frames) next to the default
timeout 300 python tools/configs.py --out "$OUT/configs_default.json" > "$OUT/configs_default.log" 2>&1
timeout 300 python tools/configs.py --quiet-zone 2 --out "$OUT/configs_quiet2.json" > "$OUT/configs_quiet2.log" 2>&1
timeout 300 python tools/configs.py --overlay-kernel 1 --out "$OUT/configs_overlay1.json" > "$OUT/configs_overlay1.log" 2>&1
# 4. launch list + one full capture of the quiet-zone march (never a bench value)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file "$OUT/launches_quiet.csv" \
    python bench.py --steps 2 --warmup 1 --quiet-zone 1 --no-cpu-baseline > "$OUT/ncu_launches.log" 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linon_march_tiles -s 2 -c 1 -o "$OUT/march_quiet" -f \
    python bench.py --steps 1 --warmup 1 --quiet-zone 1 --no-cpu-baseline > "$OUT/ncu_full.log" 2>&1
# 5. the direct overlay kernel under ncu: one full capture (bench.py's probe child runs it on the C3 frame)
timeout 600 ncu --set full --clock-control none --import-source on --target-processes all -k regex:overlay_direct_kernel -s 3 -c 1 \
    -o "$OUT/overlay_direct" -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > "$OUT/ncu_overlay_direct.log" 2>&1
ls -la "$OUT"
