#!/bin/bash
# Round 2, GPU call 13: is the NVRTC output the same in every process of this box?  (hash of the tile kernel, cubins kept)
set -u
OUT=gpurun_out/r2c13
mkdir -p "$OUT"
lscpu | grep -E "Model name|Flags" | cut -c1-600 > "$OUT/cpu.txt"
for i in 1 2; do python tools/dev/hash_probe.py "$OUT/plain$i" >> "$OUT/hashes.txt" 2>&1; done
python tools/dev/hash_probe.py "$OUT/torch1" torch >> "$OUT/hashes.txt" 2>&1
timeout 200 python tools/render_once.py --frames 2 2>&1 | tail -1 | cut -c1-330 >> "$OUT/hashes.txt"
timeout 200 python tools/dev/cull_diff.py . 96x64 2>&1 | cut -c1-260 >> "$OUT/hashes.txt"
timeout 300 python bench.py --steps 3 --warmup 3 --no-configs --no-parity --no-cpu-baseline --c5-frames 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['implementation']['march_cubin_hash'], d['roofline']['launch_ms'])" >> "$OUT/hashes.txt" 2>&1
cat "$OUT/cpu.txt" | cut -c1-200; cat "$OUT/hashes.txt"; ls -la "$OUT"
