#!/bin/bash
# Round 2, 8-GPU call:  gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_call20_n8.sh'   (final defaults)
set -u
OUT=gpurun_out/r2c30
mkdir -p "$OUT"
run() { # n tag
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $1 --steps 10 --warmup 3 > "$OUT/bench_n$1_$2.json" 2> "$OUT/bench_n$1_$2.err"; echo "bench n$1 $2 rc=$?" >> "$OUT/rc.txt"
}
run 8 peer
run 4 peer
timeout 200 python tools/sharded_bench.py --devices 0,1,2,3,4,5,6,7 --mode 0 --size 8192x8192 --frames 8 --check 1 > "$OUT/sharded_peer_n8_c5.json" 2>&1; echo "sharded peer n8 c5 rc=$?" >> "$OUT/rc.txt"
cat "$OUT/rc.txt"; cat "$OUT"/sharded_*.json | cut -c1-700; for f in "$OUT"/bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', round(d['value'],1), round(d['ms_per_step'],3), 'march', round(d['roofline']['launch_ms'],3), 'c5', d['c5_strong'] and round(d['c5_strong']['ms_per_frame'],2), 'e2e', round(d['e2e']['value'],1))"; done
