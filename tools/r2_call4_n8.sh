#!/bin/bash
# Round 2, 8-GPU call:  gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_call4_n8.sh'
set -u
OUT=gpurun_out/r2c4
mkdir -p "$OUT"
run() { # n peer-stores tag
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $1 --steps 10 --warmup 3 --peer-stores $2 > "$OUT/bench_n$1_$3.json" 2> "$OUT/bench_n$1_$3.err"; echo "bench n$1 $3 rc=$?" >> "$OUT/rc.txt"
}
run 8 1 peer
run 8 0 gather
run 4 1 peer
timeout 200 python tools/sharded_bench.py --devices 0,1,2,3,4,5,6,7 --mode 0 --check 0 > "$OUT/sharded_peer_n8.json" 2>&1; echo "sharded peer n8 rc=$?" >> "$OUT/rc.txt"
timeout 200 python tools/sharded_bench.py --devices 0,1,2,3,4,5,6,7 --mode 0 --size 8192x8192 --frames 8 --check 1 > "$OUT/sharded_peer_n8_c5.json" 2>&1; echo "sharded peer n8 c5 rc=$?" >> "$OUT/rc.txt"
cat "$OUT/rc.txt"; cat "$OUT"/sharded_*.json; for f in "$OUT"/bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', round(d['value'],1), round(d['ms_per_step'],3), 'march', round(d['roofline']['launch_ms'],3), 'c5', d['c5_strong'] and round(d['c5_strong']['ms_per_frame'],2))"; done
