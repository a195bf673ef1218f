#!/bin/bash
# Round 2, 2-GPU call (final defaults): the N = 2 bench line only
set -u
OUT=gpurun_out/r2c31
mkdir -p "$OUT"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > "$OUT/bench_n2_peer.json" 2> "$OUT/bench_n2_peer.err"; echo "bench n2 rc=$?"
python -c "
import json
d=json.loads(open('$OUT/bench_n2_peer.json').read().strip().splitlines()[-1])
print('n2', round(d['value'],1), round(d['ms_per_step'],3), 'march', round(d['roofline']['launch_ms'],3), 'c5', d['c5_strong'] and round(d['c5_strong']['ms_per_frame'],2), round(d['e2e']['value'],1))"
