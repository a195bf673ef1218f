#!/bin/bash
# Round 2, 2-GPU call:  gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_call17_n2.sh'   (final defaults: TMA overlay, radius 1500)
set -u
OUT=gpurun_out/r2c25
mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_gpu_sharded.py -q -m gpu -rxXs -p no:cacheprovider > "$OUT/tests_sharded.log" 2>&1; echo "sharded tests rc=$?" >> "$OUT/rc.txt"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > "$OUT/bench_n2_peer.json" 2> "$OUT/bench_n2_peer.err"; echo "bench n2 rc=$?" >> "$OUT/rc.txt"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > "$OUT/bench_n2_reference.json" 2> "$OUT/bench_n2_reference.err"; echo "reference n2 rc=$?" >> "$OUT/rc.txt"
timeout 200 python tools/sharded_bench.py --devices 0,1 --mode 0 --size 8192x8192 --frames 6 > "$OUT/sharded_peer_c5.json" 2>&1; echo "sharded peer c5 rc=$?" >> "$OUT/rc.txt"
cat "$OUT/rc.txt"; tail -3 "$OUT/tests_sharded.log"; cat "$OUT"/sharded_*.json | cut -c1-600; python -c "
import json
d=json.loads(open('$OUT/bench_n2_peer.json').read().strip().splitlines()[-1])
print('n2', round(d['value'],1), round(d['ms_per_step'],3), 'march', round(d['roofline']['launch_ms'],3), 'c5', d['c5_strong'] and round(d['c5_strong']['ms_per_frame'],2), d['e2e'])
r=json.loads(open('$OUT/bench_n2_reference.json').read().strip().splitlines()[-1]); print('ref', r.get('value'), r.get('cpu_baseline',{}).get('cores'))"
