#!/bin/bash
# Round 2, multi-GPU call (2 GPUs):  gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_call3_n2.sh'
set -u
OUT=gpurun_out/r2c3
mkdir -p "$OUT"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
timeout 300 python -m pytest tests/test_gpu_sharded.py -q -m gpu -rxXs -p no:cacheprovider > "$OUT/tests_sharded.log" 2>&1; echo "sharded tests rc=$?" >> "$OUT/rc.txt"
for ps in 0 1; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --peer-stores $ps > "$OUT/bench_n2_ps$ps.json" 2> "$OUT/bench_n2_ps$ps.err"; echo "bench n2 ps$ps rc=$?" >> "$OUT/rc.txt"
done
timeout 200 python tools/sharded_bench.py --devices 0,1 --mode 0 > "$OUT/sharded_peer.json" 2>&1; echo "sharded peer rc=$?" >> "$OUT/rc.txt"
timeout 200 python tools/sharded_bench.py --devices 0,1 --mode 1 > "$OUT/sharded_nccl.json" 2>&1; echo "sharded nccl rc=$?" >> "$OUT/rc.txt"
timeout 200 python tools/sharded_bench.py --devices 0,1 --mode 0 --size 8192x8192 --frames 6 > "$OUT/sharded_peer_c5.json" 2>&1; echo "sharded peer c5 rc=$?" >> "$OUT/rc.txt"
cat "$OUT/rc.txt"; tail -3 "$OUT/tests_sharded.log"; cat "$OUT"/sharded_*.json
