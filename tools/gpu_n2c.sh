#!/bin/bash
# N GPUs (NG, default 2): row-sharded parity + strong-scaling bench (rows), relation sharding for comparison
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
N=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== parity rows"; SHARD=rows timeout 600 $TR --master-port 29512 tests/sharded_check.py 2>&1 | grep -v "^W\|warn\|^\*\|OMP_NUM" | tail -8 | tee $O/r2j_sharded_rows_n$N.log
echo "== bench rows"; timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 --shard rows > $O/r2j_bench_rows_n$N.json 2> $O/r2j_bench_rows_n$N.err; tail -3 $O/r2j_bench_rows_n$N.err | cut -c1-300
python tools/benchline.py < $O/r2j_bench_rows_n$N.json
for c in 2 8; do RGCN_SHARD_CHUNKS=$c timeout 600 $TR --master-port 2952$c bench.py --gpus $N --steps 20 --warmup 3 --shard rows > $O/r2j_bench_rows_n${N}_c$c.json 2>/dev/null; python tools/benchline.py < $O/r2j_bench_rows_n${N}_c$c.json; done
