#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
N=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== parity rows (symm)"; SHARD=rows timeout 600 $TR --master-port 29512 tests/sharded_check.py 2>&1 | grep -v "^W\|warn\|^\*\|OMP_NUM" | tail -6
echo "== bench rows"; timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 > $O/r2l_bench_rows_n$N.json 2> $O/r2l_bench_rows_n$N.err; tail -4 $O/r2l_bench_rows_n$N.err | cut -c1-300; python tools/benchline.py < $O/r2l_bench_rows_n$N.json; grep -o '"sharded_parity": "[^"]*"' $O/r2l_bench_rows_n$N.json; grep -o '"e2e": {[^}]*}' $O/r2l_bench_rows_n$N.json
