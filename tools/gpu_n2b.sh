#!/bin/bash
# 2 GPUs: sharded parity (relation and row sharding), strong-scaling bench of both
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
N=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== parity relations"; timeout 600 $TR --master-port 29511 tests/sharded_check.py 2>&1 | grep -v "^W\|warn" | tail -6 | tee $O/r2i_sharded_rel_n$N.log
echo "== parity rows"; SHARD=rows timeout 600 $TR --master-port 29512 tests/sharded_check.py 2>&1 | grep -v "^W\|warn" | tail -6 | tee $O/r2i_sharded_rows_n$N.log
echo "== bench relations"; timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 > $O/r2i_bench_rel_n$N.json 2> $O/r2i_bench_rel_n$N.err; tail -c 1200 $O/r2i_bench_rel_n$N.json; tail -3 $O/r2i_bench_rel_n$N.err
echo "== bench rows"; timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 --shard rows > $O/r2i_bench_rows_n$N.json 2> $O/r2i_bench_rows_n$N.err; tail -c 1200 $O/r2i_bench_rows_n$N.json; tail -3 $O/r2i_bench_rows_n$N.err
