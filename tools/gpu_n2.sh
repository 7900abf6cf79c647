#!/bin/bash
# 2-GPU sanity: sharded layer parity vs the single-GPU layer, then the bench line at N=2 (both arms)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== sharded check"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/sharded_check.py 2>&1 | tail -8 | tee $O/sharded_n2.log
echo "== bench n2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; python tools/benchline.py < $O/bench_n2.json; tail -3 $O/bench_n2.err
echo "== reference arm n2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; tail -c 400 $O/bench_ref_n2.json; echo; tail -2 $O/bench_ref_n2.err
