#!/bin/bash
# N GPUs (NG, default 2): row-sharded layer with the symmetric-memory exchange vs NCCL all-gathers
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
N=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== parity rows (symm)"; SHARD=rows timeout 600 $TR --master-port 29512 tests/sharded_check.py 2>&1 | grep -v "^W\|warn\|^\*\|OMP_NUM" | tail -12 | tee $O/r2k_sharded_rows_symm_n$N.log
echo "== bench rows symm"; timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 --shard rows > $O/r2k_bench_rows_symm_n$N.json 2> $O/r2k_bench_rows_symm_n$N.err; tail -4 $O/r2k_bench_rows_symm_n$N.err | cut -c1-300; python tools/benchline.py < $O/r2k_bench_rows_symm_n$N.json
echo "== bench rows nccl"; RGCN_SHARD_COMM=nccl timeout 600 $TR --master-port 29515 bench.py --gpus $N --steps 20 --warmup 3 --shard rows > $O/r2k_bench_rows_nccl_n$N.json 2>/dev/null; python tools/benchline.py < $O/r2k_bench_rows_nccl_n$N.json
