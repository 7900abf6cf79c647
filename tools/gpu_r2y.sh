#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lp_ or dense_tiled or seeded" 2>&1 | tail -3
for KE in 80 32; do
RGCN_BLOCK_WGRAD_KE=$KE timeout 300 python bench.py --workload fb15k_block --steps 10 --warmup 3 --no-cpu-baseline > $O/r2y_bench_fb15k_block_ke$KE.json 2> $O/r2y.err
python tools/benchline.py < $O/r2y_bench_fb15k_block_ke$KE.json
done
