#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lp_ or dense_tiled or seeded" 2>&1 | tail -6
timeout 300 python bench.py --workload fb15k_block --steps 10 --warmup 3 --no-cpu-baseline > $O/r2y_bench_fb15k_block.json 2> $O/r2y.err; tail -2 $O/r2y.err | cut -c1-300
python tools/benchline.py < $O/r2y_bench_fb15k_block.json
RGCN_SPLIT_SELF=0 timeout 300 python bench.py --workload fb15k_block --steps 5 --warmup 3 --no-cpu-baseline > $O/r2y_bench_fb15k_block_generic.json 2> $O/r2y.err; tail -2 $O/r2y.err | cut -c1-300
python tools/benchline.py < $O/r2y_bench_fb15k_block_generic.json
