#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lp_ or seeded" 2>&1 | tail -3
timeout 300 python bench.py --workload fb15k_block --steps 10 --warmup 3 --no-cpu-baseline > $O/r2y_bench_fb15k_block.json 2> $O/r2y.err
python tools/benchline.py < $O/r2y_bench_fb15k_block.json
RGCN_BLOCK_SQ=0 timeout 300 python bench.py --workload fb15k_block --steps 10 --warmup 3 --no-cpu-baseline 2> $O/r2y.err | python tools/benchline.py
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2y_launches_fb15k_block.csv python bench.py --workload fb15k_block --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/launch_summary.py $O/r2y_launches_fb15k_block.csv 2>/dev/null | sort -t= -k4 -r | head -6
