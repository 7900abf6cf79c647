#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
RGCN_UMMA_ORDER=l timeout 900 python bench.py --workload syn_none --steps 3 --warmup 3 --no-cpu-baseline > $O/r2w_bench_syn_none_lin.json 2> $O/r2w_syn_none.err; tail -3 $O/r2w_syn_none.err | cut -c1-300
python tools/benchline.py < $O/r2w_bench_syn_none_lin.json
