#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== lp step"; timeout 900 python bench.py --workload wn18_lp_step --steps 20 --warmup 3 > $O/r2r_bench_wn18_lp_step.json 2> $O/r2r_bench_wn18_lp_step.err; tail -5 $O/r2r_bench_wn18_lp_step.err | cut -c1-300; cat $O/r2r_bench_wn18_lp_step.json | cut -c1-1500
echo "== sampling tests"; timeout 600 python -m pytest tests/test_gpu_sampling.py -q -x 2>&1 | tail -3
