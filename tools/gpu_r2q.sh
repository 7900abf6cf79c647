#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "graphed or train_mode or block_diag or negative_sampling" 2>&1 | tail -8
echo "== wn18 graphed"; timeout 600 python bench.py --workload wn18 --steps 50 --warmup 5 > $O/r2q_bench_wn18_graph.json 2> $O/r2q_bench_wn18_graph.err; tail -3 $O/r2q_bench_wn18_graph.err | cut -c1-300; python tools/benchline.py < $O/r2q_bench_wn18_graph.json
echo "== wn18 eager"; RGCN_LP_GRAPH=0 timeout 600 python bench.py --workload wn18 --steps 50 --warmup 5 > $O/r2q_bench_wn18_eager.json 2>/dev/null; python tools/benchline.py < $O/r2q_bench_wn18_eager.json
