#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py "tests/test_gpu_headline_parity.py::test_syn_width_ten_million_edges" -q -x 2>&1 | tail -5
echo "== syn"; timeout 900 python bench.py --workload syn --steps 3 --warmup 3 --no-cpu-baseline > $O/r2p_bench_syn.json 2> $O/r2p_bench_syn.err; tail -3 $O/r2p_bench_syn.err | cut -c1-300; python tools/benchline.py < $O/r2p_bench_syn.json
echo "== syn two-phase/tiled"; RGCN_FUSED=0 timeout 900 python bench.py --workload syn --steps 3 --warmup 3 --no-cpu-baseline > $O/r2p_bench_syn_tiled.json 2>/dev/null; python tools/benchline.py < $O/r2p_bench_syn_tiled.json
echo "== syn fused both"; RGCN_FUSED=2 timeout 900 python bench.py --workload syn --steps 3 --warmup 3 --no-cpu-baseline > $O/r2p_bench_syn_fused2.json 2>/dev/null; python tools/benchline.py < $O/r2p_bench_syn_fused2.json
