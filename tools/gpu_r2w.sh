#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core_gemm or dense_tiled or seeded" 2>&1 | tail -4
timeout 300 python tools/umma_vs_fma.py 10000000 2>&1 | tail -2 | tee $O/r2w_umma_vs_fma.json
timeout 900 python bench.py --workload syn_none --steps 3 --warmup 3 > $O/r2w_bench_syn_none.json 2> $O/r2w_syn_none.err; tail -3 $O/r2w_syn_none.err | cut -c1-300
python tools/benchline.py < $O/r2w_bench_syn_none.json
