#!/bin/bash
# full verification: GPU suite, smoke, default bench line, reference arm
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/verify_pytest.log; tail -5 $O/verify_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py > $O/verify_bench.json 2> $O/verify_bench.err; tail -2 $O/verify_bench.err | cut -c1-300
python tools/benchline.py < $O/verify_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/verify_bench_ref.json 2> $O/verify_bench_ref.err; cut -c1-400 $O/verify_bench_ref.json
