#!/bin/bash
# One gpurun call: fused-path tests, full GPU suite, A/B timing sweep, bench lines, ncu captures.  Logs -> gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
echo "== fused tests"; timeout 600 python -m pytest tests/test_gpu_fused.py -q -x 2>&1 | tail -40 | tee $O/test_fused.log
echo "== sweep"; timeout 400 python tools/fused_sweep.py 0 512 384 256 2>&1 | tail -12 | tee $O/sweep.log
RGCN_FUSE_ORDER=0 timeout 200 python tools/fused_sweep.py 512 2>&1 | tail -3 | tee $O/sweep_order0.log
echo "== full gpu suite"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee $O/test_all.log
echo "== bench"; RGCN_FUSED=1 timeout 400 python bench.py > $O/bench_fused.json 2> $O/bench_fused.err; tail -c 600 $O/bench_fused.json
echo "== ncu"
RGCN_FUSED=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_fused.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_b.log 2>&1
RGCN_FUSED=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused_rows -s 2 -c 2 -f -o $O/fused_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1
timeout 120 ncu -i $O/fused_full.ncu-rep --page raw --csv > $O/fused_full_raw.csv 2>/dev/null
timeout 120 ncu -i $O/fused_full.ncu-rep --page source --csv > $O/fused_full_source.csv 2>/dev/null
ls -la $O | tail -20
