#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== ncu sampler"; timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_sample_edge -c 1 -o $O/sampler_full -f python bench.py --workload wn18_sampling --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_sampler.log 2>&1; tail -3 $O/ncu_sampler.log
echo "== ncu rank count"; timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_rank_count -c 1 -o $O/rank_full -f python bench.py --workload wn18_ranking --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_rank_full.log 2>&1; tail -3 $O/ncu_rank_full.log
ls -la $O/*.ncu-rep
