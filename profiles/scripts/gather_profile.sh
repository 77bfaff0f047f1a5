#!/bin/bash
# ncu --set full captures of the gather / profile kernels: one bench run per workload.
#   gpurun --timeout 900 -- 'bash profiles/scripts/gather_profile.sh r02a'
tag=${1:-gather}
out=gpurun_out
mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:pb_region_sums -s 2 -c 1 -o $out/${tag}_region_sums python bench.py --steps 2 --warmup 1 > $out/${tag}_c2.log 2>&1; echo "c2 rc=$?"
timeout 300 $NCU -k 'regex:pb_gather_windows|pb_window_normalize|pb_column_keys|pb_column_stats' -s 8 -c 4 -o $out/${tag}_c4 python bench.py --workload c4 --steps 2 --warmup 1 > $out/${tag}_c4.log 2>&1; echo "c4 rc=$?"
timeout 300 $NCU -k 'regex:pb_stratified_windows|pb_norm_keys|pb_column_stats' -s 6 -c 3 -o $out/${tag}_c2p python bench.py --workload c2p --steps 2 --warmup 1 > $out/${tag}_c2p.log 2>&1; echo "c2p rc=$?"
ls -la $out | tail -8
