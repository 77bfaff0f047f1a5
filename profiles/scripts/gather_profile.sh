#!/bin/bash
# ncu --set full captures of the region / window / plane-free kernels: one bench run per workload.
#   gpurun --timeout 900 -- 'bash profiles/scripts/gather_profile.sh r02g'
tag=${1:-gather}
out=gpurun_out
mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k 'regex:pb_region_sums|pb_chain_items|pb_chain_slices|pb_read_index|pb_chain_finish' -s 10 -c 6 -o $out/${tag}_c2 python bench.py --steps 2 --warmup 1 > $out/${tag}_c2.log 2>&1; echo "c2 rc=$?"
timeout 300 $NCU -k 'regex:pb_gather_windows' -s 2 -c 1 -o $out/${tag}_c4 python bench.py --workload c4 --steps 2 --warmup 1 > $out/${tag}_c4.log 2>&1; echo "c4 rc=$?"
ls -la $out | tail -5
