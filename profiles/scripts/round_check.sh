#!/bin/bash
# One gpurun call that re-establishes the state of the hot path on a fresh B200 (about 8 GPU-minutes):
#   gpurun --timeout 1800 -- 'bash profiles/scripts/round_check.sh r02'
# 1. pytest -m gpu   2. smoke()   3. bench.py (own arm, then the reference arm)   4. ncu launch list of the same
# bench command (share of the step per kernel)   5. `ncu --set full` captures of the tiles kernel and of the region /
# window / plane-free kernels.  Everything lands in gpurun_out/<tag>_*; profiles/summarize.py turns the ncu artefacts
# into the committed text.
tag=${1:-check}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q > $out/${tag}_gpu_tests.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_gpu_tests.log
python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "reference arm rc=$?"
for wl in c1 c3 c4 c5 c2p; do python bench.py --workload $wl --steps 10 --warmup 3 > $out/${tag}_bench_$wl.json 2> $out/${tag}_bench_$wl.err; echo "$wl rc=$?"; done
python bench.py --workload peaks > $out/${tag}_peaks.json 2> $out/${tag}_peaks.err; echo "peaks rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 > $out/${tag}_launches.log 2>&1; echo "launch list rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:pb_point_tiles_kernel -s 2 -c 1 -o $out/${tag}_point_tiles python bench.py --steps 2 --warmup 1 > $out/${tag}_ncu_tiles.log 2>&1; echo "ncu tiles rc=$?"
timeout 400 $NCU -k 'regex:pb_block_sums|pb_chain_totals|pb_read_index|pb_chain_first_items|pb_chain_items' -s 8 -c 6 -o $out/${tag}_regions python bench.py --steps 2 --warmup 1 > $out/${tag}_ncu_regions.log 2>&1; echo "ncu regions rc=$?"
timeout 300 $NCU -k 'regex:pb_window_blocks|pb_window_fill|pb_window_normalize|pb_column_keys|pb_column_stats' -s 10 -c 5 -o $out/${tag}_c4 python bench.py --workload c4 --steps 2 --warmup 1 > $out/${tag}_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
tail -3 $out/${tag}_gpu_tests.log; head -c 400 $out/${tag}_bench.json
