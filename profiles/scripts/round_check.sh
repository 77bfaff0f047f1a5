#!/bin/bash
# One gpurun call that re-establishes the state of the hot path on a fresh B200 (about 6 GPU-minutes):
#   gpurun --timeout 900 -- 'bash profiles/scripts/round_check.sh r02'
# 1. pytest -m gpu   2. smoke()   3. bench.py (own arm, then the reference arm)   4. ncu launch list of the same
# bench command (share of the step per kernel)   5. one `ncu --set full` capture of the dominant kernel.
# Everything lands in gpurun_out/<tag>_*; turn the ncu artefacts into the committed text with profiles/summarize.py.
tag=${1:-check}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_gpu_tests.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_gpu_tests.log
python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "reference arm rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 > $out/${tag}_launches.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pb_point_tiles_kernel -s 2 -c 1 -f \
    -o $out/${tag}_point_tiles python bench.py --steps 2 --warmup 1 > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 $out/${tag}_gpu_tests.log; cat $out/${tag}_bench.json
