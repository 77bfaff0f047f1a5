#!/bin/bash
# One gpurun call: quick GPU tests, then one bench line per workload (N = 1).
#   gpurun --timeout 1800 -- 'bash profiles/scripts/bench_all.sh r02f'
tag=${1:-all}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py > $out/${tag}_tests.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" $out/${tag}_tests.log | tail -20
for wl in c2 c1 c4 c2p c5 c3; do
  python bench.py --workload $wl --steps 10 --warmup 3 > $out/${tag}_bench_$wl.json 2> $out/${tag}_bench_$wl.err; echo "$wl rc=$?"
  tail -2 $out/${tag}_bench_$wl.err
done
