#!/bin/bash
# quick N=1 check: GPU tests without the full-size / multi-rank files, then C2 and C4 bench lines
tag=${1:-quick}
python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py --deselect tests/test_gpu_multirank.py > gpurun_out/${tag}_tests.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${tag}_tests.log | tail -20
for wl in c2 c4 c5; do python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/${tag}_bench_$wl.json 2> gpurun_out/${tag}_bench_$wl.err; echo $wl rc=$?; done
