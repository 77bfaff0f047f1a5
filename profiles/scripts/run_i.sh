python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py --deselect tests/test_gpu_multirank.py > gpurun_out/r02i_tests.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02i_tests.log | tail -20
for wl in c2 c4 c3; do python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/r02i_bench_$wl.json 2> gpurun_out/r02i_bench_$wl.err; echo $wl rc=$?; done
timeout 300 ncu --set full --clock-control none --import-source on -f -k 'regex:pb_block_sums|pb_chain_items|pb_chain_totals' -s 3 -c 3 -o gpurun_out/r02i_c2 python bench.py --steps 2 --warmup 1 > gpurun_out/r02i_ncu.log 2>&1; echo ncu rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -f -k 'regex:pb_window_blocks|pb_window_fill' -s 2 -c 2 -o gpurun_out/r02i_c4 python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/r02i_ncu4.log 2>&1; echo ncu4 rc=$?
