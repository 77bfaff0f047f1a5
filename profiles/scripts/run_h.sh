python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py > gpurun_out/r02h_tests.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02h_tests.log | tail -20
python bench.py --workload peaks > gpurun_out/r02h_peaks.json 2> gpurun_out/r02h_peaks.err; echo peaks rc=$?
python bench.py --steps 10 --warmup 3 > gpurun_out/r02h_bench_c2.json 2> gpurun_out/r02h_bench_c2.err; echo c2 rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -f -k 'regex:pb_region_sums|pb_chain_items' -s 2 -c 4 -o gpurun_out/r02h_c2 python bench.py --steps 2 --warmup 1 > gpurun_out/r02h_ncu.log 2>&1; echo ncu rc=$?
