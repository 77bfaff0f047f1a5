# re-run of the three tests that failed in run_t (tolerances), multirank programs, then ncu of the C3 binning passes
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_goldens.py tests/test_gpu_multirank.py -m gpu -q > $out/r02u_tests.log 2>&1; echo "pytest rc=$?"
tail -4 $out/r02u_tests.log
python bench.py --steps 10 --warmup 3 > $out/r02u_bench_c2.json 2> $out/r02u_bench_c2.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:pb_bin_kernel -s 4 -c 2 -o $out/r02u_bin python bench.py --workload c3 --steps 2 --warmup 1 > $out/r02u_ncu_bin.log 2>&1; echo "ncu bin rc=$?"
