# full GPU suite + the default bench on a fresh box (state check after the indexed-access / quirk commits)
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -q -x > $out/r02s_tests.log 2>&1; echo "pytest rc=$?"
tail -5 $out/r02s_tests.log
python bench.py --steps 10 --warmup 3 > $out/r02s_bench_c2.json 2> $out/r02s_bench_c2.err; echo "bench rc=$?"
head -c 600 $out/r02s_bench_c2.json
