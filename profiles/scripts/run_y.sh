# owned-block sub-table for region sums on position-sharded ranks: parity of every ranged path, then the virtual ranks again
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_regions.py tests/test_gpu_parity.py tests/test_gpu_multirank.py tests/test_gpu_scripts.py -m gpu -q > $out/r02y_tests.log 2>&1; echo "pytest rc=$?"
tail -4 $out/r02y_tests.log
python profiles/scripts/virtual_ranks.py --world 8 --cut cost > $out/r02y_vranks_cost.json 2> $out/r02y_vranks_cost.err; echo "vranks rc=$?"; cat $out/r02y_vranks_cost.json
