out=gpurun_out; mkdir -p $out
PB_REGION_TMA=1 timeout 300 python -m pytest tests/test_gpu_regions.py -m gpu -q -x > $out/r02tma_tests.log 2>&1; echo "pytest (TMA) rc=$?"
tail -3 $out/r02tma_tests.log
timeout 300 python profiles/scripts/region_tma_ab.py > $out/r02tma_ab.json 2> $out/r02tma_ab.err; echo "ab rc=$?"; cat $out/r02tma_ab.json; tail -2 $out/r02tma_ab.err
