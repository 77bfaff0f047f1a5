for split in 32768 8192 4096; do
  for wl in c2 c5; do
    PB_POINT_SPLIT=$split python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/r02o_${wl}_$split.json 2> gpurun_out/r02o_${wl}_$split.err; echo "$wl split=$split rc=$?"
  done
done
python -m pytest tests/test_gpu_regions.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
