# spliced Center batches mapped range by range on two streams: parity at full size, then A/B of the number of ranges on C3
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -k "center or Center or sharding or spliced" > $out/r02w_tests.log 2>&1; echo "pytest rc=$?"
tail -4 $out/r02w_tests.log
for k in 1 2 4 8; do
  PB_MAP_RANGES=$k python bench.py --workload c3 --steps 10 --warmup 3 > $out/r02w_c3_ranges$k.json 2> $out/r02w_c3_ranges$k.err; echo "c3 ranges=$k rc=$?"
  python - <<PY
import json
d=json.load(open("$out/r02w_c3_ranges$k.json")); print("ranges $k", d["ms_per_step"], d["per_rank"]["map_ms"], d["per_rank"]["tiles_kernel_ms"], d["roofline"]["frac"], d["e2e"]["ms_per_step"])
PY
done
