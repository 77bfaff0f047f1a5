# binning kernel rewrite (RED count pass, carry-over dense list, plain fill atomics): parity, then A/B on C3 and C5-like spliced point rule
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q > $out/r02v_tests.log 2>&1; echo "pytest rc=$?"
tail -4 $out/r02v_tests.log
for v in "OCC=6 AGG=0" "OCC=8 AGG=0" "OCC=6 AGG=1" "OCC=8 AGG=1"; do
  set -- $v; o=${1#OCC=}; a=${2#AGG=}
  PB_BIN_OCC=$o PB_BIN_AGG=$a python bench.py --workload c3 --steps 10 --warmup 3 > $out/r02v_c3_occ${o}_agg${a}.json 2> $out/r02v_c3_occ${o}_agg${a}.err; echo "c3 occ$o agg$a rc=$?"
  python - <<PY
import json
d=json.load(open("$out/r02v_c3_occ${o}_agg${a}.json")); print("occ$o agg$a", d["ms_per_step"], d["per_rank"]["map_ms"], d["per_rank"]["tiles_kernel_ms"], d["e2e"]["ms_per_step"])
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pb_ -c 80 --csv --log-file $out/r02v_launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 > $out/r02v_launches_c3.log 2>&1; echo "launch list rc=$?"
