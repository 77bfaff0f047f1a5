# site-table counting, quad-level slow paths + compile-time mask variant: parity (whole GPU suite) and numbers
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -q > $out/r02z2_tests.log 2>&1; echo "pytest rc=$?"
tail -4 $out/r02z2_tests.log
for wl in c2 c5 c1; do
  python bench.py --workload $wl --steps 10 --warmup 3 > $out/r02z2_bench_$wl.json 2> $out/r02z2_bench_$wl.err; echo "$wl rc=$?"
  python - <<PY
import json
d=json.load(open("$out/r02z2_bench_$wl.json")); t=d.get("table_only") or {}
print("$wl", d["ms_per_step"], "table_only", t.get("ms_per_step"), t.get("identical_to_plane_path"), (t.get("roofline") or {}).get("frac"), "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("table_only",{}).get("ms_per_step"))
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -f -k 'regex:pb_chain_first_items|pb_chain_items|pb_read_index' -s 3 -c 3 -o $out/r02z2_chain_counts python bench.py --steps 2 --warmup 1 > $out/r02z2_ncu.log 2>&1; echo "ncu rc=$?"
