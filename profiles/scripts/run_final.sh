# last state check of the round: whole GPU suite, smoke, default bench, reference arm
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -q > $out/r02d_gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -3 $out/r02d_gpu_tests.log
python __graft_entry__.py smoke > $out/r02d_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/r02d_smoke.log
python bench.py --steps 10 --warmup 3 > $out/r02d_bench.json 2> $out/r02d_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $out/r02d_bench_ref.json 2> $out/r02d_bench_ref.err; echo "reference arm rc=$?"
python bench.py --workload c3 --steps 10 --warmup 3 > $out/r02d_bench_c3.json 2> $out/r02d_bench_c3.err; echo "c3 rc=$?"
python - <<PY
import json
for n in ("bench", "bench_ref", "bench_c3"):
    d=json.load(open("$out/r02d_%s.json" % n)); print(n, d["ms_per_step"], d["value"], (d.get("e2e") or {}).get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("table_only") or {}).get("ms_per_step"))
PY
