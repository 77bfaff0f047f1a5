# region sums in one launch (the last block's warp finishes its chain): parity, then the C2 line's roofline_gather
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_regions.py tests/test_gpu_parity.py tests/test_gpu_scripts.py tests/test_gpu_fullsize.py tests/test_gpu_ref_goldens.py -m gpu -q > $out/r02g_tests.log 2>&1; echo "pytest rc=$?"
tail -3 $out/r02g_tests.log
python __graft_entry__.py smoke > $out/r02g_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 10 --warmup 3 > $out/r02g_bench.json 2> $out/r02g_bench.err; echo "bench rc=$?"
python bench.py --workload c1 --steps 10 --warmup 3 > $out/r02g_bench_c1.json 2> $out/r02g_bench_c1.err; echo "c1 rc=$?"
python - <<PY
import json
d=json.load(open("$out/r02g_bench.json")); print("c2", d["ms_per_step"], d["roofline_gather"]["kernel_ms"], d["roofline_gather"]["frac"], d["e2e"]["ms_per_step"], d["per_rank"]["region_sums_ms"])
d=json.load(open("$out/r02g_bench_c1.json")); print("c1", d["ms_per_step"], d.get("graph_replay_ms"), d["e2e"]["ms_per_step"])
PY
