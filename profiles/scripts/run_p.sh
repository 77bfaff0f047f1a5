# psite / phase passes with the site table in the stratified kernel: parity, then the c2p numbers (generic path for A/B)
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_scripts.py tests/test_gpu_ref_goldens.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q > $out/r02p_tests.log 2>&1; echo "pytest rc=$?"
tail -4 $out/r02p_tests.log
PB_STRAT_GENERIC=1 python bench.py --workload c2p --steps 10 --warmup 3 > $out/r02p_c2p_generic.json 2> $out/r02p_c2p_generic.err; echo "generic rc=$?"
python bench.py --workload c2p --steps 10 --warmup 3 > $out/r02p_c2p.json 2> $out/r02p_c2p.err; echo "c2p rc=$?"
python - <<PY
import json
for n in ("generic", ""):
    d=json.load(open("$out/r02p_c2p%s.json" % ("_"+n if n else ""))); print(n or "site table", d["ms_per_step"], d["stratified_kernel_ms"], d["profile_checksum"])
PY
timeout 300 ncu --set full --clock-control none --import-source on -f -k 'regex:pb_stratified_windows_kernel|pb_norm_keys|pb_column_stats' -s 3 -c 3 -o $out/r02p_c2p python bench.py --workload c2p --steps 2 --warmup 1 > $out/r02p_ncu.log 2>&1; echo "ncu rc=$?"
