# parity after the phase-multiplicity fix / multi-block region sums / sub-range overflow jobs, then the A/B runs
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -q > $out/r02t_tests.log 2>&1; echo "pytest rc=$?"
tail -4 $out/r02t_tests.log
# (gather_ab.py: A/B of 1-8 exon blocks per warp and 4-12 loads in flight per lane; both slower, code and script removed, result in profiles/gather_ab_r02.json)
python profiles/scripts/virtual_ranks.py --world 8 --cut cost > $out/r02t_vranks_cost.json 2> $out/r02t_vranks_cost.err; echo "vranks cost rc=$?"; cat $out/r02t_vranks_cost.json
python profiles/scripts/virtual_ranks.py --world 8 --cut reads > $out/r02t_vranks_reads.json 2> $out/r02t_vranks_reads.err; echo "vranks reads rc=$?"; cat $out/r02t_vranks_reads.json
python bench.py --steps 10 --warmup 3 > $out/r02t_bench_c2.json 2> $out/r02t_bench_c2.err; echo "bench rc=$?"
python bench.py --workload c4 --steps 10 --warmup 3 > $out/r02t_bench_c4.json 2> $out/r02t_bench_c4.err; echo "c4 rc=$?"
