# Center pile-up splitting: C3 with a chrM-like pile-up (10 % of the reads in the last 16.5 kb), split and unsplit
python bench.py --workload c3 --pileup 10000000 --steps 10 --warmup 3 > gpurun_out/r02q_c3_pile.json 2> gpurun_out/r02q_c3_pile.err; echo "c3 pile rc=$?"; tail -c 300 gpurun_out/r02q_c3_pile.err
PB_CENTER_SPLIT=0 python bench.py --workload c3 --pileup 10000000 --steps 5 --warmup 3 > gpurun_out/r02q_c3_pile_nosplit.json 2> gpurun_out/r02q_c3_pile_nosplit.err; echo "c3 pile nosplit rc=$?"; tail -c 300 gpurun_out/r02q_c3_pile_nosplit.err
