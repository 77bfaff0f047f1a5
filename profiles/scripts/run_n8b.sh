# 8 GPUs: C4 (metagene pass) with the matrix left in place vs all-reduced, then the C2 default once more (owned-block region sums)
out=gpurun_out; mkdir -p $out
for ex in slices matrix; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --workload c4 --c4-exchange $ex --steps 20 --warmup 3 > $out/r02n8b_c4_$ex.json 2> $out/r02n8b_c4_$ex.err; echo "c4 $ex rc=$?"
  python -c "
import json; d=json.load(open('$out/r02n8b_c4_$ex.json')); print('$ex', d['ms_per_step'], d['profile_checksum'], d['regions_counted_max'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 3 > $out/r02n8b_c2.json 2> $out/r02n8b_c2.err; echo "c2 rc=$?"
python -c "
import json; d=json.load(open('$out/r02n8b_c2.json')); print('c2', d['ms_per_step'], d.get('extended'), d['per_rank']['region_sums_ms'], d['e2e']['ms_per_step'])"
