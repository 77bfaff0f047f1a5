out=gpurun_out; mkdir -p $out
for n in 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 3 > $out/r02n${n}_c2.json 2> $out/r02n${n}_c2.err; echo "n$n c2 rc=$?"
python -c "
import json; d=json.load(open('$out/r02n${n}_c2.json')); print('n$n', d['ms_per_step'], d.get('extended'), d['per_rank']['tiles_kernel_ms'], d['per_rank']['bins'], d['e2e']['ms_per_step'], d['e2e']['table_equals_device_resident_leg'])"
done
