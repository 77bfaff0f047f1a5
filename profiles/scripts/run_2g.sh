# two GPUs: the programs under torchrun (gloo on one device, NCCL on two), then C4 with the matrix left in place vs all-reduced
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_multirank.py -m gpu -q > $out/r02g2_tests.log 2>&1; echo "pytest rc=$?"
tail -4 $out/r02g2_tests.log
for ex in slices matrix; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --workload c4 --c4-exchange $ex --steps 20 --warmup 3 > $out/r02g2_c4_$ex.json 2> $out/r02g2_c4_$ex.err; echo "c4 $ex rc=$?"
  python -c "
import json; d=json.load(open('$out/r02g2_c4_$ex.json')); print('$ex', d['ms_per_step'], d['profile_checksum'], d['regions_counted_max'])"
done
python bench.py --workload c4 --steps 20 --warmup 3 > $out/r02g2_c4_n1.json 2> $out/r02g2_c4_n1.err; python -c "
import json; d=json.load(open('$out/r02g2_c4_n1.json')); print('n1', d['ms_per_step'], d['profile_checksum'], d['regions_counted_max'])"
