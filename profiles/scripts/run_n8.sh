# strong scaling of ONE C2 batch on 8 GPUs with cost-balanced position cuts (gpurun --gpus 8)
out=gpurun_out; mkdir -p $out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > $out/r02x_n8_c2.json 2> $out/r02x_n8_c2.err; echo "n8 c2 rc=$?"
head -c 1500 $out/r02x_n8_c2.json
