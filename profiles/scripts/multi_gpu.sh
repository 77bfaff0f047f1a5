#!/bin/bash
# One gpurun --gpus N call: the multi-rank program tests over NCCL (N >= 2), then position-sharded bench lines.
#   gpurun --gpus 2 --timeout 1500 -- 'bash profiles/scripts/multi_gpu.sh r02 2 "c2 c3 c4 c5"'
tag=${1:-multi}; n=${2:-2}; wls=${3:-c2}
out=gpurun_out
mkdir -p $out
nvidia-smi topo -m > $out/${tag}_n${n}_topo.txt 2>&1
if [ "$n" = "2" ]; then
  python -m pytest tests/test_gpu_multirank.py -q > $out/${tag}_n${n}_multirank.log 2>&1
  grep -E "^(FAILED|ERROR)|passed|failed|skipped" $out/${tag}_n${n}_multirank.log | tail -5
fi
port=29700
for wl in $wls; do
  extra=""
  if [ "$wl" = "c5" ]; then extra="--sharding chromosomes"; fi
  port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --workload $wl --steps 20 --warmup 3 $extra > $out/${tag}_n${n}_$wl.json 2> $out/${tag}_n${n}_$wl.err
  echo "$wl n=$n rc=$?"; tail -2 $out/${tag}_n${n}_$wl.err | cut -c1-300
  head -c 600 $out/${tag}_n${n}_$wl.json; echo
done
port=$((port+1))
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    profiles/scripts/h2d_ranks.py > $out/${tag}_n${n}_h2d.json 2> $out/${tag}_n${n}_h2d.err
echo "h2d n=$n rc=$?"; cat $out/${tag}_n${n}_h2d.json
