out=gpurun_out; mkdir -p $out
for cfg in "4 6" "8 10" "12 14" "2 4"; do
  set -- $cfg
  PB_STAGE_THREADS=$1 PB_STAGE_RING=$2 python bench.py --steps 5 --warmup 3 > $out/r02st_t$1.json 2> $out/r02st_t$1.err
  python -c "
import json; d=json.load(open('$out/r02st_t$1.json')); print('threads $1 ring $2', d['e2e']['soa_input']['ms_per_step'], d['e2e']['soa_input']['table_equals_packed_path'])"
done
