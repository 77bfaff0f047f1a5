out=gpurun_out; mkdir -p $out
for cfg in "2048 2048" "4096 4096" "8192 8192" "8192 2048" "16384 2048" "4096 1024"; do
  set -- $cfg; f=$1; i=$2
  PB_FIRST_READS=$f PB_ITEM_READS=$i python bench.py --steps 10 --warmup 3 > $out/r02f2_c2_f${f}_i$i.json 2> $out/r02f2_c2_f${f}_i$i.err; echo "rc=$?"
  python -c "
import json; d=json.load(open('$out/r02f2_c2_f${f}_i$i.json')); t=d['table_only']; print('first $f item $i', t['ms_per_step'], t['identical_to_plane_path'])"
done
