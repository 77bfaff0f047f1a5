# plane-free counts: how many reads of a block its slice-finding warp counts itself (the rest go to the persistent item kernel)
out=gpurun_out; mkdir -p $out
for f in 2048 1024 512 256 0; do
  PB_FIRST_READS=$f python bench.py --steps 10 --warmup 3 > $out/r02f_c2_first$f.json 2> $out/r02f_c2_first$f.err; echo "rc=$?"
  python -c "
import json; d=json.load(open('$out/r02f_c2_first$f.json')); t=d['table_only']; print('first $f', t['ms_per_step'], t['identical_to_plane_path'])"
done
PB_FIRST_READS=512 python bench.py --workload c5 --steps 10 --warmup 3 > $out/r02f_c5_first512.json 2> $out/r02f_c5_first512.err
python -c "
import json; d=json.load(open('$out/r02f_c5_first512.json')); t=d['table_only']; print('c5 first 512', t['ms_per_step'], t['identical_to_plane_path'])"
