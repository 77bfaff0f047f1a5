# end-to-end C2 (the headline e2e): chunks x lanes A/B
out=gpurun_out; mkdir -p $out
for cfg in "2 6" "2 8" "2 12" "2 16" "2 24" "1 8" "1 16" "3 12"; do
  set -- $cfg; lanes=$1; chunks=$2
  PB_POINT_LANES=$lanes PB_UPLOAD_CHUNKS=$chunks python bench.py --steps 10 --warmup 3 > $out/r02e2_c2_l${lanes}_c${chunks}.json 2> $out/r02e2_c2_l${lanes}_c${chunks}.err; echo "rc=$?"
  python -c "
import json; d=json.load(open('$out/r02e2_c2_l${lanes}_c${chunks}.json')); print('lanes $lanes chunks $chunks', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('table_equals_device_resident_leg'), 'table only', d['e2e']['table_only']['ms_per_step'])"
done
