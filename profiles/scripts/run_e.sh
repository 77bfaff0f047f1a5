# end-to-end C3 (spliced batch, Center rule, chunked upload): lanes x chunks A/B
out=gpurun_out; mkdir -p $out
for lanes in 1 2; do for chunks in 4 8 12; do
  PB_CENTER_LANES=$lanes PB_UPLOAD_CHUNKS=$chunks python bench.py --workload c3 --steps 5 --warmup 3 > $out/r02e_c3_l${lanes}_c${chunks}.json 2> $out/r02e_c3_l${lanes}_c${chunks}.err; echo "rc=$?"
  python -c "
import json; d=json.load(open('$out/r02e_c3_l${lanes}_c${chunks}.json')); print('lanes $lanes chunks $chunks', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('table_equals_device_resident_leg'))"
done; done
