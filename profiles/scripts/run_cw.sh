# end-to-end C2: relative sizes of the eight upload chunks
out=gpurun_out; mkdir -p $out
i=0
for w in "1,1,1,1,1,1,1,1" "0.5,1.1,1.2,1.2,1.2,1.2,1.0,0.6" "0.6,1.2,1.2,1.2,1.2,1.2,0.9,0.5" "0.4,1.0,1.3,1.3,1.3,1.3,0.9,0.5" "1,1.1,1.1,1.1,1.1,1.1,1.0,0.5" "0.5,1,1.2,1.3,1.3,1.3,0.9,0.5,0.3"; do
  i=$((i+1))
  n=$(echo $w | tr ',' '\n' | wc -l)
  PB_CHUNK_WEIGHTS=$w PB_UPLOAD_CHUNKS=$n python bench.py --steps 10 --warmup 3 > $out/r02cw_$i.json 2> $out/r02cw_$i.err; echo "rc=$?"
  python -c "
import json; d=json.load(open('$out/r02cw_$i.json')); print('weights $w', 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('table_equals_device_resident_leg'))"
done
