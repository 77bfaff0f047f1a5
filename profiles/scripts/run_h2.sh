out=gpurun_out; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:pb_bin_kernel -s 4 -c 2 -o $out/r02h2_bin python bench.py --workload c3 --steps 2 --warmup 1 > $out/r02h2_ncu_bin.log 2>&1; echo "ncu bin rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:pb_center_tiles_kernel -s 2 -c 1 -o $out/r02h2_center python bench.py --workload c3 --steps 2 --warmup 1 > $out/r02h2_ncu_center.log 2>&1; echo "ncu center rc=$?"
