B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$B > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:knn_lane_kernel -s 2 -c 1 -f -o gpurun_out/r2_knn_final $B > /dev/null 2>&1
ls -la gpurun_out/r2_knn_final.ncu-rep
