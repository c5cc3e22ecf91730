set -x
timeout 300 python scripts/kernel_ab.py > gpurun_out/r2_sweep_ab.json || exit 1
NBK_ORDER=passes timeout 300 python scripts/kernel_ab.py >> gpurun_out/r2_sweep_ab.json
cat gpurun_out/r2_sweep_ab.json
python -m pytest tests -m gpu -x -q -k "not config5 and not config4 and not config3" 2>&1 | tail -3
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r2_prof_bench.json 2> gpurun_out/r2_prof_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv $B > /dev/null 2> gpurun_out/r2_ncu_launch.err
ncu --set full --clock-control none --import-source on -k regex:knn_lane_kernel -s 2 -c 2 -f -o gpurun_out/r2_knn_lane_1e8 $B > /dev/null 2> gpurun_out/r2_ncu_full.err
K="python scripts/kernel_ab.py --queries 12500000 --steps 1"
$K > gpurun_out/r2_prof_lowdensity.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_lowdensity.csv $K > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_lane_kernel -s 4 -c 1 -f -o gpurun_out/r2_knn_lane_12m $K > /dev/null 2>&1
NBK_LIBRARY=nbodyhpc_b200/lib/variants/libnbk_pfleaf.so python scripts/kernel_ab.py > gpurun_out/r2_pfleaf.json
cat gpurun_out/r2_pfleaf.json gpurun_out/r2_prof_lowdensity.json
ls -la gpurun_out/*.ncu-rep
