for k in 32 64; do
ncu --set full --clock-control none --import-source on -k regex:knn_lane_kernel -s 4 -c 1 -f -o gpurun_out/r2_heap_k$k python scripts/kernel_ab.py --queries 10000000 -k $k --steps 1 > /dev/null 2>&1
done
ls -la gpurun_out/r2_heap_k*.ncu-rep
