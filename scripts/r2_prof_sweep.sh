ncu --set full --clock-control none -k regex:sweep_kernel -s 3 -c 3 -f -o gpurun_out/r2_sweep python scripts/kernel_ab.py --queries 100000000 --steps 1 > /dev/null 2>&1
ls -la gpurun_out/r2_sweep.ncu-rep
