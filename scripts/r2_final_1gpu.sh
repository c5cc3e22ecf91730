python -m pytest tests -m gpu -x -q --durations=5 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r2_tests_final.log; echo "pytest rc=$?" >> gpurun_out/r2_tests_final.log
tail -4 gpurun_out/r2_tests_final.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_final.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
