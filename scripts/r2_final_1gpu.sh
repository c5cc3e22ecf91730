python -m pytest tests -m gpu -x -q --durations=5 -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_tests_final.log
tail -3 gpurun_out/r2_tests_final.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 300 gpurun_out/r2_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/r2_config4_final.json 2> gpurun_out/r2_config4_final.err; tail -c 300 gpurun_out/r2_config4_final.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
