set -x
python -m pytest tests -m gpu -q -k "multi or second_device or sharded or device_arrays or in_process" 2>&1 | tail -5 > gpurun_out/r2_multi_tests.log
for N in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_strong_n$N.json 2> gpurun_out/r2_strong_n$N.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 --config 4 > gpurun_out/r2_config4_n8.json 2> gpurun_out/r2_config4_n8.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_strong_n1.json 2> gpurun_out/r2_strong_n1.err
tail -3 gpurun_out/r2_multi_tests.log
