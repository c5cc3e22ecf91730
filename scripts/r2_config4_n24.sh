for N in 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 5 --warmup 3 --config 4 > gpurun_out/r2_config4_n$N.json 2> gpurun_out/r2_config4_n$N.err
done
