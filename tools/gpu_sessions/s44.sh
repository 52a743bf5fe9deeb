mkdir -p gpurun_out/s44
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 300 --warmup 100 --no-cpu-baseline > gpurun_out/s44/bench_cheetah_n2.json 2> gpurun_out/s44/n2.err; grep '^{' gpurun_out/s44/bench_cheetah_n2.json | tail -1 | cut -c1-260
