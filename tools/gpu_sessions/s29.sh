mkdir -p gpurun_out/s29
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 300 --warmup 100 > gpurun_out/s29/bench_cheetah_n$N.log 2> gpurun_out/s29/bench_cheetah_n$N.err
grep '^{' gpurun_out/s29/bench_cheetah_n$N.log | tail -1 | cut -c1-260
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/s29/bench_cheetah_n${N}_driver.log 2> gpurun_out/s29/bench_cheetah_n${N}_driver.err
grep '^{' gpurun_out/s29/bench_cheetah_n${N}_driver.log | tail -1 | cut -c1-260
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --config humanoid --steps 200 --warmup 100 --no-cpu-baseline > gpurun_out/s29/bench_humanoid_n$N.log 2> gpurun_out/s29/bench_humanoid_n$N.err
grep '^{' gpurun_out/s29/bench_humanoid_n$N.log | tail -1 | cut -c1-260
timeout 600 python bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/s29/bench_ref_n$N.log 2> gpurun_out/s29/bench_ref_n$N.err
grep '^{' gpurun_out/s29/bench_ref_n$N.log | tail -1 | cut -c1-300
