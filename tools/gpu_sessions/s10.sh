set -x
mkdir -p gpurun_out/s10
nvidia-smi -L > gpurun_out/s10/gpus.txt
timeout 300 python tools/group_bench.py --gpus 2 --steps 500 > gpurun_out/s10/group_n2.json 2> gpurun_out/s10/group_n2.err
timeout 300 python tools/group_bench.py --gpus 1 --steps 500 > gpurun_out/s10/group_n1.json 2> gpurun_out/s10/group_n1.err
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 500 --warmup 100 > gpurun_out/s10/bench_n2.log 2> gpurun_out/s10/bench_n2.err
grep "^{" gpurun_out/s10/bench_n2.log > gpurun_out/s10/bench_n2.json
timeout 300 python -m pytest tests/test_group.py tests/test_gpu_canary.py -q -m gpu > gpurun_out/s10/pytest.log 2>&1; tail -3 gpurun_out/s10/pytest.log
cat gpurun_out/s10/group_n2.json gpurun_out/s10/group_n1.json; grep -c "NCCL INFO" gpurun_out/s10/bench_n2.log; tail -2 gpurun_out/s10/*.err
