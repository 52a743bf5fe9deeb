mkdir -p gpurun_out/s30
timeout 300 python bench.py --steps 300 --warmup 100 --no-cpu-baseline > gpurun_out/s30/bench_cheetah.json 2> gpurun_out/s30/bench_cheetah.err; tail -1 gpurun_out/s30/bench_cheetah.json | cut -c1-220
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/s30/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s30/pytest_gpu.log
tail -6 gpurun_out/s30/pytest_gpu.log
