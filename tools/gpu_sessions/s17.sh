mkdir -p gpurun_out/s17
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s17/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s17/pytest_gpu.log
tail -5 gpurun_out/s17/pytest_gpu.log
timeout 300 python bench.py --steps 200 --warmup 100 > gpurun_out/s17/bench_cheetah.json 2> gpurun_out/s17/bench_cheetah.err; tail -1 gpurun_out/s17/bench_cheetah.json | cut -c1-600
timeout 300 python bench.py --config humanoid --steps 200 --warmup 100 --no-cpu-baseline > gpurun_out/s17/bench_humanoid.json 2> gpurun_out/s17/bench_humanoid.err; tail -1 gpurun_out/s17/bench_humanoid.json | cut -c1-400
