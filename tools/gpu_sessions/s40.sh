mkdir -p gpurun_out/s40
timeout 3000 python -m pytest tests -m gpu -q > gpurun_out/s40/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s40/pytest_gpu.log
tail -8 gpurun_out/s40/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --steps 300 --warmup 100 --no-cpu-baseline > gpurun_out/s40/bench_cheetah.json 2> gpurun_out/s40/bench_cheetah.err; tail -1 gpurun_out/s40/bench_cheetah.json | cut -c1-200
timeout 300 python bench.py --config humanoid --steps 200 --warmup 100 --no-cpu-baseline --no-e2e > gpurun_out/s40/bench_humanoid.json 2> /dev/null; tail -1 gpurun_out/s40/bench_humanoid.json | cut -c1-200
