mkdir -p gpurun_out/s37
timeout 1500 python -m pytest tests/test_fluid.py tests/test_zoo_parity.py tests/test_golden.py tests/test_gpu_coop.py tests/test_binary_model.py -m gpu -q -k "fluid or zoo_r or coop or binary" > gpurun_out/s37/pytest_fluid.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s37/pytest_fluid.log
tail -15 gpurun_out/s37/pytest_fluid.log
timeout 300 python bench.py --steps 300 --warmup 100 --no-cpu-baseline > gpurun_out/s37/bench_cheetah.json 2> gpurun_out/s37/bench_cheetah.err; tail -1 gpurun_out/s37/bench_cheetah.json | cut -c1-200
