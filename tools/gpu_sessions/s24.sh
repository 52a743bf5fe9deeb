mkdir -p gpurun_out/s24
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/s24/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s24/pytest_gpu.log
tail -8 gpurun_out/s24/pytest_gpu.log
