mkdir -p gpurun_out/s21
timeout 1200 python -m pytest tests/test_dual_solvers.py tests/test_binary_model.py -m gpu -q 2>&1 | tail -15
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/s21/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s21/pytest_gpu.log
tail -8 gpurun_out/s21/pytest_gpu.log
